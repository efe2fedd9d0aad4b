// RoPE + KV append and the two attention flavours of the hot path, CUDA-core flash-decoding style:
//   * FlowLM: causal attention of each new row over the sequence's paged KV (decode: one row per sequence;
//     prefill: all rows of the prompt).  Replaces modules/attention.py:29-64,150-182 + modules/rope.py.
//   * Mimi: 16-step chunk against the 250-slot ring buffer with the reference's write-before-attend
//     visibility rule.  Replaces modules/attention.py:67-105,220-264.
// Work split: 8 lanes own one key (16-byte slices of the 64-wide head), so a warp covers 4 keys per step,
// keeps an online-softmax state per 8-lane group and merges groups/warps at the end.
#include "kernels.cuh"

namespace ptts {
namespace {

template <typename KT> struct KVec;
template <> struct KVec<__nv_bfloat16> {
  static __device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};
template <> struct KVec<float> {
  static __device__ __forceinline__ void load8(const float* p, float (&o)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void store2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
  }
};

// ---- FlowLM -------------------------------------------------------------------------------------------
// grid = M rows, block = H*32 threads: thread (h, i) rotates pair i of head h for q and k, copies v.
template <typename KT>
__global__ void flow_rope_append_kernel(const FlowAttnParams p) {
  const int m = blockIdx.x;
  const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
  const int D = p.H * kHeadDim;
  const int seq = p.row_seq ? p.row_seq[m] : m;
  const int pos = p.row_pos[m];
  const float* row = p.qkv + (long long)m * 3 * D;
  float sn, cs;
  sincosf((float)pos * p.freqs[i], &sn, &cs);
  const int off = h * kHeadDim + 2 * i;
  const float qr = row[off], qi = row[off + 1];
  const float kr = row[D + off], ki = row[D + off + 1];
  const float vr = row[2 * D + off], vi = row[2 * D + off + 1];
  p.q_rot[(long long)m * D + off] = qr * cs - qi * sn;
  p.q_rot[(long long)m * D + off + 1] = qr * sn + qi * cs;
  const int page = p.page_table[(long long)seq * p.max_pages + pos / kPageTokens];
  const int slot = pos % kPageTokens;
  KT* base = reinterpret_cast<KT*>(p.pool) + p.layer * p.layer_stride + page * p.page_stride;
  const long long kv_half = (long long)p.H * kPageTokens * kHeadDim;
  KT* kdst = base + ((long long)h * kPageTokens + slot) * kHeadDim + 2 * i;
  KVec<KT>::store2(kdst, kr * cs - ki * sn, kr * sn + ki * cs);
  KVec<KT>::store2(kdst + kv_half, vr, vi);
}

struct SoftState {
  float m, l, acc[8];
};

__device__ __forceinline__ void soft_init(SoftState& s) {
  s.m = -INFINITY; s.l = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s.acc[i] = 0.f;
}

__device__ __forceinline__ void soft_update(SoftState& s, float score, const float (&v)[8]) {
  const float mn = fmaxf(s.m, score);
  const float corr = __expf(s.m - mn);      // exp(-inf) = 0 on the first key
  const float pexp = __expf(score - mn);
  s.l = s.l * corr + pexp;
#pragma unroll
  for (int i = 0; i < 8; ++i) s.acc[i] = s.acc[i] * corr + pexp * v[i];
  s.m = mn;
}

// merge the state of the lane `delta` away (same head-slice index) into this lane's
__device__ __forceinline__ void soft_merge_shfl(SoftState& s, int delta) {
  const float om = __shfl_xor_sync(0xffffffffu, s.m, delta);
  const float ol = __shfl_xor_sync(0xffffffffu, s.l, delta);
  const float mn = fmaxf(s.m, om);
  const float ca = (s.m == -INFINITY) ? 0.f : __expf(s.m - mn);
  const float cb = (om == -INFINITY) ? 0.f : __expf(om - mn);
  s.l = s.l * ca + ol * cb;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float oa = __shfl_xor_sync(0xffffffffu, s.acc[i], delta);
    s.acc[i] = s.acc[i] * ca + oa * cb;
  }
  s.m = mn;
}

// grid (M, H), block 128 = 4 warps; keys 0..pos of the row's sequence are strided over 16 lane-groups.
template <typename KT>
__global__ void __launch_bounds__(128) flow_attention_kernel(const FlowAttnParams p) {
  __shared__ float sh_m[4], sh_l[4], sh_acc[4][kHeadDim];
  const int m = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sl = lane & 7;     // key group within the warp, 8-wide slice of the head
  const int D = p.H * kHeadDim;
  const int seq = p.row_seq ? p.row_seq[m] : m;
  const int n_keys = p.row_pos[m] + 1;
  float q[8];
  {
    const float* qp = p.q_rot + (long long)m * D + h * kHeadDim + sl * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = qp[i] * 0.125f;   // 1/sqrt(64)
  }
  const KT* pool = reinterpret_cast<const KT*>(p.pool) + p.layer * p.layer_stride;
  const long long kv_half = (long long)p.H * kPageTokens * kHeadDim;
  const int* pt = p.page_table + (long long)seq * p.max_pages;
  SoftState st;
  soft_init(st);
  for (int k0 = (warp * 4); k0 < n_keys; k0 += 16) {
    const int key = k0 + grp;
    const bool ok = key < n_keys;
    float kf[8], vf[8];
    float s = 0.f;
    if (ok) {
      const int page = pt[key / kPageTokens];
      const KT* kp = pool + page * p.page_stride + ((long long)h * kPageTokens + (key % kPageTokens)) * kHeadDim + sl * 8;
      KVec<KT>::load8(kp, kf);
      KVec<KT>::load8(kp + kv_half, vf);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) soft_update(st, s, vf);
  }
  soft_merge_shfl(st, 8);
  soft_merge_shfl(st, 16);
  if (lane < 8) {
    if (lane == 0) { sh_m[warp] = st.m; sh_l[warp] = st.l; }
#pragma unroll
    for (int i = 0; i < 8; ++i) sh_acc[warp][sl * 8 + i] = st.acc[i];
  }
  __syncthreads();
  if (threadIdx.x < kHeadDim) {
    float mx = fmaxf(fmaxf(sh_m[0], sh_m[1]), fmaxf(sh_m[2], sh_m[3]));
    float l = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = (sh_m[w] == -INFINITY) ? 0.f : __expf(sh_m[w] - mx);
      l += sh_l[w] * c;
      a += sh_acc[w][threadIdx.x] * c;
    }
    const long long oi = (long long)m * D + h * kHeadDim + threadIdx.x;
    if (p.out16) p.out16[oi] = __float2bfloat16_rn(a / l);
    else p.out[oi] = a / l;
  }
}

// ---- Mimi ---------------------------------------------------------------------------------------------
// grid = B*T rows, block = H*32
template <typename KT>
__global__ void mimi_rope_ring_kernel(const MimiAttnParams p) {
  const int m = blockIdx.x;
  const int b = m / p.T, t = m % p.T;
  const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
  const int D = p.H * kHeadDim;
  const int e = p.offset[b];
  const int pos = e + t;
  const float* row = p.qkv + (long long)m * 3 * D;
  float sn, cs;
  sincosf((float)pos * p.freqs[i], &sn, &cs);
  const int off = h * kHeadDim + 2 * i;
  const float qr = row[off], qi = row[off + 1];
  const float kr = row[D + off], ki = row[D + off + 1];
  const float vr = row[2 * D + off], vi = row[2 * D + off + 1];
  p.q_rot[(long long)m * D + off] = qr * cs - qi * sn;
  p.q_rot[(long long)m * D + off + 1] = qr * sn + qi * cs;
  const int slot = pos % p.context;
  KT* base = reinterpret_cast<KT*>(p.ring) + p.layer * p.layer_stride;
  KT* kdst = base + (((long long)b * p.H + h) * p.context + slot) * kHeadDim + 2 * i;
  KVec<KT>::store2(kdst, kr * cs - ki * sn, kr * sn + ki * cs);
  KVec<KT>::store2(kdst + p.kv_stride, vr, vi);
}

// grid (B, H), block 128: warp w owns queries 4w..4w+3 of the 16-step chunk (T <= 16).
template <typename KT>
__global__ void __launch_bounds__(128) mimi_attention_kernel(const MimiAttnParams p) {
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sl = lane & 7;
  const int D = p.H * kHeadDim;
  const int cap = p.context;
  const int e = p.offset[b];
  const int last = e + p.T - 1;
  const int end_index = last % cap;
  const KT* kbase = reinterpret_cast<const KT*>(p.ring) + p.layer * p.layer_stride +
                    ((long long)b * p.H + h) * cap * kHeadDim;
  const KT* vbase = kbase + p.kv_stride;
  constexpr int QW = 4;
  float q[QW][8];
  SoftState st[QW];
  int qpos[QW];
#pragma unroll
  for (int j = 0; j < QW; ++j) {
    const int t = warp * QW + j;
    qpos[j] = (t < p.T) ? e + t : -1000000;
    soft_init(st[j]);
    const float* qp = p.q_rot + ((long long)(b * p.T + min(t, p.T - 1))) * D + h * kHeadDim + sl * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) q[j][i] = qp[i] * 0.125f;
  }
  for (int s0 = 0; s0 < cap; s0 += 4) {
    const int slot = s0 + grp;
    const bool in = slot < cap;
    // position held by this ring slot after the chunk was written (attention.py:88-103)
    int pos_k = -1;
    if (in) {
      const int delta = slot - end_index;
      pos_k = delta <= 0 ? last + delta : last + delta - cap;
      if (slot >= e + p.T) pos_k = -1;
    }
    float kf[8], vf[8];
    const bool any = in && pos_k >= 0;
    if (any) {
      KVec<KT>::load8(kbase + (long long)slot * kHeadDim + sl * 8, kf);
      KVec<KT>::load8(vbase + (long long)slot * kHeadDim + sl * 8, vf);
    }
#pragma unroll
    for (int j = 0; j < QW; ++j) {
      float s = 0.f;
      if (any) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[j][i], kf[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int dq = qpos[j] - pos_k;
      if (any && dq >= 0 && dq < cap) soft_update(st[j], s, vf);
    }
  }
#pragma unroll
  for (int j = 0; j < QW; ++j) {
    soft_merge_shfl(st[j], 8);
    soft_merge_shfl(st[j], 16);
    const int t = warp * QW + j;
    if (lane < 8 && t < p.T) {
      const long long oi = ((long long)(b * p.T + t)) * D + h * kHeadDim + sl * 8;
      const float inv = 1.0f / st[j].l;
      if (p.out16) {
#pragma unroll
        for (int i = 0; i < 8; i += 2)
          *reinterpret_cast<__nv_bfloat162*>(p.out16 + oi + i) = __floats2bfloat162_rn(st[j].acc[i] * inv, st[j].acc[i + 1] * inv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) p.out[oi + i] = st[j].acc[i] * inv;
      }
    }
  }
}

}  // namespace

void launch_flow_rope_append(const FlowAttnParams& p, cudaStream_t s) {
  if (p.M <= 0) return;
  ProfScope ps("flow_rope_append", nullptr, 0, (double)p.M * p.H * 64 * (3 * 4 + 4 + 2 * (p.kv_bf16 ? 2 : 4)), s);
  if (p.kv_bf16) flow_rope_append_kernel<__nv_bfloat16><<<p.M, p.H * 32, 0, s>>>(p);
  else flow_rope_append_kernel<float><<<p.M, p.H * 32, 0, s>>>(p);
  ++g_launches;
}

void launch_flow_attention(const FlowAttnParams& p, cudaStream_t s) {
  if (p.M <= 0) return;
  dim3 grid(p.M, p.H);
  ProfScope ps("flow_attention", nullptr, 4.0 * p.total_keys * p.H * 64,
               2.0 * p.total_keys * p.H * 64 * (p.kv_bf16 ? 2 : 4) + 2.0 * p.M * p.H * 64 * 4, s);
  if (p.kv_bf16) flow_attention_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p);
  else flow_attention_kernel<float><<<grid, 128, 0, s>>>(p);
  ++g_launches;
}

void launch_mimi_rope_ring(const MimiAttnParams& p, cudaStream_t s) {
  ProfScope ps("mimi_rope_ring", nullptr, 0, (double)p.B * p.T * p.H * 64 * (3 * 4 + 4 + 2 * (p.kv_bf16 ? 2 : 4)), s);
  if (p.kv_bf16) mimi_rope_ring_kernel<__nv_bfloat16><<<p.B * p.T, p.H * 32, 0, s>>>(p);
  else mimi_rope_ring_kernel<float><<<p.B * p.T, p.H * 32, 0, s>>>(p);
  ++g_launches;
}

void launch_mimi_attention(const MimiAttnParams& p, cudaStream_t s) {
  dim3 grid(p.B, p.H);
  ProfScope ps("mimi_attention", nullptr, 4.0 * p.B * p.H * p.T * p.context * 64,
               2.0 * p.B * p.H * p.context * 64 * (p.kv_bf16 ? 2 : 4) + 2.0 * p.B * p.T * p.H * 64 * 4, s);
  if (p.kv_bf16) mimi_attention_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(p);
  else mimi_attention_kernel<float><<<grid, 128, 0, s>>>(p);
  ++g_launches;
}

}  // namespace ptts
