// RoPE + KV append and the two attention flavours of the hot path, CUDA-core flash-decoding style:
//   * FlowLM: causal attention of each new row over the sequence's paged KV (decode: one row per sequence;
//     prefill: all rows of the prompt).  Replaces modules/attention.py:29-64,150-182 + modules/rope.py.
//   * Mimi: 16-step chunk against the 250-slot ring buffer with the reference's write-before-attend
//     visibility rule.  Replaces modules/attention.py:67-105,220-264.
// Work split: 8 lanes own one key (16-byte slices of the 64-wide head), so a warp covers 4 keys per step,
// keeps an online-softmax state per 8-lane group and merges groups/warps at the end.
#include "kernels.cuh"
#include "tc_device.cuh"

#include <algorithm>
#include <cstdlib>

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
  pdl_sync();
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
__global__ void __launch_bounds__(128, 16) flow_attention_kernel(const FlowAttnParams p) {
  pdl_sync();
  __shared__ float sh_m[4], sh_l[4], sh_acc[4][kHeadDim];
  const int m = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sl = lane & 7;     // key group within the warp, 8-wide slice of the head
  const int D = p.H * kHeadDim;
  const int seq = p.row_seq ? p.row_seq[m] : m;
  const int n_all = p.row_pos[m] + 1;
  // split-KV (small batches): CTA z of `splits` owns a 16-aligned slice of the keys and writes a partial
  int key_lo = p.prefix_len, n_keys = n_all;     // prefix_len > 0: keys [0, prefix_len) come from the cascade partial
  if (p.splits > 1) {
    const int chunk = (((n_all + p.splits - 1) / p.splits) + 15) & ~15;
    key_lo = blockIdx.z * chunk;
    n_keys = min(n_all, key_lo + chunk);
  }
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
  // each warp takes 16 consecutive keys per step (4 groups of 4, all inside one 32-token page): the 8 16-byte
  // loads of a step are issued before any of them is consumed, so ~8 KB per warp are in flight
  constexpr int U = 1;   // measured: U=2 no gain, U=4 (108 regs) 40% slower -- occupancy beats per-warp unrolling here
  // steps start on a multiple of 4 keys so that the keys of a step share one page; keys below key_lo (the
  // cascade prefix, or another split's slice) are masked
  for (int k0 = (key_lo & ~3) + warp * 4 * U; k0 < n_keys; k0 += 16 * U) {
    const int page = pt[k0 / kPageTokens];
    const KT* pbase = pool + page * p.page_stride + ((long long)h * kPageTokens + (k0 % kPageTokens)) * kHeadDim + sl * 8;
    float kf[U][8], vf[U][8];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int key = k0 + 4 * u + grp;
      ok[u] = key >= key_lo && key < n_keys;
      if (ok[u]) {
        const KT* kp = pbase + (4 * u + grp) * kHeadDim;
        KVec<KT>::load8(kp, kf[u]);
        KVec<KT>::load8(kp + kv_half, vf[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = 0.f;
      if (ok[u]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[u][i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok[u]) soft_update(st, s, vf[u]);
    }
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
    const float* pp = (p.prefix_len > 0) ? p.prefix_part + ((long long)m * p.H + h) * 66 : nullptr;
    if (pp) mx = fmaxf(mx, pp[64]);
    float l = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = (sh_m[w] == -INFINITY) ? 0.f : __expf(sh_m[w] - mx);
      l += sh_l[w] * c;
      a += sh_acc[w][threadIdx.x] * c;
    }
    if (pp) {
      const float c = __expf(pp[64] - mx);
      l += pp[65] * c;
      a += pp[threadIdx.x] * c;
    }
    if (p.splits > 1) {
      // partial of this key slice: [m][h][z][64 values | max | sum]; an empty slice contributes l = 0
      float* part = p.part + (((long long)m * p.H + h) * p.splits + blockIdx.z) * 66;
      part[threadIdx.x] = a;
      if (threadIdx.x == 0) { part[64] = mx; part[65] = l; }
    } else {
      const long long oi = (long long)m * D + h * kHeadDim + threadIdx.x;
      if (p.out16) p.out16[oi] = __float2bfloat16_rn(a / l);
      else p.out[oi] = a / l;
    }
  }
}

// merge the split-KV partials of one (row, head): grid (M, H), 64 threads
__global__ void flow_attention_merge_kernel(const FlowAttnParams p) {
  pdl_sync();
  const int m = blockIdx.x, h = blockIdx.y;
  const float* part = p.part + ((long long)m * p.H + h) * p.splits * 66;
  float mx = -INFINITY;
  for (int z = 0; z < p.splits; ++z)
    if (part[z * 66 + 65] > 0.f) mx = fmaxf(mx, part[z * 66 + 64]);
  float l = 0.f, a = 0.f;
  for (int z = 0; z < p.splits; ++z) {
    const float lz = part[z * 66 + 65];
    if (lz > 0.f) {
      const float c = __expf(part[z * 66 + 64] - mx);
      l += lz * c;
      a += part[z * 66 + threadIdx.x] * c;
    }
  }
  const long long oi = (long long)m * p.H * kHeadDim + h * kHeadDim + threadIdx.x;
  if (p.out16) p.out16[oi] = __float2bfloat16_rn(a / l);
  else p.out[oi] = a / l;
}

// ---- Mimi ---------------------------------------------------------------------------------------------
// grid = B*T rows, block = H*32
template <typename KT>
__global__ void mimi_rope_ring_kernel(const MimiAttnParams p) {
  pdl_sync();
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
  pdl_sync();
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


// ---- Mimi attention on tensor cores (bf16 ring) --------------------------------------------------------------
// One CTA per (sequence, head); the 16-step chunk is exactly one m16 MMA tile.  Warp w owns ring slots
// [64w, 64w+64): it stages its K/V slice in shared memory with coalesced 16-byte loads (the only HBM traffic of
// the kernel), computes S = Q K^T and O = P V with mma.sync.m16n8k16 (ldmatrix / ldmatrix.trans operands),
// keeps its own softmax statistics and the four partial results are merged through shared memory.  The
// visibility rule is the reference's (modules/attention.py:88-103, 244-254).
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kMimiLd = 72;                                   // padded smem row (bf16 elements): conflict-free ldmatrix
constexpr int kMimiWarpSmem = 2 * 64 * kMimiLd * 2;           // bytes per warp: K + V slices

__global__ void __launch_bounds__(128) mimi_attention_mma_kernel(const MimiAttnParams p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char mimi_smem[];
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int D = p.H * kHeadDim;
  const int cap = p.context;
  const int e = p.offset[b];
  const int last = e + p.T - 1;
  const int end_index = last % cap;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(mimi_smem + warp * kMimiWarpSmem);
  __nv_bfloat16* Vs = Ks + 64 * kMimiLd;
  const __nv_bfloat16* kbase = reinterpret_cast<const __nv_bfloat16*>(p.ring) + p.layer * p.layer_stride +
                               ((long long)b * p.H + h) * cap * kHeadDim;
  const __nv_bfloat16* vbase = kbase + p.kv_stride;
  const int slot0 = warp * 64;
  // ---- stage this warp's 64 slots of K and V (16-byte chunks, fully coalesced) ----
  // cp.async (16 B, L1-bypassing) straight into shared memory: all 32 copies of a lane are in flight at once and
  // no registers are spent on staging; slots beyond the ring are zero-filled (src-size 0)
  // (the ring is read once per frame and layer: L2 evict-first, see tc_device.cuh)
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int idx = lane + 32 * i;
    const int row = idx >> 3, ch = idx & 7;
    const int slot = slot0 + row;
    const int sz = slot < cap ? 16 : 0;
    const long long off = (long long)(slot < cap ? slot : 0) * kHeadDim + ch * 8;
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"((uint32_t)__cvta_generic_to_shared(Ks + row * kMimiLd + ch * 8)),
                 "l"(kbase + off), "r"(sz), "l"(pol) : "memory");
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"((uint32_t)__cvta_generic_to_shared(Vs + row * kMimiLd + ch * 8)),
                 "l"(vbase + off), "r"(sz), "l"(pol) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // ---- Q fragments (rows g and g+8 of the chunk), pre-scaled by 1/sqrt(64) ----
  uint32_t qa[4][4];
  {
    const bool r0 = g < p.T, r1 = g + 8 < p.T;
    const float* q0 = p.q_rot + ((long long)(b * p.T + (r0 ? g : 0))) * D + h * kHeadDim;
    const float* q1 = p.q_rot + ((long long)(b * p.T + (r1 ? g + 8 : 0))) * D + h * kHeadDim;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = 16 * ks + 2 * t4;
      const float2 a0 = *reinterpret_cast<const float2*>(q0 + c), a2 = *reinterpret_cast<const float2*>(q0 + c + 8);
      const float2 a1 = *reinterpret_cast<const float2*>(q1 + c), a3 = *reinterpret_cast<const float2*>(q1 + c + 8);
      qa[ks][0] = r0 ? pack2_bf16(a0.x * 0.125f, a0.y * 0.125f) : 0u;
      qa[ks][1] = r1 ? pack2_bf16(a1.x * 0.125f, a1.y * 0.125f) : 0u;
      qa[ks][2] = r0 ? pack2_bf16(a2.x * 0.125f, a2.y * 0.125f) : 0u;
      qa[ks][3] = r1 ? pack2_bf16(a3.x * 0.125f, a3.y * 0.125f) : 0u;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  // ---- S = Q K^T : 8 key tiles x 4 k-steps ----
  float sc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t kb[4];
      ldsm_x4(kb, (uint32_t)__cvta_generic_to_shared(Ks + (nt * 8 + (lane & 7)) * kMimiLd + 32 * kk + 8 * (lane >> 3)));
      mma_bf16_16816(sc[nt], qa[2 * kk], kb[0], kb[1]);
      mma_bf16_16816(sc[nt], qa[2 * kk + 1], kb[2], kb[3]);
    }
  }
  // ---- visibility mask + per-row max over this warp's slots ----
  const int qp0 = (g < p.T) ? e + g : -(1 << 30), qp1 = (g + 8 < p.T) ? e + g + 8 : -(1 << 30);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int slot = slot0 + nt * 8 + 2 * t4 + j;
      int pos_k = -1;
      if (slot < cap) {
        const int delta = slot - end_index;
        pos_k = delta <= 0 ? last + delta : last + delta - cap;
        if (slot >= e + p.T) pos_k = -1;
      }
      const int d0 = qp0 - pos_k, d1 = qp1 - pos_k;
      const bool v0 = pos_k >= 0 && d0 >= 0 && d0 < cap, v1 = pos_k >= 0 && d1 >= 0 && d1 < cap;
      if (!v0) sc[nt][j] = -INFINITY;
      if (!v1) sc[nt][2 + j] = -INFINITY;
      m0 = fmaxf(m0, sc[nt][j]);
      m1 = fmaxf(m1, sc[nt][2 + j]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float mm0 = (m0 == -INFINITY) ? 0.f : m0, mm1 = (m1 == -INFINITY) ? 0.f : m1;
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[4][4];                       // P as A fragments: k-step j covers slots 16j..16j+15 of the slice
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float p00 = __expf(sc[nt][0] - mm0), p01 = __expf(sc[nt][1] - mm0);
    const float p10 = __expf(sc[nt][2] - mm1), p11 = __expf(sc[nt][3] - mm1);
    l0 += p00 + p01;
    l1 += p10 + p11;
    pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p00, p01);
    pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p10, p11);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- O = P V : 8 dim tiles x 4 k-steps ----
  float oc[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {       // pairs of dim tiles
      uint32_t vb[4];
      const int mid = lane >> 3;
      ldsm_x4_t(vb, (uint32_t)__cvta_generic_to_shared(Vs + (16 * j + 8 * (mid & 1) + (lane & 7)) * kMimiLd +
                                                        8 * (2 * dp + (mid >> 1))));
      mma_bf16_16816(oc[2 * dp], pa[j], vb[0], vb[1]);
      mma_bf16_16816(oc[2 * dp + 1], pa[j], vb[2], vb[3]);
    }
  }
  // ---- merge the four slices ----
  __syncwarp();
  float* sO = reinterpret_cast<float*>(mimi_smem + warp * kMimiWarpSmem);      // [16][64]
  float* sM = sO + 16 * 64;                                                      // [16] max, [16] sum
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    *reinterpret_cast<float2*>(sO + g * 64 + dt * 8 + 2 * t4) = make_float2(oc[dt][0], oc[dt][1]);
    *reinterpret_cast<float2*>(sO + (g + 8) * 64 + dt * 8 + 2 * t4) = make_float2(oc[dt][2], oc[dt][3]);
  }
  if (t4 == 0) {
    sM[g] = m0; sM[g + 8] = m1;
    sM[16 + g] = l0; sM[16 + g + 8] = l1;
  }
  __syncthreads();
  {
    const int row = threadIdx.x >> 3, c0 = (threadIdx.x & 7) * 8;      // 16 rows x 8 column groups
    if (row < p.T) {
      float mx = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mx = fmaxf(mx, reinterpret_cast<const float*>(mimi_smem + w * kMimiWarpSmem)[16 * 64 + row]);
      float l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float* base = reinterpret_cast<const float*>(mimi_smem + w * kMimiWarpSmem);
        const float mw = base[16 * 64 + row];
        const float cw = (mw == -INFINITY) ? 0.f : __expf(mw - mx);
        l += base[16 * 64 + 16 + row] * cw;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += base[row * 64 + c0 + i] * cw;
      }
      const float inv = 1.0f / l;
      const long long oi = ((long long)(b * p.T + row)) * D + h * kHeadDim + c0;
      if (p.out16) {
        *reinterpret_cast<uint4*>(p.out16 + oi) =
            make_uint4(pack2_bf16(acc[0] * inv, acc[1] * inv), pack2_bf16(acc[2] * inv, acc[3] * inv),
                       pack2_bf16(acc[4] * inv, acc[5] * inv), pack2_bf16(acc[6] * inv, acc[7] * inv));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) p.out[oi + i] = acc[i] * inv;
      }
    }
  }
}

// ---- cascade: attention of every row over the SHARED voice prefix, on tensor cores ---------------------------
// grid (ceil(M/64), H), 128 threads.  The prefix K/V of head h (<= 128 keys, read from the voice's own pages)
// are staged once per CTA; warp w takes rows 16w..16w+15 of the CTA's 64 as one m16 MMA tile against all 128
// key slots (S = Q K^T: 16 n-tiles x 4 k-steps, O = P V: 8 dim tiles x 8 k-steps) and writes the un-normalised
// partial {O[64], max, sum} that flow_attention_kernel merges with the per-sequence keys.
constexpr int kPrefixKeys = 128;
__global__ void __launch_bounds__(128) flow_prefix_attention_kernel(const FlowAttnParams p) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char pre_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(pre_smem);
  __nv_bfloat16* Vs = Ks + kPrefixKeys * kMimiLd;
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int D = p.H * kHeadDim;
  const int P = p.prefix_len;
  const __nv_bfloat16* pool = reinterpret_cast<const __nv_bfloat16*>(p.pool) + p.layer * p.layer_stride;
  const long long kv_half = (long long)p.H * kPageTokens * kHeadDim;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + 128 * i;                  // 128 keys x 8 chunks of 16 bytes
    const int key = idx >> 3, ch = idx & 7;
    uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
    if (key < P) {
      const int page = p.prefix_pages[key / kPageTokens];
      const __nv_bfloat16* src = pool + page * p.page_stride + ((long long)h * kPageTokens + (key % kPageTokens)) * kHeadDim + ch * 8;
      kv = *reinterpret_cast<const uint4*>(src);
      vv = *reinterpret_cast<const uint4*>(src + kv_half);
    }
    *reinterpret_cast<uint4*>(Ks + key * kMimiLd + ch * 8) = kv;
    *reinterpret_cast<uint4*>(Vs + key * kMimiLd + ch * 8) = vv;
  }
  const int row0 = blockIdx.x * 64 + warp * 16;
  const int m0 = row0 + g, m1 = row0 + g + 8;
  const bool r0 = m0 < p.M, r1 = m1 < p.M;
  uint32_t qa[4][4];
  {
    const float* q0 = p.q_rot + (long long)(r0 ? m0 : 0) * D + h * kHeadDim;
    const float* q1 = p.q_rot + (long long)(r1 ? m1 : 0) * D + h * kHeadDim;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = 16 * ks + 2 * t4;
      const float2 a0 = *reinterpret_cast<const float2*>(q0 + c), a2 = *reinterpret_cast<const float2*>(q0 + c + 8);
      const float2 a1 = *reinterpret_cast<const float2*>(q1 + c), a3 = *reinterpret_cast<const float2*>(q1 + c + 8);
      qa[ks][0] = pack2_bf16(a0.x * 0.125f, a0.y * 0.125f);
      qa[ks][1] = pack2_bf16(a1.x * 0.125f, a1.y * 0.125f);
      qa[ks][2] = pack2_bf16(a2.x * 0.125f, a2.y * 0.125f);
      qa[ks][3] = pack2_bf16(a3.x * 0.125f, a3.y * 0.125f);
    }
  }
  __syncthreads();
  float sc[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t kb[4];
      ldsm_x4(kb, (uint32_t)__cvta_generic_to_shared(Ks + (nt * 8 + (lane & 7)) * kMimiLd + 32 * kk + 8 * (lane >> 3)));
      mma_bf16_16816(sc[nt], qa[2 * kk], kb[0], kb[1]);
      mma_bf16_16816(sc[nt], qa[2 * kk + 1], kb[2], kb[3]);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (nt * 8 + 2 * t4 + j >= P) { sc[nt][j] = -INFINITY; sc[nt][2 + j] = -INFINITY; }
      mx0 = fmaxf(mx0, sc[nt][j]);
      mx1 = fmaxf(mx1, sc[nt][2 + j]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
  uint32_t pa[8][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    const float p00 = __expf(sc[nt][0] - mx0), p01 = __expf(sc[nt][1] - mx0);
    const float p10 = __expf(sc[nt][2] - mx1), p11 = __expf(sc[nt][3] - mx1);
    l0 += p00 + p01;
    l1 += p10 + p11;
    pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p00, p01);
    pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p10, p11);
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  float oc[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t vb[4];
      const int mid = lane >> 3;
      ldsm_x4_t(vb, (uint32_t)__cvta_generic_to_shared(Vs + (16 * j + 8 * (mid & 1) + (lane & 7)) * kMimiLd +
                                                        8 * (2 * dp + (mid >> 1))));
      mma_bf16_16816(oc[2 * dp], pa[j], vb[0], vb[1]);
      mma_bf16_16816(oc[2 * dp + 1], pa[j], vb[2], vb[3]);
    }
  }
  if (r0) {
    float* o = p.prefix_part + ((long long)m0 * p.H + h) * 66;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<float2*>(o + dt * 8 + 2 * t4) = make_float2(oc[dt][0], oc[dt][1]);
    if (t4 == 0) { o[64] = mx0; o[65] = l0; }
  }
  if (r1) {
    float* o = p.prefix_part + ((long long)m1 * p.H + h) * 66;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<float2*>(o + dt * 8 + 2 * t4) = make_float2(oc[dt][2], oc[dt][3]);
    if (t4 == 0) { o[64] = mx1; o[65] = l1; }
  }
}


// ---- decode attention as a persistent stream of TMA boxes + tensor-core math ------------------------------------
// ncu on the per-(row, head) CUDA-core kernel above: 28.8 M warp instructions per launch, ~31 per key -- it is bound by
// instruction issue at about half the issue peak, not by HBM (62 % of the measured bandwidth), and three dependent
// global loads sit in front of every short-lived CTA.  Two persistent warp-specialised versions with one producer warp
// feeding four consumer warps through a 12-stage ring were SLOWER (61-70 us vs 52): ncu showed the single producer
// never waiting for a free stage (it was the bottleneck at ~1000 cycles of mbarrier / shuffle / TMA-issue overhead per
// 8 KB) and the consumers meeting at a CTA barrier per item.  This version keeps the idea and removes both:
//   * a CTA is ONE consumer warp + ONE producer warp (64 threads), 6 CTAs per SM, persistent; a CTA owns a contiguous
//     range of (row, head) items, so nothing is ever merged across warps and there is no CTA-wide barrier;
//   * the producer moves 64 keys per stage: each 32-token page of the head is ONE 5-D TMA box (K and V slices, 8 KB,
//     128-byte swizzle); only the first / last page of an item, where [key_lo, n_all) cuts the page, go as 8-key
//     boxes; two 16 KB stages per CTA, so ~100-190 KB per SM are in flight;
//   * the consumer works on 64 keys at a time with mma.sync m16n8k16 (the query is row 0 of the 16-row A tile):
//     S = q K^T (32 MMAs, ldmatrix on the swizzled rows), one online-softmax update on the 4 lanes that hold row 0,
//     O += P V (32 MMAs, ldmatrix.trans): ~3 warp instructions per key;
//   * the item's query row and meta data go through a two-slot side buffer; the cascade partial of the shared voice
//     prefix is fetched one item ahead and merged at the end of the item.
// Masked keys get p = 0 exactly; their V bytes are whatever the pool (or a stale stage) holds, which is why the pool
// and the ring are zeroed once (0 x finite = 0; never-written memory could hold NaN patterns).
constexpr int kAttnPageBytes = 2 * kPageTokens * kHeadDim * 2;    // K + V of one page and head, bf16: 8 KB
constexpr int kAttnThreads = 64;
constexpr int attn_smem_bytes(int stages, int pg) { return stages * pg * kAttnPageBytes + 2 * 256 + 2 * 16 + 256 + 16 * stages + 32 + 1024; }

// The private keys / values of a sequence are read once per frame and layer and never again before ~1.5 GB of other
// keys have gone by: the boxes carry an L2 evict-first policy so that they do not flush the weights and activations
// that the GEMMs around the attention re-read every frame (FlowAttnParams::kv_evict_first).
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4,
                                            unsigned long long policy = 0ull) {
  if (policy)
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "l"(policy) : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// One ring stage (PG pages = 32 * PG keys of one head) against the query tile: S = Q K^T, online softmax in the log2
// domain, O += P V.  TWO = false: the tile holds ONE query (row 0: lanes 0..3 carry it, a1 / a3 are zero, element [1] of
// mrun / lrun is unused); TWO = true: 16 queries (rows g and g + 8 of every lane), used for the shared voice prefix.
template <int PG, bool TWO>
__device__ __forceinline__ void attn_stage(uint32_t st0, bool two_pages, const uint32_t (&qa)[4][4], int k0, int key_lo, int n_all,
                                           float (&oc)[8][4], float (&mrun)[2], float (&lrun)[2], int lane) {
  const int t4 = lane & 3;
  float sc[4 * PG][4];
#pragma unroll
  for (int e = 0; e < PG; ++e) {
    if (e == 1 && !two_pages) break;
    const uint32_t Ks = st0 + (uint32_t)e * kAttnPageBytes;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float (&acc)[4] = sc[4 * e + nt];
      acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
      const int r = nt * 8 + (lane & 7);                     // key row this lane addresses for ldmatrix
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t kb[4];
        const int c = 4 * kk + (lane >> 3);                   // 16-byte chunk (8 dims) of the row
        ldsm_x4(kb, Ks + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
        mma_bf16_16816(acc, qa[2 * kk], kb[0], kb[1]);
        mma_bf16_16816(acc, qa[2 * kk + 1], kb[2], kb[3]);
      }
    }
  }
  if (k0 < key_lo || k0 + PG * kPageTokens > n_all) {        // warp-uniform: only an item's first / last stage is cut
#pragma unroll
    for (int nt = 0; nt < 4 * PG; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = k0 + nt * 8 + 2 * t4 + e;
        if (key < key_lo || key >= n_all) {                   // (covers an absent second page: its keys are >= n_all)
          sc[nt][e] = -INFINITY;
          if (TWO) sc[nt][2 + e] = -INFINITY;
        }
      }
    }
  }
  float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 4 * PG; ++nt) {
    bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
    if (TWO) bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
  }
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
  if (TWO) {
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
  }
  const float nm0 = fmaxf(mrun[0], bm0);                       // finite: every stage holds at least one valid key
  const float corr0 = (mrun[0] == -INFINITY) ? 0.f : exp2f(mrun[0] - nm0);
  mrun[0] = nm0;
  float nm1 = 0.f, corr1 = 0.f;
  if (TWO) {
    nm1 = fmaxf(mrun[1], bm1);
    corr1 = (mrun[1] == -INFINITY) ? 0.f : exp2f(mrun[1] - nm1);
    mrun[1] = nm1;
  }
  float s0 = 0.f, s1 = 0.f;
  uint32_t pa[2 * PG][4];
#pragma unroll
  for (int nt = 0; nt < 4 * PG; ++nt) {
    const float p0 = exp2f(sc[nt][0] - nm0), p1 = exp2f(sc[nt][1] - nm0);
    s0 += p0 + p1;
    pa[nt >> 1][(nt & 1) * 2] = pack2_bf16(p0, p1);
    if (TWO) {
      const float q0 = exp2f(sc[nt][2] - nm1), q1 = exp2f(sc[nt][3] - nm1);
      s1 += q0 + q1;
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(q0, q1);
    } else {
      pa[nt >> 1][(nt & 1) * 2 + 1] = 0u;                      // rows 8..15 of the A tile
    }
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  lrun[0] = lrun[0] * corr0 + s0;
  if (TWO) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    lrun[1] = lrun[1] * corr1 + s1;
  }
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    oc[dt][0] *= corr0; oc[dt][1] *= corr0;
    if (TWO) { oc[dt][2] *= corr1; oc[dt][3] *= corr1; }
  }
  // O += P V: per page 2 k-steps of 16 keys x 8 n-tiles of 8 dims
#pragma unroll
  for (int e = 0; e < PG; ++e) {
    if (e == 1 && !two_pages) break;
    const uint32_t Vs = st0 + (uint32_t)e * kAttnPageBytes + 4096u;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int mid = lane >> 3;
        const int r = 16 * jj + 8 * (mid & 1) + (lane & 7);
        const int c = 2 * dp + (mid >> 1);
        ldsm_x4_t(vb, Vs + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4));
        mma_bf16_16816(oc[2 * dp], pa[2 * e + jj], vb[0], vb[1]);
        mma_bf16_16816(oc[2 * dp + 1], pa[2 * e + jj], vb[2], vb[3]);
      }
    }
  }
}

// PG = pages (32 keys) per stage; CTAs per SM follow the stage size: 2 x 2 pages -> 6, 3 x 2 -> 4, 2 x 1 -> 9
template <int STAGES, int PG, bool FOLD>
__global__ void __launch_bounds__(kAttnThreads, PG == 1 ? 9 : (STAGES == 2 ? 6 : 4)) flow_attention_stream_kernel(const __grid_constant__ CUtensorMap tm_page,
                                                                                const __grid_constant__ CUtensorMap tm_box,
                                                                                const FlowAttnParams p, const int items) {
  pdl_sync();
  if (p.tstamp && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(p.tstamp + 2 * p.layer, t);
  }
  extern __shared__ __align__(1024) unsigned char attn_smem[];
  constexpr int kAttnStageBytes = PG * kAttnPageBytes;
  const uint32_t ring = (smem_u32(attn_smem) + 1023u) & ~1023u;
  const uint32_t qbuf = ring + STAGES * kAttnStageBytes;            // 2 x 64 floats
  const uint32_t metab = qbuf + 2 * 256;                            // 2 x {n_all, key_lo, pg_lo, n_pg}
  const uint32_t obuf = metab + 2 * 16;                             // 64 floats: the item's un-normalised output row
  const uint32_t bars = obuf + 256;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, qfull0 = empty0 + 8 * STAGES, qempty0 = qfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, D = H * kHeadDim;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_page) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_box) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(qfull0 + 8 * i, 1);
      mbar_init(qempty0 + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // pages / boxes that are never loaded (keys outside [key_lo, n_all)) must still hold finite values
  for (uint32_t i = threadIdx.x; i < (uint32_t)(STAGES * kAttnStageBytes / 16); i += kAttnThreads)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ring + i * 16u), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // contiguous, balanced item range of this CTA
  const int it_lo = (int)(((long long)blockIdx.x * items) / gridDim.x);
  const int it_hi = (int)(((long long)(blockIdx.x + 1) * items) / gridDim.x);
  const int page0 = p.layer * (int)(p.layer_stride / p.page_stride);     // page coordinate of this layer's page 0
  // folded cascade (p.pflags): the shared-prefix partials are computed by this kernel too, as (16 rows, head) tiles with
  // real 16-row MMAs.  Tiles are CLAIMED (atomic counter p.pflags[0]) by whichever CTAs are running, so a tile always
  // belongs to a resident CTA and the per-sequence items that later wait for its flag (p.pflags[1 + tile]) cannot
  // deadlock on a CTA that has not been scheduled.  This launch also zeroes the NEXT layer's counter and flags (the
  // last layer those of layer 0, for the next frame): launches of one batch are ordered on its stream.
  const bool fold = FOLD && p.pflags != nullptr && p.prefix_len > 0;     // (compiled out of the default instantiation)
  const int n_pitems = fold ? ((p.M + 15) / 16) * H : 0;
  if (fold && blockIdx.x == 0)
    for (int i = threadIdx.x; i <= n_pitems; i += kAttnThreads) p.pflags_next[i] = 0;

  if (warp == 1) {
    // ---------------- producer ----------------
    uint32_t gst = 0, qcnt = 0;
    // folded cascade: (16 rows, head) tiles against the shared voice prefix first (pages of the voice itself, L2-resident);
    // the claimed tile id travels to the consumer through the query ring's metadata slot, -1 ends the phase
    if (fold) {
      const int n_ppg = (p.prefix_len + kPageTokens - 1) / kPageTokens;
      const int my_page = lane < n_ppg ? p.prefix_pages[lane] : 0;             // <= 4 pages (prefix <= 128 keys)
      for (;;) {
        int pi = 0;
        if (lane == 0) pi = atomicAdd(p.pflags, 1);
        pi = __shfl_sync(0xffffffffu, pi, 0);
        const bool end = pi >= n_pitems;
        const uint32_t qs = qcnt & 1, qph = (qcnt >> 1) & 1;
        if (lane == 0) {
          mbar_wait(qempty0 + 8 * qs, qph ^ 1);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(metab + 16 * qs), "r"(end ? -1 : pi), "r"(-1), "r"(0), "r"(n_ppg) : "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(qfull0 + 8 * qs) : "memory");
        }
        ++qcnt;
        if (end) break;
        const int h = pi % H;
        for (int j0 = 0; j0 < n_ppg; j0 += PG, ++gst) {
          const uint32_t s = gst % STAGES, ph = (gst / STAGES) & 1;
          int pages[PG];
#pragma unroll
          for (int e = 0; e < PG; ++e) pages[e] = __shfl_sync(0xffffffffu, my_page, (j0 + e) & 31);
          if (lane == 0) {
            uint32_t bytes = 0;
            int bhi[PG];
#pragma unroll
            for (int e = 0; e < PG; ++e) {
              const int k0 = (j0 + e) * kPageTokens;
              bhi[e] = (j0 + e < n_ppg) ? (min(p.prefix_len, k0 + kPageTokens) - 1 - k0) >> 3 : -1;
              bytes += (uint32_t)(bhi[e] + 1) * 2048u;
            }
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_expect_tx(full0 + 8 * s, bytes);
#pragma unroll
            for (int e = 0; e < PG; ++e) {
              const uint32_t dst = ring + s * kAttnStageBytes + (uint32_t)e * kAttnPageBytes;
              const int pc = page0 + pages[e];
              if (bhi[e] == 3) {
                tma_load_5d(dst, &tm_page, full0 + 8 * s, 0, 0, h, 0, pc);
              } else {
                for (int b = 0; b <= bhi[e]; ++b) {
                  tma_load_5d(dst + (uint32_t)b * 1024u, &tm_box, full0 + 8 * s, 0, 8 * b, h, 0, pc);
                  tma_load_5d(dst + 4096u + (uint32_t)b * 1024u, &tm_box, full0 + 8 * s, 0, 8 * b, h, 1, pc);
                }
              }
            }
          }
          __syncwarp();
        }
      }
    }
    unsigned long long kv_policy = 0ull;
    if (p.kv_evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(kv_policy));
    const int pg_first = p.prefix_len / kPageTokens;
    // lane l holds pages pg_first + l and pg_first + 32 + l of a row (2048 keys); rows beyond that reload in the loop.
    // Entries past the row's last page are zero in the table, so the two loads do not depend on each other.
    auto fetch = [&](int m, int& n_all, int& pa, int& pb) {
      const int seq = p.row_seq ? p.row_seq[m] : m;
      const int* pt = p.page_table + (long long)seq * p.max_pages;
      n_all = p.row_pos[m] + 1;
      pa = (pg_first + lane < p.max_pages) ? pt[pg_first + lane] : 0;
      pb = (pg_first + 32 + lane < p.max_pages) ? pt[pg_first + 32 + lane] : 0;
    };
    int m_cur = -1, cur_n = 0, cur_a = 0, cur_b = 0, nx_n = 0, nx_a = 0, nx_b = 0;
    if (it_lo < it_hi) fetch(it_lo / H, nx_n, nx_a, nx_b);
    for (int it = it_lo; it < it_hi; ++it) {
      const int m = it / H, h = it - m * H;
      if (m != m_cur) {
        m_cur = m; cur_n = nx_n; cur_a = nx_a; cur_b = nx_b;
        if ((m + 1) * H < it_hi) fetch(m + 1, nx_n, nx_a, nx_b);     // consumed a whole row (H items) later
      }
      const int n_all = cur_n;
      const int key_lo = p.prefix_len;
      const int pg_lo = pg_first, pg_hi = (n_all - 1) / kPageTokens;
      const int n_pg = pg_hi - pg_lo + 1;
      const uint32_t qs = qcnt & 1, qph = (qcnt >> 1) & 1;
      if (lane == 0) {
        mbar_wait(qempty0 + 8 * qs, qph ^ 1);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(metab + 16 * qs), "r"(n_all), "r"(key_lo), "r"(pg_lo), "r"(n_pg) : "memory");
        mbar_expect_tx(qfull0 + 8 * qs, 256);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                     ::"r"(qbuf + 256 * qs), "l"(p.q_rot + (long long)m * D + h * kHeadDim), "r"(qfull0 + 8 * qs) : "memory");
      }
      ++qcnt;
      const int seq = p.row_seq ? p.row_seq[m] : m;
      int far = 0;
      for (int j0 = 0; j0 < n_pg; j0 += PG, ++gst) {
        const uint32_t s = gst % STAGES, ph = (gst / STAGES) & 1;
        int pages[PG];
#pragma unroll
        for (int e = 0; e < PG; ++e) {
          const int j = j0 + e;
          if (j >= 64 && (j & 31) == 0) {     // more than 64 pages (> 2048 keys): further page ids, fetched in place
            const int pg = pg_lo + j + lane;
            far = (pg <= pg_hi) ? p.page_table[(long long)seq * p.max_pages + pg] : 0;
          }
          pages[e] = __shfl_sync(0xffffffffu, j < 32 ? cur_a : (j < 64 ? cur_b : far), j & 31);
        }
        if (lane == 0) {
          uint32_t bytes = 0;
          int blo[PG], bhi[PG];
#pragma unroll
          for (int e = 0; e < PG; ++e) {
            const int k0 = (pg_lo + j0 + e) * kPageTokens;
            blo[e] = (max(key_lo, k0) - k0) >> 3;                                    // 8-key boxes that hold valid keys
            bhi[e] = (j0 + e < n_pg) ? (min(n_all, k0 + kPageTokens) - 1 - k0) >> 3 : -1;
            if (bhi[e] >= blo[e]) bytes += (uint32_t)(bhi[e] - blo[e] + 1) * 2048u;
          }
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, bytes);
#pragma unroll
          for (int e = 0; e < PG; ++e) {
            const uint32_t dst = ring + s * kAttnStageBytes + (uint32_t)e * kAttnPageBytes;
            const int pc = page0 + pages[e];
            if (blo[e] == 0 && bhi[e] == 3) {
              tma_load_5d(dst, &tm_page, full0 + 8 * s, 0, 0, h, 0, pc, kv_policy);  // K and V of the whole page
            } else {
              for (int b = blo[e]; b <= bhi[e]; ++b) {
                tma_load_5d(dst + (uint32_t)b * 1024u, &tm_box, full0 + 8 * s, 0, 8 * b, h, 0, pc, kv_policy);
                tma_load_5d(dst + 4096u + (uint32_t)b * 1024u, &tm_box, full0 + 8 * s, 0, 8 * b, h, 1, pc, kv_policy);
              }
            }
          }
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- consumer ----------------
  const int g = lane >> 2, t4 = lane & 3;
  uint32_t qcnt = 0, gst = 0;
  constexpr float kQs = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) and log2(e) folded into q: scores in the log2 domain
  if (fold) {
    const int n_ppg = (p.prefix_len + kPageTokens - 1) / kPageTokens;
    for (;;) {
      const uint32_t qs = qcnt & 1, qph = (qcnt >> 1) & 1;
      mbar_wait(qfull0 + 8 * qs, qph);
      int pi;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pi) : "r"(metab + 16 * qs));
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(qempty0 + 8 * qs) : "memory");
      ++qcnt;
      if (pi < 0) break;
      const int rg = pi / H, h = pi - rg * H;
      const int m0 = rg * 16 + g, m1 = m0 + 8;
      const bool r0 = m0 < p.M, r1 = m1 < p.M;
      uint32_t qa[4][4];
      {
        const float* q0 = p.q_rot + (long long)(r0 ? m0 : 0) * D + h * kHeadDim;
        const float* q1 = p.q_rot + (long long)(r1 ? m1 : 0) * D + h * kHeadDim;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int c = 16 * ks + 2 * t4;
          const float2 a0 = *reinterpret_cast<const float2*>(q0 + c), a2 = *reinterpret_cast<const float2*>(q0 + c + 8);
          const float2 a1 = *reinterpret_cast<const float2*>(q1 + c), a3 = *reinterpret_cast<const float2*>(q1 + c + 8);
          qa[ks][0] = r0 ? pack2_bf16(a0.x * kQs, a0.y * kQs) : 0u;
          qa[ks][1] = r1 ? pack2_bf16(a1.x * kQs, a1.y * kQs) : 0u;
          qa[ks][2] = r0 ? pack2_bf16(a2.x * kQs, a2.y * kQs) : 0u;
          qa[ks][3] = r1 ? pack2_bf16(a3.x * kQs, a3.y * kQs) : 0u;
        }
      }
      float oc[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f;
      float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
      for (int j0 = 0; j0 < n_ppg; j0 += PG, ++gst) {
        const uint32_t s = gst % STAGES, ph = (gst / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        attn_stage<PG, true>(ring + s * kAttnStageBytes, PG == 2 && j0 + 1 < n_ppg, qa, j0 * kPageTokens, 0, p.prefix_len, oc, mrun,
                             lrun, lane);
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
      }
      // un-normalised partial {O[64], max (natural log domain), sum} per (row, head), as flow_prefix_attention_kernel writes it
      if (r0) {
        float* o = p.prefix_part + ((long long)m0 * H + h) * 66;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<float2*>(o + dt * 8 + 2 * t4) = make_float2(oc[dt][0], oc[dt][1]);
        if (t4 == 0) { o[64] = mrun[0] * 0.6931471805599453f; o[65] = lrun[0]; }
      }
      if (r1) {
        float* o = p.prefix_part + ((long long)m1 * H + h) * 66;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) *reinterpret_cast<float2*>(o + dt * 8 + 2 * t4) = make_float2(oc[dt][2], oc[dt][3]);
        if (t4 == 0) { o[64] = mrun[1] * 0.6931471805599453f; o[65] = lrun[1]; }
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p.pflags + 1 + pi), "r"(1) : "memory");
    }
  }
  const bool has_pp = p.prefix_len > 0;
  // not folded: the partial comes from flow_prefix_attention_kernel (an earlier launch) and is fetched one item ahead
  float2 nx_a = make_float2(0.f, 0.f);
  float nx_m = -INFINITY, nx_l = 0.f;
  if (has_pp && !fold && it_lo < it_hi) {
    const float* pp = p.prefix_part + (long long)it_lo * 66;
    nx_a = *reinterpret_cast<const float2*>(pp + 2 * lane); nx_m = pp[64]; nx_l = pp[65];
  }
  for (int it = it_lo; it < it_hi; ++it, ++qcnt) {
    const int m = it / H, h = it - m * H;
    const uint32_t qs = qcnt & 1, qph = (qcnt >> 1) & 1;
    float2 pp_a = nx_a;
    float pp_m = nx_m, pp_l = nx_l;
    bool pp_ready = has_pp && !fold;
    const float* ppc = p.prefix_part + (long long)it * 66;
    const int* my_flag = fold ? p.pflags + 1 + (m >> 4) * H + h : nullptr;
    if (fold) {
      // published already (the usual case after the first microseconds)?  then its loads ride under this item
      int f;
      asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(f) : "l"(my_flag) : "memory");
      pp_ready = f != 0;
      if (pp_ready) { pp_a = __ldcg(reinterpret_cast<const float2*>(ppc + 2 * lane)); pp_m = __ldcg(ppc + 64); pp_l = __ldcg(ppc + 65); }
    } else if (has_pp && it + 1 < it_hi) {
      const float* pp = p.prefix_part + (long long)(it + 1) * 66;
      nx_a = *reinterpret_cast<const float2*>(pp + 2 * lane); nx_m = pp[64]; nx_l = pp[65];
    }
    mbar_wait(qfull0 + 8 * qs, qph);
    uint32_t qa[4][4];           // A fragments of row 0 (a0, a2); rows 8..15 (a1, a3) are zero
    int n_all, key_lo, pg_lo, n_pg;
    {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        float2 x, y;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x.x), "=f"(x.y) : "r"(qbuf + 256 * qs + (uint32_t)(16 * ks + 2 * t4) * 4u));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(y.x), "=f"(y.y) : "r"(qbuf + 256 * qs + (uint32_t)(16 * ks + 8 + 2 * t4) * 4u));
        qa[ks][0] = g == 0 ? pack2_bf16(x.x * kQs, x.y * kQs) : 0u;
        qa[ks][1] = 0u;
        qa[ks][2] = g == 0 ? pack2_bf16(y.x * kQs, y.y * kQs) : 0u;
        qa[ks][3] = 0u;
      }
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(n_all), "=r"(key_lo), "=r"(pg_lo), "=r"(n_pg) : "r"(metab + 16 * qs));
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(qempty0 + 8 * qs) : "memory");
    float oc[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f;
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
    for (int j0 = 0; j0 < n_pg; j0 += PG, ++gst) {
      const uint32_t s = gst % STAGES, ph = (gst / STAGES) & 1;
      mbar_wait(full0 + 8 * s, ph);
      attn_stage<PG, false>(ring + s * kAttnStageBytes, PG == 2 && j0 + 1 < n_pg, qa, (pg_lo + j0) * kPageTokens, key_lo, n_all, oc,
                            mrun, lrun, lane);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
    }
    // row 0 of the tile lives in lanes 0..3: through shared memory to "lane l owns dims 2l, 2l+1"
    if (g == 0) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(obuf + (uint32_t)(dt * 8 + 2 * t4) * 4u), "f"(oc[dt][0]), "f"(oc[dt][1]) : "memory");
    }
    const float m_own = __shfl_sync(0xffffffffu, mrun[0], 0) * 0.6931471805599453f;   // back to the natural-log domain
    const float l_own = __shfl_sync(0xffffffffu, lrun[0], 0);
    __syncwarp();
    float2 o;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o.x), "=f"(o.y) : "r"(obuf + (uint32_t)lane * 8u));
    __syncwarp();
    float l = l_own;
    if (has_pp) {
      if (!pp_ready) {           // folded cascade, tile not published when this item started: wait for it now
        int f;
        do {
          asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(f) : "l"(my_flag) : "memory");
        } while (f == 0);
        pp_a = __ldcg(reinterpret_cast<const float2*>(ppc + 2 * lane)); pp_m = __ldcg(ppc + 64); pp_l = __ldcg(ppc + 65);
      }
      const float mx = fmaxf(m_own, pp_m);
      const float c0 = __expf(m_own - mx), c1 = __expf(pp_m - mx);
      l = l_own * c0 + pp_l * c1;
      o.x = o.x * c0 + pp_a.x * c1;
      o.y = o.y * c0 + pp_a.y * c1;
    }
    const float inv = 1.0f / l;
    const long long oi = (long long)m * D + h * kHeadDim + 2 * lane;
    if (p.out16) *reinterpret_cast<__nv_bfloat162*>(p.out16 + oi) = __floats2bfloat162_rn(o.x * inv, o.y * inv);
    else *reinterpret_cast<float2*>(p.out + oi) = make_float2(o.x * inv, o.y * inv);
  }
  if (p.tstamp && lane == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(p.tstamp + 2 * p.layer + 1, t);
  }
}

}  // namespace

// cos / sin of every row's position, shared by the 6 layers of a step: table[m] = cos[32] | sin[32]
// (angles in fp32 like modules/rope.py:17-24; consumed by the fused qkv epilogue of the tcgen05 GEMM)
// T > 1: row m = (sequence m / T, step m % T) at position row_pos[m / T] + m % T (Mimi: 16 steps per frame)
__global__ void rope_table_kernel(const int* __restrict__ row_pos, const float* __restrict__ freqs,
                                  float* __restrict__ table, int M, int T) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 32) return;
  const int m = idx >> 5, i = idx & 31;
  const int pos = (T > 1) ? row_pos[m / T] + m % T : row_pos[m];
  float sn, cs;
  sincosf((float)pos * freqs[i], &sn, &cs);
  table[(long long)m * 64 + i] = cs;
  table[(long long)m * 64 + 32 + i] = sn;
}

void launch_rope_table(const int* row_pos, const float* freqs, float* table, int M, int T, cudaStream_t s) {
  if (M <= 0) return;
  ProfScope ps("rope_table", nullptr, 0, (double)M * 64 * 4, s);
  launch_k(rope_table_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, s, row_pos, freqs, table, M, T);
  ++g_launches;
}

void launch_flow_rope_append(const FlowAttnParams& p, cudaStream_t s) {
  if (p.M <= 0) return;
  ProfScope ps("flow_rope_append", nullptr, 0, (double)p.M * p.H * 64 * (3 * 4 + 4 + 2 * (p.kv_bf16 ? 2 : 4)), s);
  if (p.kv_bf16) launch_k(flow_rope_append_kernel<__nv_bfloat16>, dim3(p.M), dim3(p.H * 32), 0, s, p);
  else launch_k(flow_rope_append_kernel<float>, dim3(p.M), dim3(p.H * 32), 0, s, p);
  ++g_launches;
}

void launch_flow_prefix_attention(const FlowAttnParams& p, cudaStream_t s) {
  if (p.M <= 0 || p.prefix_len <= 0) return;
  const size_t smem = (size_t)2 * kPrefixKeys * kMimiLd * 2;
  ProfScope ps("flow_prefix_attention", nullptr, 4.0 * p.M * p.H * p.prefix_len * 64,
               2.0 * p.H * p.prefix_len * 64 * 2 + (double)p.M * p.H * (64 * 4 + 66 * 4), s);
  launch_k(flow_prefix_attention_kernel, dim3((p.M + 63) / 64, p.H), dim3(128), smem, s, p);
  ++g_launches;
}

unsigned long long* flow_attention_dbg_buffer() {
  static const bool on = [] { const char* v = getenv("PTTS_ATTN_DBG"); return v && v[0] == '1'; }();
  static unsigned long long* buf = nullptr;
  if (on && !buf) cudaMalloc((void**)&buf, 64 * 2 * sizeof(unsigned long long));
  return on ? buf : nullptr;
}

// bf16 decode rows, enough (row, head) items to fill the machine: the persistent bulk-copy kernel
static bool flow_attention_stream_ok(const FlowAttnParams& p) {
  static const int mode = [] { const char* v = getenv("PTTS_ATTN_STREAM"); return v ? atoi(v) : 1; }();
  return mode != 0 && p.kv_tmap && p.kv_bf16 && p.splits <= 1 && !p.row_seq && p.M * p.H >= 296 && p.q_rot && (p.out16 || p.out);
}

// the folded-cascade instantiation exists for the default stage geometry only
bool flow_attention_streams(const FlowAttnParams& p) {
  const char* st = getenv("PTTS_ATTN_STAGES");
  const char* pg = getenv("PTTS_ATTN_PAGES");
  return flow_attention_stream_ok(p) && !(st && atoi(st) == 3) && !(pg && atoi(pg) == 1);
}

void launch_flow_attention(const FlowAttnParams& p, cudaStream_t s) {
  if (p.M <= 0) return;
  dim3 grid(p.M, p.H, p.splits > 1 ? p.splits : 1);
  const double keys = (double)p.total_keys - (double)p.M * p.prefix_len;     // keys streamed by this kernel
  ProfScope ps("flow_attention", nullptr, 4.0 * keys * p.H * 64,
               2.0 * keys * p.H * 64 * (p.kv_bf16 ? 2 : 4) + 2.0 * p.M * p.H * 64 * 4, s);
  if (flow_attention_stream_ok(p)) {
    // stage geometry: PTTS_ATTN_STAGES (2 | 3 ring stages) x PTTS_ATTN_PAGES (2 | 1 pages of 32 keys per stage)
    static const int stages = [] { const char* v = getenv("PTTS_ATTN_STAGES"); return (v && atoi(v) == 3) ? 3 : 2; }();
    static const int pg = [] { const char* v = getenv("PTTS_ATTN_PAGES"); return (v && atoi(v) == 1) ? 1 : 2; }();
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(flow_attention_stream_kernel<2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(2, 2));
      cudaFuncSetAttribute(flow_attention_stream_kernel<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(2, 2));
      cudaFuncSetAttribute(flow_attention_stream_kernel<3, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(3, 2));
      cudaFuncSetAttribute(flow_attention_stream_kernel<2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(2, 1));
      cudaFuncSetAttribute(flow_attention_stream_kernel<3, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes(3, 1));
      attr_done = true;
    }
    const int items = p.M * p.H;
    static const int per_sm = [] { const char* v = getenv("PTTS_ATTN_CTAS_PER_SM"); return v ? std::max(1, atoi(v)) : 0; }();
    const int ctas = per_sm > 0 ? per_sm : (pg == 1 ? 9 : (stages == 2 ? 6 : 4));
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(p.kv_tmap);
    const dim3 grid((unsigned)std::min(items, 148 * ctas)), block(kAttnThreads);
    const size_t smem = (size_t)attn_smem_bytes(stages, pg);
    const bool fold = p.pflags != nullptr && p.prefix_len > 0;       // only with the default stage geometry
    if (fold && stages == 2 && pg == 2) launch_k(flow_attention_stream_kernel<2, 2, true>, grid, block, smem, s, maps[0], maps[1], p, items);
    else if (stages == 3 && pg == 2) launch_k(flow_attention_stream_kernel<3, 2, false>, grid, block, smem, s, maps[0], maps[1], p, items);
    else if (stages == 2 && pg == 1) launch_k(flow_attention_stream_kernel<2, 1, false>, grid, block, smem, s, maps[0], maps[1], p, items);
    else if (stages == 3 && pg == 1) launch_k(flow_attention_stream_kernel<3, 1, false>, grid, block, smem, s, maps[0], maps[1], p, items);
    else launch_k(flow_attention_stream_kernel<2, 2, false>, grid, block, smem, s, maps[0], maps[1], p, items);
    ++g_launches;
    return;
  }
  if (p.kv_bf16) launch_k(flow_attention_kernel<__nv_bfloat16>, dim3(grid), dim3(128), 0, s, p);
  else launch_k(flow_attention_kernel<float>, dim3(grid), dim3(128), 0, s, p);
  ++g_launches;
  if (p.splits > 1) {
    launch_k(flow_attention_merge_kernel, dim3(p.M, p.H), dim3(64), 0, s, p);
    ++g_launches;
  }
}

// ---- FlowLM prefill attention on tensor cores ---------------------------------------------------------------
// Text / voice prefill: a sequence contributes T rows at positions pos0 .. pos0+T-1 that attend causally to the paged
// keys 0 .. pos (modules/attention.py:164-182 with T > 1).  The per-row decode kernel reads every key once per ROW;
// here a CTA owns 64 query rows of one (sequence, head): the keys are staged 64 at a time in shared memory (cp.async
// from the pages, 4 KB per page and head) and S = Q K^T, online softmax and O += P V run on mma.sync m16n8k16 in
// the FlashAttention-2 arrangement (4 warps x 16 rows, running max / sum per row, O rescaled per key block).
__global__ void __launch_bounds__(128) flow_prefill_attention_kernel(const FlowAttnParams p) {
  pdl_sync();
  __shared__ __align__(16) __nv_bfloat16 Ks[64 * kMimiLd];
  __shared__ __align__(16) __nv_bfloat16 Vs[64 * kMimiLd];
  const int sq = blockIdx.x, h = blockIdx.y, qt = blockIdx.z;
  const int row_lo = p.seq_row0[sq], T = p.seq_row0[sq + 1] - row_lo;
  if (qt * 64 >= T) return;
  const int pos0 = p.seq_pos0[sq];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int D = p.H * kHeadDim;
  const __nv_bfloat16* pool = reinterpret_cast<const __nv_bfloat16*>(p.pool) + p.layer * p.layer_stride;
  const long long kv_half = (long long)p.H * kPageTokens * kHeadDim;
  const int* pt = p.page_table + (long long)sq * p.max_pages;
  const int tl0 = qt * 64 + warp * 16 + g, tl1 = tl0 + 8;         // local row indices of this lane's two rows
  const bool r0 = tl0 < T, r1 = tl1 < T;
  const int qpos0 = pos0 + tl0, qpos1 = pos0 + tl1;
  uint32_t qa[4][4];
  {
    const float* q0 = p.q_rot + (long long)(row_lo + (r0 ? tl0 : 0)) * D + h * kHeadDim;
    const float* q1 = p.q_rot + (long long)(row_lo + (r1 ? tl1 : 0)) * D + h * kHeadDim;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = 16 * ks + 2 * t4;
      const float2 a0 = *reinterpret_cast<const float2*>(q0 + c), a2 = *reinterpret_cast<const float2*>(q0 + c + 8);
      const float2 a1 = *reinterpret_cast<const float2*>(q1 + c), a3 = *reinterpret_cast<const float2*>(q1 + c + 8);
      qa[ks][0] = pack2_bf16(a0.x * 0.125f, a0.y * 0.125f);
      qa[ks][1] = pack2_bf16(a1.x * 0.125f, a1.y * 0.125f);
      qa[ks][2] = pack2_bf16(a2.x * 0.125f, a2.y * 0.125f);
      qa[ks][3] = pack2_bf16(a3.x * 0.125f, a3.y * 0.125f);
    }
  }
  float oc[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int n_keys = pos0 + min(T, qt * 64 + 64);                  // keys any row of this tile may see
  for (int k0 = 0; k0 < n_keys; k0 += 64) {
    __syncthreads();                                               // previous block fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 128 * i;                       // 64 keys x 8 chunks of 16 bytes
      const int kr = idx >> 3, ch = idx & 7;
      const int key = k0 + kr;
      const int sz = key < n_keys ? 16 : 0;
      const int kk = key < n_keys ? key : 0;
      const __nv_bfloat16* src = pool + pt[kk / kPageTokens] * p.page_stride +
                                 ((long long)h * kPageTokens + (kk % kPageTokens)) * kHeadDim + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(Ks + kr * kMimiLd + ch * 8)),
                   "l"(src), "r"(sz) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(Vs + kr * kMimiLd + ch * 8)),
                   "l"(src + kv_half), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t kb[4];
        ldsm_x4(kb, (uint32_t)__cvta_generic_to_shared(Ks + (nt * 8 + (lane & 7)) * kMimiLd + 32 * kk + 8 * (lane >> 3)));
        mma_bf16_16816(sc[nt], qa[2 * kk], kb[0], kb[1]);
        mma_bf16_16816(sc[nt], qa[2 * kk + 1], kb[2], kb[3]);
      }
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int key = k0 + nt * 8 + 2 * t4 + j;
        if (key > qpos0 || !r0) sc[nt][j] = -INFINITY;             // causal: key position <= query position
        if (key > qpos1 || !r1) sc[nt][2 + j] = -INFINITY;
        bm0 = fmaxf(bm0, sc[nt][j]);
        bm1 = fmaxf(bm1, sc[nt][2 + j]);
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float e0 = (nm0 == -INFINITY) ? 0.f : nm0, e1 = (nm1 == -INFINITY) ? 0.f : nm1;
    const float c0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - e0), c1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - e1);
    m0 = nm0; m1 = nm1;
    float s0 = 0.f, s1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p00 = __expf(sc[nt][0] - e0), p01 = __expf(sc[nt][1] - e0);
      const float p10 = __expf(sc[nt][2] - e1), p11 = __expf(sc[nt][3] - e1);
      s0 += p00 + p01;
      s1 += p10 + p11;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p00, p01);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p10, p11);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    l0 = l0 * c0 + s0;
    l1 = l1 * c1 + s1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { oc[dt][0] *= c0; oc[dt][1] *= c0; oc[dt][2] *= c1; oc[dt][3] *= c1; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const int mid = lane >> 3;
        ldsm_x4_t(vb, (uint32_t)__cvta_generic_to_shared(Vs + (16 * j + 8 * (mid & 1) + (lane & 7)) * kMimiLd +
                                                          8 * (2 * dp + (mid >> 1))));
        mma_bf16_16816(oc[2 * dp], pa[j], vb[0], vb[1]);
        mma_bf16_16816(oc[2 * dp + 1], pa[j], vb[2], vb[3]);
      }
    }
  }
  const float i0 = r0 ? 1.0f / l0 : 0.f, i1 = r1 ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int d = dt * 8 + 2 * t4;
    if (r0) {
      const long long oi = (long long)(row_lo + tl0) * D + h * kHeadDim + d;
      if (p.out16) *reinterpret_cast<__nv_bfloat162*>(p.out16 + oi) = __floats2bfloat162_rn(oc[dt][0] * i0, oc[dt][1] * i0);
      else { p.out[oi] = oc[dt][0] * i0; p.out[oi + 1] = oc[dt][1] * i0; }
    }
    if (r1) {
      const long long oi = (long long)(row_lo + tl1) * D + h * kHeadDim + d;
      if (p.out16) *reinterpret_cast<__nv_bfloat162*>(p.out16 + oi) = __floats2bfloat162_rn(oc[dt][2] * i1, oc[dt][3] * i1);
      else { p.out[oi] = oc[dt][2] * i1; p.out[oi + 1] = oc[dt][3] * i1; }
    }
  }
}

bool launch_flow_prefill_attention(const FlowAttnParams& p, cudaStream_t s) {
  if (!p.kv_bf16 || !p.seq_row0 || !p.seq_pos0 || p.n_seq <= 0 || p.max_rows_per_seq <= 0) return false;
  ProfScope ps("flow_prefill_attention", nullptr, 4.0 * p.total_keys * p.H * 64, 2.0 * p.total_keys * p.H * 64 * 2 / 32, s);
  launch_k(flow_prefill_attention_kernel, dim3(p.n_seq, p.H, (p.max_rows_per_seq + 63) / 64), dim3(128), 0, s, p);
  ++g_launches;
  return true;
}

// ---- Mimi ENCODER transformer (voice cloning; one-off per voice, fp32, not a hot path) -----------------------
// Non-streaming call of MimiStreamingMultiheadAttention (modules/attention.py:210-264 with model_state=None):
// positions 0..T-1, interleaved-pair RoPE, key j visible to query t iff 0 <= t - j < context.
__global__ void enc_rope_kernel(float* __restrict__ qkv, const float* __restrict__ freqs, int T, int H) {
  const int t = blockIdx.x;
  const int h = threadIdx.x >> 5, i = threadIdx.x & 31;
  const int D = H * kHeadDim;
  float sn, cs;
  sincosf((float)t * freqs[i], &sn, &cs);
  float* row = qkv + (long long)t * 3 * D + h * kHeadDim + 2 * i;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float xr = row[which * D], xi = row[which * D + 1];
    row[which * D] = xr * cs - xi * sn;
    row[which * D + 1] = xr * sn + xi * cs;
  }
}

// grid (T, H), 128 threads: scores of the <= 256-key window in shared memory, two-pass softmax, P.V
__global__ void __launch_bounds__(128) enc_window_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                   int T, int H, int context) {
  __shared__ float sq[kHeadDim], sc[256], red[4], so[2][kHeadDim];
  const int t = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int D = H * kHeadDim;
  const int lo = max(0, t - context + 1), n = t - lo + 1;
  if (tid < kHeadDim) sq[tid] = qkv[(long long)t * 3 * D + h * kHeadDim + tid] * 0.125f;
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 128) {
    const float* k = qkv + (long long)(lo + j) * 3 * D + D + h * kHeadDim;
    float a = 0.f;
#pragma unroll 16
    for (int d = 0; d < kHeadDim; ++d) a = fmaf(sq[d], k[d], a);
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float l = 0.f;
  for (int j = tid; j < n; j += 128) {
    const float p = __expf(sc[j] - mx);
    sc[j] = p;
    l += p;
  }
  l = warp_sum(l);
  if ((tid & 31) == 0) red[tid >> 5] = l;
  __syncthreads();
  l = (red[0] + red[1]) + (red[2] + red[3]);
  const int d = tid & 63, half = tid >> 6;
  float a = 0.f;
  for (int j = half; j < n; j += 2) a = fmaf(sc[j], qkv[(long long)(lo + j) * 3 * D + 2 * D + h * kHeadDim + d], a);
  so[half][d] = a;
  __syncthreads();
  if (tid < kHeadDim) out[(long long)t * D + h * kHeadDim + tid] = (so[0][tid] + so[1][tid]) / l;
}

void launch_enc_attention(float* qkv, float* out, const float* freqs, int T, int H, int context, cudaStream_t s) {
  if (T <= 0) return;
  launch_k(enc_rope_kernel, dim3(T), dim3(H * 32), 0, s, qkv, freqs, T, H);
  launch_k(enc_window_attention_kernel, dim3(T, H), dim3(128), 0, s, (const float*)qkv, out, T, H, context);
  g_launches += 2;
}

void launch_mimi_rope_ring(const MimiAttnParams& p, cudaStream_t s) {
  ProfScope ps("mimi_rope_ring", nullptr, 0, (double)p.B * p.T * p.H * 64 * (3 * 4 + 4 + 2 * (p.kv_bf16 ? 2 : 4)), s);
  if (p.kv_bf16) launch_k(mimi_rope_ring_kernel<__nv_bfloat16>, dim3(p.B * p.T), dim3(p.H * 32), 0, s, p);
  else launch_k(mimi_rope_ring_kernel<float>, dim3(p.B * p.T), dim3(p.H * 32), 0, s, p);
  ++g_launches;
}

void launch_mimi_attention(const MimiAttnParams& p, cudaStream_t s) {
  dim3 grid(p.B, p.H);
  ProfScope ps("mimi_attention", nullptr, 4.0 * p.B * p.H * p.T * p.context * 64,
               2.0 * p.B * p.H * p.context * 64 * (p.kv_bf16 ? 2 : 4) + 2.0 * p.B * p.T * p.H * 64 * 4, s);
  if (p.kv_bf16 && p.T <= 16 && p.context <= 256) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(mimi_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kMimiWarpSmem);
      attr_done = true;
    }
    launch_k(mimi_attention_mma_kernel, dim3(grid), dim3(128), 4 * kMimiWarpSmem, s, p);
  } else if (p.kv_bf16) launch_k(mimi_attention_kernel<__nv_bfloat16>, dim3(grid), dim3(128), 0, s, p);
  else launch_k(mimi_attention_kernel<float>, dim3(grid), dim3(128), 0, s, p);
  ++g_launches;
}

}  // namespace ptts
