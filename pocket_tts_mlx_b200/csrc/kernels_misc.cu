// Row-wise normalisation and the small fused ops around the GEMMs (LayerNorm,
// Embedding, NaN->BOS, noise, quantizer + upsample, last SEANet conv, carried conv state).
#include <nvtx3/nvToolsExt.h>
#include "kernels.cuh"

namespace ptts {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// one warp per row (8 rows per CTA), row held in registers; two-pass mean / biased variance exactly as
// LayerNorm is defined, reductions by warp shuffles only.  C <= 1024, C % 128 == 0.
template <int VEC>   // float4 vectors per lane: C = 128 * VEC
__global__ void __launch_bounds__(256) layernorm_kernel(const NormParams p) {
  pdl_sync();
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= p.nb * p.T) return;
  const int b = m / p.T, t = m % p.T;
  const float4* x = reinterpret_cast<const float4*>(p.X + b * p.x_bs + t * p.x_rs);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = x[lane + 32 * i];
  if (p.acc_n > 0) {
    float4* xw = reinterpret_cast<float4*>(const_cast<float*>(p.X) + b * p.x_bs + t * p.x_rs);
#pragma unroll 4
    for (int k = 0; k < p.acc_n; ++k) {
      const float4* a = reinterpret_cast<const float4*>(p.acc + k * p.acc_stride + (long long)m * p.C);
      float4 q[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) q[i] = a[lane + 32 * i];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { v[i].x += q[i].x; v[i].y += q[i].y; v[i].z += q[i].z; v[i].w += q[i].w; }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) xw[lane + 32 * i] = v[i];
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)p.C + p.eps);
  const float4* sc = p.scale ? reinterpret_cast<const float4*>(p.scale + (long long)m * p.mod_rs) : nullptr;
  const float4* sh = p.shift ? reinterpret_cast<const float4*>(p.shift + (long long)m * p.mod_rs) : nullptr;
  const float4* w4 = reinterpret_cast<const float4*>(p.w);
  const float4* b4 = reinterpret_cast<const float4*>(p.b);
  const long long yo = b * p.y_bs + t * p.y_rs;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c4 = lane + 32 * i;
    float4 o = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
    if (p.w) {
      const float4 w = w4[c4], bb = b4[c4];
      o.x = o.x * w.x + bb.x; o.y = o.y * w.y + bb.y; o.z = o.z * w.z + bb.z; o.w = o.w * w.w + bb.w;
    }
    if (sc) {
      const float4 a = sc[c4], d = sh[c4];
      o.x = o.x * (1.0f + a.x) + d.x; o.y = o.y * (1.0f + a.y) + d.y;
      o.z = o.z * (1.0f + a.z) + d.z; o.w = o.w * (1.0f + a.w) + d.w;
    }
    if (p.Y16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(p.Y16 + yo + 4 * c4) = pk;
    } else {
      *reinterpret_cast<float4*>(p.Y + yo + 4 * c4) = o;
    }
  }
}

// w_in_t is the transposed input_linear weight [L][D] so that consecutive threads read consecutive addresses.
// One thread = one output feature of 4 rows; all L weight loads are issued before the first FMA (the kernel is
// pure latency: 256 x 1024 outputs of a 32-deep dot product), grid = (rows / 4, D / 256).
template <int kL>
__global__ void __launch_bounds__(256) input_rows_kernel(const float* __restrict__ w_in_t, const float* __restrict__ bos,
                                                         const float* __restrict__ prev, const int* __restrict__ bos_flag,
                                                         float* __restrict__ x, int B, int D, int L) {
  pdl_sync();
  extern __shared__ float lat[];     // [4][L]
  const int b0 = blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * L; i += blockDim.x) {
    const int r = i / L, k = i - r * L, b = b0 + r;
    lat[i] = (b < B) ? (bos_flag[b] ? bos[k] : prev[b * L + k]) : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= D) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if constexpr (kL > 0) {
    float w[kL];
#pragma unroll
    for (int k = 0; k < kL; ++k) w[k] = __ldg(w_in_t + (long long)k * D + n);
#pragma unroll
    for (int k = 0; k < kL; ++k) {
      a0 = fmaf(w[k], lat[k], a0); a1 = fmaf(w[k], lat[kL + k], a1);
      a2 = fmaf(w[k], lat[2 * kL + k], a2); a3 = fmaf(w[k], lat[3 * kL + k], a3);
    }
  } else {
#pragma unroll 8
    for (int k = 0; k < L; ++k) {
      const float w = __ldg(w_in_t + (long long)k * D + n);
      a0 = fmaf(w, lat[k], a0); a1 = fmaf(w, lat[L + k], a1);
      a2 = fmaf(w, lat[2 * L + k], a2); a3 = fmaf(w, lat[3 * L + k], a3);
    }
  }
  if (b0 < B) x[(long long)b0 * D + n] = a0;
  if (b0 + 1 < B) x[(long long)(b0 + 1) * D + n] = a1;
  if (b0 + 2 < B) x[(long long)(b0 + 2) * D + n] = a2;
  if (b0 + 3 < B) x[(long long)(b0 + 3) * D + n] = a3;
}

template <typename WT>
__global__ void embed_rows_kernel(const WT* __restrict__ table, const int* __restrict__ ids,
                                  float* __restrict__ rows, int D) {
  pdl_sync();
  const int m = blockIdx.x;
  const WT* src = table + (long long)ids[m] * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) rows[(long long)m * D + c] = (float)src[c];
}

// Philox4x32-10 (Salmon et al. 2011), one 128-bit block per pair of outputs is plenty here.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// x0[i] = clamp(g_i * std): g from the host buffer z or from Philox(seed, frame counter, i)
__device__ __forceinline__ float noise_prep_one(const float* __restrict__ z, float* __restrict__ x0, int i, float std,
                                                float clamp, int use_philox,
                                                const unsigned long long* __restrict__ counter) {
  float g;
  if (use_philox) {
    const unsigned long long step = counter[0], seed = counter[1];   // {frame counter, seed} live on the device
    uint4 r = philox4x32(make_uint4((unsigned)i, (unsigned)step, (unsigned)(step >> 32), 0u),
                         make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float u1 = ((r.x >> 8) + 1u) * (1.0f / 16777216.0f);   // (0,1]
    const float u2 = (r.y >> 8) * (1.0f / 16777216.0f);
    g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  } else {
    g = z[i];
  }
  float v = g * std;
  if (clamp >= 0.f) v = fminf(fmaxf(v, -clamp), clamp);
  x0[i] = v;
  return v;
}

__global__ void __launch_bounds__(128) final_norm_eos_kernel(const float* __restrict__ x,
                                                             const int* __restrict__ row_of,
                                                             const float* __restrict__ ln_w,
                                                             const float* __restrict__ ln_b,
                                                             const float* __restrict__ w_eos,
                                                             const float* __restrict__ b_eos,
                                                             float* __restrict__ cout,
                                                             __nv_bfloat16* __restrict__ cout16,
                                                             float* __restrict__ logit, int D,
                                                             const float* __restrict__ acc, int acc_n,
                                                             long long acc_stride,
                                                             // flow-head start noise of the same row (x0 null: skip)
                                                             const float* __restrict__ nz, float* __restrict__ x0, int nL,
                                                             float nstd, float nclamp, int use_philox,
                                                             const unsigned long long* __restrict__ counter,
                                                             // bf16 copy of x0 as rows of 64 (zero beyond nL): A operand of
                                                             // the flow head's input projection in the chain kernel
                                                             __nv_bfloat16* __restrict__ x0_16) {
  pdl_sync();
  __shared__ float red[32];
  const int b = blockIdx.x;
  if (x0 && (int)threadIdx.x < nL) {
    const float v0 = noise_prep_one(nz, x0, b * nL + threadIdx.x, nstd, nclamp, use_philox, counter);
    if (x0_16) x0_16[b * 64 + threadIdx.x] = __float2bfloat16_rn(v0);
  }
  const long long row = row_of ? row_of[b] : b;
  const float* xr = x + row * D;
  if (D == 1024 && acc_n <= 8) {
    // vectorised path: two float4 per thread, every plane load issued before the first add (latency-bound kernel)
    const int tid = threadIdx.x;
    float4 u[2], q4[8][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) u[i] = reinterpret_cast<const float4*>(xr)[tid + 128 * i];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < acc_n) {
#pragma unroll
        for (int i = 0; i < 2; ++i) q4[k][i] = reinterpret_cast<const float4*>(acc + k * acc_stride + row * D)[tid + 128 * i];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < acc_n) {
#pragma unroll
        for (int i = 0; i < 2; ++i) { u[i].x += q4[k][i].x; u[i].y += q4[k][i].y; u[i].z += q4[k][i].z; u[i].w += q4[k][i].w; }
      }
    float s = (u[0].x + u[0].y) + (u[0].z + u[0].w) + (u[1].x + u[1].y) + (u[1].z + u[1].w);
    const float mean = block_sum(s, red) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      u[i].x -= mean; u[i].y -= mean; u[i].z -= mean; u[i].w -= mean;
      q += (u[i].x * u[i].x + u[i].y * u[i].y) + (u[i].z * u[i].z + u[i].w * u[i].w);
    }
    const float rstd = 1.0f / sqrtf(block_sum(q, red) / (float)D + 1e-5f);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c4 = tid + 128 * i;
      const float4 w = reinterpret_cast<const float4*>(ln_w)[c4], bb = reinterpret_cast<const float4*>(ln_b)[c4];
      const float4 we = reinterpret_cast<const float4*>(w_eos)[c4];
      const float4 o = make_float4(u[i].x * rstd * w.x + bb.x, u[i].y * rstd * w.y + bb.y, u[i].z * rstd * w.z + bb.z,
                                   u[i].w * rstd * w.w + bb.w);
      reinterpret_cast<float4*>(cout + (long long)b * D)[c4] = o;
      if (cout16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        reinterpret_cast<uint2*>(cout16 + (long long)b * D)[c4] =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
      dot = fmaf(o.x, we.x, dot); dot = fmaf(o.y, we.y, dot); dot = fmaf(o.z, we.z, dot); dot = fmaf(o.w, we.w, dot);
    }
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) logit[b] = dot + b_eos[0];
    return;
  }
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    v[i] = c < D ? xr[c] : 0.f;
    for (int k = 0; k < acc_n; ++k)
      if (c < D) v[i] += acc[k * acc_stride + row * D + c];
    s += v[i];
  }
  const float mean = block_sum(s, red) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    const float d = c < D ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(block_sum(q, red) / (float)D + 1e-5f);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    if (c < D) {
      const float o = (v[i] - mean) * rstd * ln_w[c] + ln_b[c];
      cout[(long long)b * D + c] = o;
      if (cout16) cout16[(long long)b * D + c] = __float2bfloat16_rn(o);
      dot = fmaf(o, w_eos[c], dot);
    }
  }
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) logit[b] = dot + b_eos[0];
}

__global__ void noise_prep_kernel(const float* __restrict__ z, float* __restrict__ x0, int n, float std,
                                  float clamp, int use_philox,
                                  const unsigned long long* __restrict__ counter) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  noise_prep_one(z, x0, i, std, clamp, use_philox, counter);
}

// wq_t [L][C] and wu_t [2S][C] are stored transposed (channel fastest) for coalesced reads.  One thread = one
// channel of one sequence; the L + 2S weight loads are all issued up front (latency-bound kernel), grid = (B, C/128).
template <int kL, int kS>
__global__ void __launch_bounds__(128) quant_upsample_kernel(const float* __restrict__ lat, const float* __restrict__ emb_std,
                                                             const float* __restrict__ emb_mean,
                                                             const float* __restrict__ wq_t, const float* __restrict__ wu_t,
                                                             float* __restrict__ zprev, float* __restrict__ out,
                                                             long long out_bs, int L, int C, int S) {
  pdl_sync();
  extern __shared__ float u[];   // normalised latent
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) u[i] = lat[b * L + i] * emb_std[i] + emb_mean[i];
  __syncthreads();
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if constexpr (kL > 0) {
    float wq[kL], wa[kS], wb[kS];
#pragma unroll
    for (int i = 0; i < kL; ++i) wq[i] = __ldg(wq_t + (long long)i * C + c);
#pragma unroll
    for (int t = 0; t < kS; ++t) {
      wa[t] = __ldg(wu_t + (long long)t * C + c);
      wb[t] = __ldg(wu_t + (long long)(kS + t) * C + c);
    }
    const float zp = zprev[(long long)b * C + c];
    float z = 0.f;
#pragma unroll
    for (int i = 0; i < kL; ++i) z = fmaf(wq[i], u[i], z);
    zprev[(long long)b * C + c] = z;
#pragma unroll
    for (int t = 0; t < kS; ++t) out[b * out_bs + (long long)t * C + c] = fmaf(wa[t], z, wb[t] * zp);
  } else {
    float z = 0.f;
#pragma unroll 8
    for (int i = 0; i < L; ++i) z = fmaf(__ldg(wq_t + (long long)i * C + c), u[i], z);
    const float zp = zprev[(long long)b * C + c];
    zprev[(long long)b * C + c] = z;
#pragma unroll 8
    for (int t = 0; t < S; ++t)
      out[b * out_bs + (long long)t * C + c] =
          fmaf(__ldg(wu_t + (long long)t * C + c), z, __ldg(wu_t + (long long)(S + t) * C + c) * zp);
  }
}

// 128 outputs per CTA; the (128 + taps - 1) x C input tile is staged with ELU applied, padded rows
template <typename XT, bool kElu>
__global__ void __launch_bounds__(128) final_conv_kernel(const XT* __restrict__ x, long long x_bs,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         float* __restrict__ audio, long long audio_bs,
                                                         int T, int C, int taps, short* __restrict__ pcm) {
  pdl_sync();
  extern __shared__ float sm[];
  float* ws = sm;                       // [taps*C]
  float* xs = sm + taps * C;            // [128+taps-1][C+1]
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  const int rows = min(128, T - t0) + taps - 1;
  for (int i = threadIdx.x; i < taps * C; i += 128) ws[i] = w[i];
  const XT* src = x + b * x_bs + (long long)t0 * C;
  if constexpr (sizeof(XT) == 2) {
    const uint4* src8 = reinterpret_cast<const uint4*>(src);     // rows are C*2 bytes, C % 8 == 0
    for (int i = threadIdx.x; i < rows * C / 8; i += 128) {
      const uint4 q = src8[i];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
      const int r = (i * 8) / C, c = (i * 8) - r * C;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        xs[r * (C + 1) + c + 2 * j] = kElu ? act_apply(f.x, ACT_ELU) : f.x;
        xs[r * (C + 1) + c + 2 * j + 1] = kElu ? act_apply(f.y, ACT_ELU) : f.y;
      }
    }
  } else {
    for (int i = threadIdx.x; i < rows * C; i += 128) {
      const int r = i / C, c = i - r * C;
      const float v = (float)src[i];
      xs[r * (C + 1) + c] = kElu ? act_apply(v, ACT_ELU) : v;
    }
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float a = bias[0];
  for (int j = 0; j < taps; ++j) {
    const float* xr = xs + (threadIdx.x + j) * (C + 1);
    const float* wr = ws + j * C;
#pragma unroll 16
    for (int c = 0; c < C; ++c) a = fmaf(xr[c], wr[c], a);
  }
  audio[b * audio_bs + t] = a;
  if (pcm) pcm[b * audio_bs + t] = pcm16_of(a);
}

// End-of-frame bookkeeping of the Mimi decoder in one launch: blockIdx.y < n_entries moves the last taps-1 rows of a
// streaming conv input to the front (its carried state); blockIdx.y == n_entries finishes the fused SEANet tail for
// sequence blockIdx.x (first two samples of every 128-step tile + carry of the boundary partials, see seanet_tail.cu)
// and advances the sequence's ring offset.
__global__ void state_shift_kernel(const ShiftEntry* __restrict__ entries, int n_entries, float* __restrict__ audio,
                                   long long audio_bs, float* __restrict__ bnd, int tiles_t,
                                   int* __restrict__ mimi_offset, int inc_mimi, short* __restrict__ pcm) {
  pdl_sync();
  const int b = blockIdx.x;
  if ((int)blockIdx.y == n_entries) {
    if (bnd) {
      for (int k = threadIdx.x; k < tiles_t; k += blockDim.x) {
        const float* slot = bnd + ((long long)b * (tiles_t + 1) + k) * 4;
        float* a = audio + b * audio_bs + (long long)k * 128;
        const float a0 = a[0] + slot[0] + slot[2], a1 = a[1] + slot[1];
        a[0] = a0;
        a[1] = a1;
        if (pcm) {                                  // the tail kernel converted every other sample of the tile
          short* q = pcm + b * audio_bs + (long long)k * 128;
          q[0] = pcm16_of(a0);
          q[1] = pcm16_of(a1);
        }
      }
      __syncthreads();
      if (threadIdx.x < 3) {
        float* first = bnd + (long long)b * (tiles_t + 1) * 4;
        first[threadIdx.x] = first[(long long)tiles_t * 4 + threadIdx.x];
      }
    }
    if (mimi_offset && threadIdx.x == 0) mimi_offset[b] += inc_mimi;
    return;
  }
  const ShiftEntry e = entries[blockIdx.y];
  // rows*C*esz is a multiple of 4 bytes for every buffer of the decoder (C >= 64)
  unsigned* base = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(e.buf) + b * e.bs * e.esz);
  const int n = e.rows * e.C * e.esz / 4;
  const unsigned* src = base + (long long)e.T * e.C * e.esz / 4;
  for (int i = threadIdx.x; i < n; i += blockDim.x) base[i] = src[i];
}

__global__ void advance_kernel(int* seq_len, int* bos_flag, int* mimi_offset, unsigned long long* counter,
                               int B, int inc_len, int inc_mimi, const int* active) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const bool on = !active || active[i];         // parked slots of a continuous batch do not grow their KV
    if (seq_len && on) seq_len[i] += inc_len;
    if (bos_flag && inc_len && on) bos_flag[i] = 0;
    if (mimi_offset) mimi_offset[i] += inc_mimi;
  }
  if (i == 0 && counter) *counter += 1ull;
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float a, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

template <typename KT>
__global__ void copy_pages_kernel(KT* pool, long long layer_stride, long long page_stride,
                                  const int* __restrict__ src_pages, const int* __restrict__ dst_pages) {
  pdl_sync();
  const int pair = blockIdx.x, layer = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(pool + layer * layer_stride + src_pages[pair] * page_stride);
  uint4* d = reinterpret_cast<uint4*>(pool + layer * layer_stride + dst_pages[pair] * page_stride);
  const int n = (int)(page_stride * sizeof(KT) / 16);
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

__global__ void fill_u32_kernel(unsigned int* dst, unsigned int v, long long n) {
  pdl_sync();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

__global__ void gather_frame_kernel(const float* __restrict__ all, float* __restrict__ lat, int F, int L,
                                    const int* __restrict__ frame_idx) {
  pdl_sync();
  const int b = blockIdx.x, f = *frame_idx;
  for (int i = threadIdx.x; i < L; i += blockDim.x) lat[b * L + i] = all[((long long)b * F + f) * L + i];
}

__global__ void scatter_audio_kernel(const float* __restrict__ audio, float* __restrict__ all, int F, int n,
                                     const int* __restrict__ frame_idx) {
  pdl_sync();
  const int b = blockIdx.x, f = *frame_idx;
  for (int i = threadIdx.x; i < n; i += blockDim.x) all[((long long)b * F + f) * n + i] = audio[(long long)b * n + i];
}

__global__ void inc_kernel(int* v, int inc) {
  pdl_sync(); *v += inc; }

}  // namespace

// Few rows (decode at batch <= 512): four warps per row, C = 512 * VEC.  Every load of the row and of its split-K
// planes is issued before the first use, the two reductions go through shared memory; same arithmetic order per
// element as layernorm_kernel (planes added in index order), so the two kernels agree to rounding of the sums.
template <int VEC>
__global__ void __launch_bounds__(128) layernorm_wide_kernel(const NormParams p) {
  pdl_sync();
  __shared__ float red[2][4];
  const int m = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = m / p.T, t = m % p.T;
  float* xrow = const_cast<float*>(p.X) + b * p.x_bs + t * p.x_rs;
  const float4* x = reinterpret_cast<const float4*>(xrow);
  float4 v[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = x[tid + 128 * i];
  if (p.acc_n > 0) {
    float4 q[8][VEC];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < p.acc_n) {
        const float4* a = reinterpret_cast<const float4*>(p.acc + k * p.acc_stride + (long long)m * p.C);
#pragma unroll
        for (int i = 0; i < VEC; ++i) q[k][i] = a[tid + 128 * i];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < p.acc_n) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) { v[i].x += q[k][i].x; v[i].y += q[k][i].y; v[i].z += q[k][i].z; v[i].w += q[k][i].w; }
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) reinterpret_cast<float4*>(xrow)[tid + 128 * i] = v[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  s = warp_sum(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  const float mean = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) / (float)p.C;
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q2 += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  q2 = warp_sum(q2);
  if (lane == 0) red[1][warp] = q2;
  __syncthreads();
  const float rstd = 1.0f / sqrtf(((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])) / (float)p.C + p.eps);
  const float4* sc = p.scale ? reinterpret_cast<const float4*>(p.scale + (long long)m * p.mod_rs) : nullptr;
  const float4* sh = p.shift ? reinterpret_cast<const float4*>(p.shift + (long long)m * p.mod_rs) : nullptr;
  const float4* w4 = reinterpret_cast<const float4*>(p.w);
  const float4* b4 = reinterpret_cast<const float4*>(p.b);
  const long long yo = b * p.y_bs + t * p.y_rs;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c4 = tid + 128 * i;
    float4 o = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
    if (p.w) {
      const float4 w = w4[c4], bb = b4[c4];
      o.x = o.x * w.x + bb.x; o.y = o.y * w.y + bb.y; o.z = o.z * w.z + bb.z; o.w = o.w * w.w + bb.w;
    }
    if (sc) {
      const float4 a = sc[c4], d = sh[c4];
      o.x = o.x * (1.0f + a.x) + d.x; o.y = o.y * (1.0f + a.y) + d.y;
      o.z = o.z * (1.0f + a.z) + d.z; o.w = o.w * (1.0f + a.w) + d.w;
    }
    if (p.Y16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(p.Y16 + yo + 4 * c4) = pk;
    } else {
      *reinterpret_cast<float4*>(p.Y + yo + 4 * c4) = o;
    }
  }
}

// ---- Mimi encoder helpers (voice cloning, one-off per voice) --------------------------------------------------
// First SEANet-encoder conv: 1 input channel, k taps, causal (k-1 zeros in front of x): y[t][n] = b[n] + sum_j w[n][j] x~[t+j]
__global__ void enc_conv0_kernel(const float* __restrict__ xpad, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ y, long long T, int N, int k) {
  extern __shared__ float sw[];          // [N][k] | [N]
  for (int i = threadIdx.x; i < N * k; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) sw[N * k + i] = b[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * N) return;
  const long long t = idx / N;
  const int n = (int)(idx - t * N);
  float a = sw[N * k + n];
  for (int j = 0; j < k; ++j) a = fmaf(sw[n * k + j], xpad[t + j], a);
  y[idx] = a;
}

void launch_enc_conv0(const float* xpad, const float* w, const float* b, float* y, long long T, int N, int k, cudaStream_t s) {
  const long long total = T * N;
  launch_k(enc_conv0_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(N * k + N) * sizeof(float), s, xpad, w, b,
           y, T, N, k);
  ++g_launches;
}

// dst rows 0..n-1 <- src row (replicate padding of the downsample conv)
__global__ void replicate_row_kernel(float* __restrict__ dst, const float* __restrict__ src, int n, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * C) dst[i] = src[i % C];
}

void launch_replicate_row(float* dst, const float* src, int n, int C, cudaStream_t s) {
  launch_k(replicate_row_kernel, dim3((n * C + 255) / 256), dim3(256), 0, s, dst, src, n, C);
  ++g_launches;
}

void launch_layernorm(const NormParams& p, cudaStream_t s) {
  ProfScope ps("layernorm", nullptr, 0, 2.0 * p.nb * p.T * p.C * 4, s);
  const int rows = p.nb * p.T;
  if (rows <= 512 && p.acc_n <= 8 && (p.C == 512 || p.C == 1024)) {
    if (p.C == 1024) launch_k(layernorm_wide_kernel<2>, dim3(rows), dim3(128), 0, s, p);
    else launch_k(layernorm_wide_kernel<1>, dim3(rows), dim3(128), 0, s, p);
    ++g_launches;
    return;
  }
  const int rpc = rows >= 2048 ? 8 : (rows >= 512 ? 4 : 1);     // rows per CTA: keep >= ~256 CTAs in flight
  const int grid = (rows + rpc - 1) / rpc, block = 32 * rpc;
  switch (p.C / 128) {
    case 8: launch_k(layernorm_kernel<8>, dim3(grid), dim3(block), 0, s, p); break;
    case 4: launch_k(layernorm_kernel<4>, dim3(grid), dim3(block), 0, s, p); break;
    case 2: launch_k(layernorm_kernel<2>, dim3(grid), dim3(block), 0, s, p); break;
    default: launch_k(layernorm_kernel<1>, dim3(grid), dim3(block), 0, s, p); break;
  }
  ++g_launches;
}

void launch_input_rows(const float* w_in, const float* bos, const float* prev, const int* bos_flag, float* x,
                       int B, int D, int L, cudaStream_t s) {
  ProfScope ps("input_rows", nullptr, 0, (double)B * (D + L) * 4 + (double)D * L * 4, s);
  const dim3 grid((B + 3) / 4, (D + 255) / 256);
  if (L == 32) launch_k(input_rows_kernel<32>, grid, dim3(256), 4 * L * sizeof(float), s, w_in, bos, prev, bos_flag, x, B, D, L);
  else launch_k(input_rows_kernel<0>, grid, dim3(256), 4 * L * sizeof(float), s, w_in, bos, prev, bos_flag, x, B, D, L);
  ++g_launches;
}

void launch_embed_rows(const void* table, int table_bf16, const int* ids, float* rows, int M, int D,
                       cudaStream_t s) {
  ProfScope ps("embed_rows", nullptr, 0, (double)M * D * 6, s);
  if (table_bf16)
    launch_k(embed_rows_kernel<__nv_bfloat16>, dim3(M), dim3(256), 0, s, (const __nv_bfloat16*)table, ids, rows, D);
  else
    launch_k(embed_rows_kernel<float>, dim3(M), dim3(256), 0, s, (const float*)table, ids, rows, D);
  ++g_launches;
}

void launch_final_norm_eos(const float* x, const int* row_of, const float* ln_w, const float* ln_b,
                           const float* w_eos, const float* b_eos, float* c, __nv_bfloat16* c16, float* logit,
                           int B, int D, const float* acc, int acc_n, long long acc_stride, cudaStream_t s,
                           const float* nz, float* x0, int nL, float nstd, float nclamp, int use_philox,
                           const unsigned long long* counter, __nv_bfloat16* x0_16) {
  ProfScope ps("final_norm_eos", nullptr, 0, 2.0 * B * D * 4, s);
  if (nL > 128) { x0 = nullptr; }
  if (nL > 64) x0_16 = nullptr;
  launch_k(final_norm_eos_kernel, dim3(B), dim3(128), 0, s, x, row_of, ln_w, ln_b, w_eos, b_eos, c, c16, logit, D, acc, acc_n, acc_stride,
           nz, x0, nL, nstd, nclamp, use_philox, counter, x0_16);
  ++g_launches;
}

void launch_noise_prep(const float* z, float* x0, int n, float std, float clamp, int use_philox,
                       const unsigned long long* counter, cudaStream_t s) {
  ProfScope ps("noise_prep", nullptr, 0, 2.0 * n * 4, s);
  launch_k(noise_prep_kernel, dim3((n + 255) / 256), dim3(256), 0, s, z, x0, n, std, clamp, use_philox, counter);
  ++g_launches;
}

void launch_quant_upsample(const float* lat, const float* emb_std, const float* emb_mean, const float* wq,
                           const float* wu, float* zprev, float* out, long long out_bs, int B, int L, int C,
                           int S, cudaStream_t s) {
  ProfScope ps("quant_upsample", nullptr, 0, (double)B * S * C * 4, s);
  const dim3 grid(B, (C + 127) / 128);
  if (L == 32 && S == 16)
    launch_k(quant_upsample_kernel<32, 16>, grid, dim3(128), L * sizeof(float), s, lat, emb_std, emb_mean, wq, wu, zprev, out,
             out_bs, L, C, S);
  else
    launch_k(quant_upsample_kernel<0, 0>, grid, dim3(128), L * sizeof(float), s, lat, emb_std, emb_mean, wq, wu, zprev, out,
             out_bs, L, C, S);
  ++g_launches;
}

void launch_final_conv(const float* x, long long x_bs, const float* w, const float* bias, float* audio,
                       long long audio_bs, int B, int T, int C, int taps, cudaStream_t s, short* pcm) {
  ProfScope ps("final_conv", nullptr, 0, (double)B * T * (C + 1) * 4, s);
  const size_t smem = (size_t)(taps * C + (128 + taps - 1) * (C + 1)) * sizeof(float);
  dim3 grid((T + 127) / 128, B);
  launch_k(final_conv_kernel<float, true>, dim3(grid), dim3(128), smem, s, x, x_bs, w, bias, audio, audio_bs, T, C, taps, pcm);
  ++g_launches;
}

void launch_final_conv16(const __nv_bfloat16* x, long long x_bs, const float* w, const float* bias, float* audio,
                         long long audio_bs, int B, int T, int C, int taps, cudaStream_t s, short* pcm) {
  ProfScope ps("final_conv", nullptr, 0, (double)B * T * (C / 2 + 1) * 4, s);
  const size_t smem = (size_t)(taps * C + (128 + taps - 1) * (C + 1)) * sizeof(float);
  dim3 grid((T + 127) / 128, B);
  launch_k(final_conv_kernel<__nv_bfloat16, false>, dim3(grid), dim3(128), smem, s, x, x_bs, w, bias, audio, audio_bs, T, C, taps, pcm);
  ++g_launches;
}

void launch_state_shift(const ShiftEntry* entries_dev, int n_entries, int B, cudaStream_t s, float* audio,
                        long long audio_bs, float* bnd, int tiles_t, int* mimi_offset, int inc_mimi, short* pcm) {
  ProfScope ps("state_shift", nullptr, 0, 0, s);
  const bool extra = bnd || mimi_offset;
  dim3 grid(B, n_entries + (extra ? 1 : 0));
  launch_k(state_shift_kernel, dim3(grid), dim3(256), 0, s, entries_dev, n_entries, audio, audio_bs, bnd, tiles_t, mimi_offset,
           inc_mimi, pcm);
  ++g_launches;
}

// grid (pieces, slots, 8 chunks of a piece)
__global__ void __launch_bounds__(256) restore_state_kernel(const StatePieceDev* __restrict__ pieces, const int* __restrict__ slots,
                                                            const char* __restrict__ tpl) {
  pdl_sync();
  const StatePieceDev p = pieces[blockIdx.x];
  char* dst = p.base + (unsigned long long)slots[blockIdx.y] * p.stride;
  const char* src = tpl ? tpl + p.tpl_off : nullptr;
  const bool wide = (p.bytes & 15) == 0 && (reinterpret_cast<unsigned long long>(dst) & 15) == 0 &&
                    (!src || (reinterpret_cast<unsigned long long>(src) & 15) == 0);
  if (wide) {
    const long long n = (long long)(p.bytes >> 4);
    const long long per = (n + gridDim.z - 1) / gridDim.z, lo = per * blockIdx.z, hi = min(n, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x)
      reinterpret_cast<uint4*>(dst)[i] = src ? __ldg(reinterpret_cast<const uint4*>(src) + i) : make_uint4(0, 0, 0, 0);
  } else if (blockIdx.z == 0) {
    const long long n = (long long)(p.bytes >> 2);
    for (long long i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<unsigned*>(dst)[i] = src ? __ldg(reinterpret_cast<const unsigned*>(src) + i) : 0u;
  }
}

void launch_restore_state(const StatePieceDev* pieces, int n_pieces, const int* slots, int n_slots, const char* tpl, cudaStream_t s) {
  if (n_pieces <= 0 || n_slots <= 0) return;
  ProfScope ps("restore_state", nullptr, 0, 0, s);
  launch_k(restore_state_kernel, dim3(n_pieces, n_slots, 8), dim3(256), 0, s, pieces, slots, tpl);
  ++g_launches;
}

void launch_advance(int* seq_len, int* bos_flag, int* mimi_offset, unsigned long long* counter, int B,
                    int inc_len, int inc_mimi, cudaStream_t s, const int* active) {
  ProfScope ps("advance", nullptr, 0, 0, s);
  launch_k(advance_kernel, dim3((B + 255) / 256), dim3(256), 0, s, seq_len, bos_flag, mimi_offset, counter, B, inc_len, inc_mimi,
           active);
  ++g_launches;
}

void launch_axpy(const float* x, float* y, float a, int n, cudaStream_t s) {
  ProfScope ps("axpy", nullptr, 0, 3.0 * n * 4, s);
  launch_k(axpy_kernel, dim3((n + 255) / 256), dim3(256), 0, s, x, y, a, n);
  ++g_launches;
}

void launch_copy_pages(void* pool, int kv_bf16, long long layer_stride, long long page_stride, int n_layers,
                       const int* src_pages, const int* dst_pages, int n_pairs, cudaStream_t s) {
  if (n_pairs <= 0) return;
  ProfScope ps("copy_pages", nullptr, 0, 2.0 * n_pairs * n_layers * page_stride * (kv_bf16 ? 2 : 4), s);
  dim3 grid(n_pairs, n_layers);
  if (kv_bf16)
    launch_k(copy_pages_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, (__nv_bfloat16*)pool, layer_stride, page_stride,
                                                          src_pages, dst_pages);
  else
    launch_k(copy_pages_kernel<float>, dim3(grid), dim3(256), 0, s, (float*)pool, layer_stride, page_stride, src_pages, dst_pages);
  ++g_launches;
}

void launch_fill_u32(unsigned int* dst, unsigned int v, long long n, cudaStream_t s) {
  ProfScope ps("fill_u32", nullptr, 0, (double)n * 4, s);
  launch_k(fill_u32_kernel, dim3(1184), dim3(256), 0, s, dst, v, n);
  ++g_launches;
}

void launch_gather_frame(const float* lat_all, float* lat, int B, int F, int L, const int* frame_idx,
                         cudaStream_t s) {
  ProfScope ps("gather_frame", nullptr, 0, 0, s);
  launch_k(gather_frame_kernel, dim3(B), dim3(32), 0, s, lat_all, lat, F, L, frame_idx);
  ++g_launches;
}

void launch_scatter_audio(const float* audio, float* audio_all, int B, int F, int n, const int* frame_idx,
                          cudaStream_t s) {
  ProfScope ps("scatter_audio", nullptr, 0, 2.0 * B * n * 4, s);
  launch_k(scatter_audio_kernel, dim3(B), dim3(256), 0, s, audio, audio_all, F, n, frame_idx);
  ++g_launches;
}

void launch_inc(int* v, int inc, cudaStream_t s) {
  ProfScope ps("inc", nullptr, 0, 0, s);
  launch_k(inc_kernel, dim3(1), dim3(1), 0, s, v, inc);
  ++g_launches;
}

}  // namespace ptts

// ---- eager per-kernel profiler -------------------------------------------------------------------------
#include <algorithm>
#include <map>
#include <string>
#include <vector>
namespace ptts {
bool g_prof_on = false;
namespace {
struct Rec { std::string name; cudaEvent_t e0, e1; double flops, bytes; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::string g_report;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* kernel, const char* tag, double flops, double bytes, cudaStream_t stream) : s(stream) {
  static const bool nvtx_on = [] { const char* v = getenv("PTTS_NVTX"); return v && v[0] == '1'; }();
  if (nvtx_on) {
    char name[96];
    snprintf(name, sizeof name, tag ? "%s:%s" : "%s", kernel, tag);
    nvtxRangePushA(name);
    nvtx = true;
  }
  if (!g_prof_on) return;
  Rec r;
  r.name = kernel;
  if (tag) { r.name += ":"; r.name += tag; }
  r.e0 = get_event(); r.e1 = get_event(); r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, s);
  slot = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_recs[slot].e1, s);
  if (nvtx) nvtxRangePop();
}
void prof_start() { g_recs.clear(); g_prof_on = true; }
const char* prof_report() {
  g_prof_on = false;
  struct Agg { double ms = 0, flops = 0, bytes = 0; long n = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : g_recs) {
    cudaEventSynchronize(r.e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& a = agg[r.name];
    a.ms += ms; a.flops += r.flops; a.bytes += r.bytes; a.n += 1;
    g_pool.push_back(r.e0); g_pool.push_back(r.e1);
  }
  g_recs.clear();
  std::vector<std::pair<std::string, Agg>> v(agg.begin(), agg.end());
  std::sort(v.begin(), v.end(), [](auto& x, auto& y) { return x.second.ms > y.second.ms; });
  g_report.clear();
  char line[512];
  for (auto& kv : v) {
    snprintf(line, sizeof line, "%s,%ld,%.6f,%.6g,%.6g\n", kv.first.c_str(), kv.second.n, kv.second.ms,
             kv.second.flops, kv.second.bytes);
    g_report += line;
  }
  return g_report.c_str();
}
}  // namespace ptts
