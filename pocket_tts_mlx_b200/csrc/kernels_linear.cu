// Multi-tap linear operator, CUDA-core paths.
//   * linear_tile_kernel : 64x64x16 register-tiled SIMT GEMM for any M (fp32 mode, odd shapes, and the
//                          reference point the tcgen05 path is checked against).
//   * linear_gemv_kernel : M <= 16 rows (batch-1 decode, 16-step Mimi chunk): every weight byte is read
//                          exactly once with 16-byte loads; HBM-bound by construction.
// Replaces mx.matmul / nn.Linear / nn.Conv1d / mx.conv_transpose1d call sites listed in SURVEY.md 2.3.
#include "kernels.cuh"

namespace ptts {

std::atomic<long long> g_launches{0};
std::atomic<bool> g_pdl_on{false};
std::atomic<bool> g_launch_prio_on{false};
thread_local int g_launch_prio = 0;

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename WT>
__device__ __forceinline__ void load_w4(const WT* w, float (&out)[4]);
template <>
__device__ __forceinline__ void load_w4<float>(const float* w, float (&out)[4]) {
  float4 v = *reinterpret_cast<const float4*>(w);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_w4<__nv_bfloat16>(const __nv_bfloat16* w, float (&out)[4]) {
  uint2 v = *reinterpret_cast<const uint2*>(w);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  out[0] = fa.x; out[1] = fa.y; out[2] = fb.x; out[3] = fb.y;
}

// grid (ceil(N/64), ceil(M/64)); 256 threads, each a 4x4 micro-tile.  K = taps*C, C % 16 == 0.
template <typename WT>
__global__ void __launch_bounds__(256) linear_tile_kernel(const LinearParams p) {
  pdl_sync();
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = p.nb * p.T, K = p.taps * p.C;
  const WT* W = reinterpret_cast<const WT*>(p.W);

  // loader mapping: 64 rows x 4 k-quads
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lr;
  const bool a_ok = am < M;
  const int ab = a_ok ? am / p.T : 0, at = a_ok ? am % p.T : 0;
  const float* a_row = p.A + ab * p.a_bs + at * p.a_rs;
  const int wn = n0 + lr;
  const bool w_ok = wn < p.N;
  const WT* w_row = W + (long long)(w_ok ? wn : 0) * K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int kk = k0 + lk;
    const int tap = kk / p.C, c = kk - tap * p.C;
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok) {
      float4 v = *reinterpret_cast<const float4*>(a_row + tap * p.a_rs + c);
      av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
      if (p.a_pro != ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = act_apply(av[i], p.a_pro);
      }
    }
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (w_ok) load_w4<WT>(w_row + kk, wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Ws[lk + i][lr] = wv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = m / p.T, t = m % p.T;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) epilogue_store(p, b, t, n, acc[i][j]);
    }
  }
}

// ---- small-M weight-streaming path -------------------------------------------------------------------
// grid = ceil(N / 8); 128 threads = 4 warps x 2 output rows each.  The (T+taps-1) x C input rows of every
// sequence are staged in shared memory once (prologue activation applied); each lane then streams
// 16-byte weight vectors and keeps MT accumulators per output row.
template <typename WT> struct WVec;
template <> struct WVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
};
template <> struct WVec<float> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const float* p, float (&o)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
};

// grid = ceil(N / (8*R)); 256 threads = 8 warps x R output rows each.  The (T+taps-1) x C input rows of every
// sequence are staged in shared memory once (prologue activation applied).  A warp walks its R weight rows in
// steps of U x 256 elements and issues all R*U 16-byte loads of a step before consuming any of them, so each lane
// keeps R*U*16 bytes in flight (R*U = 8: 32 KB per CTA) -- the kernel is meant to sit on the HBM roofline.
template <typename WT> struct WRaw;
template <> struct WRaw<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p, bool stream = false) {
    v = stream ? __ldcs(reinterpret_cast<const uint4*>(p)) : __ldg(reinterpret_cast<const uint4*>(p));
  }
  __device__ __forceinline__ void unpack(float (&o)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
};
template <> struct WRaw<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p, bool = false) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void unpack(float (&o)[8]) const {
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
};

// grid = ceil(N / (8*R)); 256 threads = 8 warps x R output rows each.  A warp walks its R weight rows in steps of
// U x 256 elements and issues all R*U 16-byte loads of a step before consuming any of them (R*U = 8: 32 KB in
// flight per CTA); the first step's loads are issued BEFORE the input rows are staged (and normalised) in shared
// memory, so the activation round trip hides behind the weight stream -- the kernel is meant to sit on the HBM
// roofline, not on two serialised DRAM latencies.
template <typename WT, int MT, int R>
__global__ void __launch_bounds__(256) linear_gemv_kernel(const LinearParams p) {
  pdl_sync();
  extern __shared__ __align__(16) float xs[];   // [nb][T+taps-1][C]
  constexpr int U = 8 / R;
  const int rows_per_seq = p.T + p.taps - 1;
  const int C = p.C, K = p.taps * p.C;
  const int M = p.nb * p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = (blockIdx.x * 8 + warp) * R;
  const WT* W = reinterpret_cast<const WT*>(p.W);
  const WT* wrow[R];
#pragma unroll
  for (int r = 0; r < R; ++r) wrow[r] = W + (long long)min(n_base + r, p.N - 1) * K;

  WRaw<WT> wq[R][U];
  const bool w_stream = p.w_stream != 0;
  auto issue = [&](int k0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + 256 * u;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (k < K) wq[r][u].load(wrow[r] + k, w_stream);
    }
  };
  issue(lane * 8);                                  // weights do not depend on the activations: fetch first

  {
    const int c4 = C >> 2;
    const int total = p.nb * rows_per_seq * c4;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int q = i % c4, r = (i / c4) % rows_per_seq, b = i / (c4 * rows_per_seq);
      float4 v = *reinterpret_cast<const float4*>(p.A + b * p.a_bs + r * p.a_rs + q * 4);
      if (p.a_pro != ACT_NONE) {
        v.x = act_apply(v.x, p.a_pro); v.y = act_apply(v.y, p.a_pro);
        v.z = act_apply(v.z, p.a_pro); v.w = act_apply(v.w, p.a_pro);
      }
      *reinterpret_cast<float4*>(xs + ((long long)(b * rows_per_seq + r)) * C + q * 4) = v;
    }
  }
  __syncthreads();
  if (p.ln_on) {
    // fused LayerNorm (taps == 1): warp w normalises rows w, w+8, ... in place, two-pass like layernorm_kernel
    for (int m = warp; m < M; m += 8) {
      float* xr = xs + (long long)m * C;
      float sum = 0.f;
      for (int c = lane; c < C; c += 32) sum += xr[c];
      const float mean = warp_sum(sum) / (float)C;
      float q = 0.f;
      for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + p.ln_eps);
      const float* sc = p.ln_scale ? p.ln_scale + (long long)m * p.ln_mod_rs : nullptr;
      const float* sh = p.ln_shift ? p.ln_shift + (long long)m * p.ln_mod_rs : nullptr;
      for (int c = lane; c < C; c += 32) {
        float o = (xr[c] - mean) * rstd;
        if (p.ln_w) o = o * p.ln_w[c] + p.ln_b[c];
        if (sc) o = o * (1.0f + sc[c]) + sh[c];
        xr[c] = o;
      }
    }
    __syncthreads();
  }

  float acc[R][MT];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

  for (int k0 = lane * 8; k0 < K; k0 += 256 * U) {
    float w[R][U][8];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r) wq[r][u].unpack(w[r][u]);
    if (k0 + 256 * U < K) issue(k0 + 256 * U);      // next step's loads fly while this step is multiplied
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + 256 * u;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          if (m < M) {
            const int b = m / p.T, t = m - b * p.T;
            const float* xr = xs + ((long long)(b * rows_per_seq + t + tap)) * C + c;
            const float4 x0 = *reinterpret_cast<const float4*>(xr);
            const float4 x1 = *reinterpret_cast<const float4*>(xr + 4);
            const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[r][m] = fmaf(w[r][u][i], x[i], acc[r][m]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[r][m] = warp_sum(acc[r][m]);

  if constexpr (R == 2) {
    if (p.rope_on) {
      // fused RoPE + KV append: this warp owns the interleaved pair (n_base, n_base + 1) of every row
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (lane == (m & 31) && m < M && n_base + 1 < p.N) {
          const int Dm = p.kv_heads * kHeadDim;
          const int which = n_base / Dm, within = n_base - which * Dm;      // 0 q, 1 k, 2 v
          const int pos = p.kv_row_pos[m];
          float x = acc[0][m], y = acc[1][m];
          if (which < 2) {
            float sn, cs;
            sincosf((float)pos * p.rope_freqs[(within & 63) >> 1], &sn, &cs);
            const float xr = x * cs - y * sn, yr = x * sn + y * cs;
            x = xr; y = yr;
          }
          if (which == 0) {
            p.q_rot[(long long)m * Dm + within] = x;
            p.q_rot[(long long)m * Dm + within + 1] = y;
          } else {
            const int seq = p.kv_row_seq ? p.kv_row_seq[m] : m;
            const int page = p.kv_page_table[(long long)seq * p.kv_max_pages + pos / kPageTokens];
            const long long off = page * p.kv_page_stride + ((long long)(within >> 6) * kPageTokens + pos % kPageTokens) * kHeadDim +
                                  (within & 63) + (which == 2 ? (long long)p.kv_heads * kPageTokens * kHeadDim : 0);
            if (p.kv_bf16) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.kv_layer) + off) = __floats2bfloat162_rn(x, y);
            else *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.kv_layer) + off) = make_float2(x, y);
          }
        }
      }
      return;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (lane == ((r * MT + m) & 31) && m < M) {
        const int n = n_base + r;
        if (n < p.N) {
          const int b = m / p.T, t = m - b * p.T;
          epilogue_store(p, b, t, n, acc[r][m]);
        }
      }
    }
  }
}

template <typename WT, int R>
void launch_gemv_r(const LinearParams& p, cudaStream_t s) {
  const int M = p.nb * p.T;
  const size_t smem = (size_t)p.nb * (p.T + p.taps - 1) * p.C * sizeof(float);
  dim3 grid((p.N + 8 * R - 1) / (8 * R)), block(256);
#define PTTS_GEMV(MT)                                                                                   \
  do {                                                                                                  \
    auto kfn = linear_gemv_kernel<WT, MT, R>;                                                           \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    launch_k(kfn, grid, block, smem, s, p);                                                             \
  } while (0)
  if (M <= 1) PTTS_GEMV(1);
  else if (M <= 2) PTTS_GEMV(2);
  else if (M <= 4) PTTS_GEMV(4);
  else if (M <= 8) PTTS_GEMV(8);
  else PTTS_GEMV(16);
#undef PTTS_GEMV
}

template <typename WT>
void launch_gemv_t(const LinearParams& p, cudaStream_t s) {
  // rows per warp: enough CTAs to cover the chip first (N / (8R) >= ~150), then loads in flight per lane
  const int M = p.nb * p.T;
  if (M > 4 || p.N <= 1536) launch_gemv_r<WT, 1>(p, s);
  else launch_gemv_r<WT, 2>(p, s);
}

}  // namespace

bool linear_gemv_rope_supported(const LinearParams& p) {
  const int M = p.nb * p.T;
  return linear_gemv_supported(p) && M <= 4 && p.N > 1536 && (p.N % 2) == 0;     // the R = 2 instantiation
}

bool linear_gemv_supported(const LinearParams& p) {
  const int M = p.nb * p.T;
  const size_t smem = (size_t)p.nb * (p.T + p.taps - 1) * p.C * sizeof(float);
  return M <= 16 && (p.C % 8) == 0 && smem <= 160 * 1024;
}

// Tiny-K Linear over many rows (flow-head input_proj: K = 32): one thread = one output feature of four rows, its
// whole bf16 weight row (K * 2 bytes) fetched with 16-byte loads before the first FMA.  y = x W^T + b, fp32 in/out.
template <int K>
__global__ void __launch_bounds__(256) small_k_linear_kernel(const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias,
                                                             const float* __restrict__ X, float* __restrict__ Y, int M, int N) {
  pdl_sync();
  __shared__ float xs[4][K];
  const int m0 = blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * K; i += 256) {
    const int r = i / K, k = i - r * K;
    xs[r][k] = (m0 + r < M) ? X[(long long)(m0 + r) * K + k] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y * 256 + threadIdx.x;
  if (n >= N) return;
  uint4 wq[K / 8];
#pragma unroll
  for (int i = 0; i < K / 8; ++i) wq[i] = __ldg(reinterpret_cast<const uint4*>(W + (long long)n * K) + i);
  const float bb = bias ? bias[n] : 0.f;
  float a[4] = {bb, bb, bb, bb};
#pragma unroll
  for (int i = 0; i < K / 8; ++i) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&wq[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 w = __bfloat1622float2(h[j]);
      const int k = 8 * i + 2 * j;
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = fmaf(w.y, xs[r][k + 1], fmaf(w.x, xs[r][k], a[r]));
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (m0 + r < M) Y[(long long)(m0 + r) * N + n] = a[r];
}

bool launch_small_k_linear(const __nv_bfloat16* W, const float* bias, const float* X, float* Y, int M, int K, int N,
                           const char* tag, cudaStream_t s) {
  if (K != 32) return false;
  ProfScope ps("small_k_linear", tag, 2.0 * M * N * K, (double)N * K * 2 + (double)M * (K + N) * 4, s);
  launch_k(small_k_linear_kernel<32>, dim3((M + 3) / 4, (N + 255) / 256), dim3(256), 0, s, W, bias, X, Y, M, N);
  ++g_launches;
  return true;
}

void launch_linear_tile(const LinearParams& p, cudaStream_t s) {
  const int M = p.nb * p.T;
  dim3 grid((p.N + BN - 1) / BN, (M + BM - 1) / BM), block(256);
  ProfScope ps("linear_tile", p.tag, linear_flops(p), linear_bytes(p), s);
  if (p.w_bf16) launch_k(linear_tile_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, s, p);
  else launch_k(linear_tile_kernel<float>, dim3(grid), dim3(block), 0, s, p);
  ++g_launches;
}

void launch_linear_gemv(const LinearParams& p, cudaStream_t s) {
  ProfScope ps("linear_gemv", p.tag, linear_flops(p), linear_bytes(p), s);
  if (p.w_bf16) launch_gemv_t<__nv_bfloat16>(p, s);
  else launch_gemv_t<float>(p, s);
  ++g_launches;
}

}  // namespace ptts
