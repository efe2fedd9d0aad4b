// Fused SEANet tail for sm_100a: last residual block + output convolution in one persistent tcgen05 kernel.
//
// Reference maths (pocket_tts_mlx/modules/seanet.py:9-42 resblock, :150-170 last layers; streaming convs
// pocket_tts_mlx/modules/conv.py:74-150):
//     h     = ELU(conv_k3(ELU(x)) + b1)              64 -> 32 channels, causal, 2 carried input columns
//     y     = x + conv_k1(h) + b2                    32 -> 64
//     audio = conv_k3(ELU(y)) + bf                   64 -> 1,  causal, 2 carried input columns
// At batch 256 the 64-channel level is 491 520 time steps per frame; run as three kernels it moves the
// [491 520 x 64] tensor through HBM five times.  Here a 128-step tile goes
//     TMA (3 shifted boxes of ELU(x), one box of x) -> tcgen05.mma #1 (K = 3 x 64, N = 32) -> TMEM
//       -> epilogue group A: +b1, ELU, bf16, written to shared memory in the UMMA K-major 64B-swizzled layout
//       -> tcgen05.mma #2 (K = 32, N = 64) -> TMEM
//       -> epilogue group B: +b2, +x (from the TMA-staged tile), ELU, the three 64-wide dot products of the
//          output convolution, neighbour exchange through shared memory, fp32 audio store.
// Nothing but the audio (and 3 floats per tile) is written.  Every barrier is an mbarrier; all TMEM / smem
// buffers are double-buffered so the MMAs of tile i+1 overlap both epilogues of tile i.
//
// The output convolution needs y[t-2], y[t-1] from the previous tile (or the previous frame): each tile writes
// its last partial products (w0.y[126], w0.y[127], w1.y[127]) to `bnd`, and sn_tail_fix_kernel adds them to the
// first two samples of the next tile; slot 0 of `bnd` is the state carried from the previous frame.
#include "gemm_tc.cuh"
#include "tc_device.cuh"

#include <algorithm>
#include <cstdlib>

namespace ptts {
namespace {

constexpr int kTailThreads = 512;          // warps: 0 operand TMA, 1 MMA, 2 residual TMA, 4-7 epilogue A, 8-15 epilogue B
constexpr int kTailStages = 8;             // 16 KB operand chunks: 2 2/3 tiles in flight
constexpr int kResBufs = 3;                // residual tiles in flight
constexpr uint32_t kABytes = 16384;        // [128 steps][64 ch] bf16, 128B swizzle
constexpr uint32_t kW1Bytes = 3 * 4096;    // 3 taps x [32][64] bf16, 128B swizzle
constexpr uint32_t kW2Bytes = 4096;        // [64][32] bf16, 64B swizzle
constexpr uint32_t kHBytes = 8192;         // [128][32] bf16, 64B swizzle
constexpr uint32_t kResBytes = 16384;      // [128][64] bf16, 128B swizzle
constexpr uint32_t kConstFloats = 32 + 64 + 192 + 4;   // b1, b2, wf[3][64], bf
constexpr uint32_t kTailSmem = kTailStages * kABytes + kW1Bytes + kW2Bytes + 2 * kHBytes + kResBufs * kResBytes +
                               kConstFloats * 4 + 2 * 2 * 3 * 128 * 4 + 512 + 1024;

struct TailArgs {
  int nb, T, tiles_t, total_tiles;
  const float *b1, *b2, *wf, *bf;
  float* audio;
  long long audio_bs;
  float* bnd;
  short* pcm;            // optional 16-bit PCM copy of the samples (samples 0, 1 of a tile are finished by the fix-up)
  int stream_loads;      // L2 evict-first on the last reads of the input tensors
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : __expf(v) - 1.0f; }
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(kTailThreads, 1) sn_tail_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                  const __grid_constant__ CUtensorMap tm_w1,
                                                                  const __grid_constant__ CUtensorMap tm_w2,
                                                                  const __grid_constant__ CUtensorMap tm_res,
                                                                  const TailArgs g) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sW1 = sA + kTailStages * kABytes;
  const uint32_t sW2 = sW1 + kW1Bytes;
  const uint32_t sH = sW2 + kW2Bytes;
  const uint32_t sRes = sH + 2 * kHBytes;
  const uint32_t sConst = sRes + kResBufs * kResBytes;
  const uint32_t sPx = sConst + kConstFloats * 4;          // [2 tiles][2 halves][3][128] fp32
  const uint32_t bars = sPx + 2 * 2 * 3 * 128 * 4;
  const uint32_t full0 = bars, empty0 = full0 + 8 * kTailStages, wfull = empty0 + 8 * kTailStages;
  const uint32_t d1f = wfull + 8, d1e = d1f + 16, hf = d1e + 16, he = hf + 16, d2f = he + 16, d2e = d2f + 16,
                 rf = d2e + 16, re = rf + 8 * kResBufs, tptr = re + 8 * kResBufs;
  float* cst = reinterpret_cast<float*>(smem_raw + (sConst - smem_u32(smem_raw)));
  float* px = reinterpret_cast<float*>(smem_raw + (sPx - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_res) : "memory");
    for (int i = 0; i < kTailStages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1f + 8 * i, 1); mbar_init(d1e + 8 * i, 4);
      mbar_init(hf + 8 * i, 4);  mbar_init(he + 8 * i, 1);
      mbar_init(d2f + 8 * i, 1); mbar_init(d2e + 8 * i, 8);
    }
    for (int i = 0; i < kResBufs; ++i) {
      mbar_init(rf + 8 * i, 1);
      mbar_init(re + 8 * i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights-only constants (not produced by the preceding kernel)
  for (int i = threadIdx.x; i < 32 + 64 + 192 + 1; i += kTailThreads) {
    float v;
    if (i < 32) v = g.b1[i];
    else if (i < 96) v = g.b2[i - 32];
    else if (i < 288) v = g.wf[i - 96];
    else v = g.bf[0];
    cst[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  pdl_wait();

  const int n_my = (g.total_tiles > (int)blockIdx.x) ? (g.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(wfull, kW1Bytes + kW2Bytes);
      for (int tap = 0; tap < 3; ++tap) tma_load_2d(sW1 + tap * 4096u, &tm_w1, wfull, tap * 64, 0);
      tma_load_2d(sW2, &tm_w2, wfull, 0, 0);
      uint32_t git = 0;
      // the [491 520 x 64] tensors are read here for the last time: the third (last) tap box of a tile and the residual
      // box go in as L2 evict-first
      const unsigned long long pol = l2_policy(g.stream_loads ? L2_EVICT_FIRST : L2_DEFAULT);
      for (int i = 0; i < n_my; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int b = tile / g.tiles_t, t0 = (tile % g.tiles_t) * 128;
        for (int tap = 0; tap < 3; ++tap, ++git) {
          const uint32_t s = git % kTailStages, ph = (git / kTailStages) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, kABytes);
          tma_load_3d(sA + s * kABytes, &tm_a, full0 + 8 * s, 0, t0 + tap, b, tap == 2 ? pol : 0ull);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc<32>(), idesc2 = make_idesc<64>();
      mbar_wait(wfull, 0);
      uint32_t git = 0;
      for (int i = 0; i <= n_my; ++i) {
        if (i < n_my) {                                  // conv_k3: D1 = sum_tap A_tap . W1_tap^T
          const uint32_t as = i & 1, aph = (i >> 1) & 1;
          mbar_wait(d1e + 8 * as, aph ^ 1);
          tc_fence_after();
          const uint32_t d1 = tmem_base + as * 32;
          for (int tap = 0; tap < 3; ++tap, ++git) {
            const uint32_t s = git % kTailStages, ph = (git / kTailStages) & 1;
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint64_t da = make_desc<128>(sA + s * kABytes);
            const uint64_t db = make_desc<128>(sW1 + tap * 4096u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma(d1, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc1, (tap | k) != 0);
            tc_commit(empty0 + 8 * s);
          }
          tc_commit(d1f + 8 * as);
        }
        if (i >= 1) {                                    // conv_k1 of the previous tile: D2 = H . W2^T
          const int j = i - 1;
          const uint32_t as = j & 1, aph = (j >> 1) & 1;
          mbar_wait(hf + 8 * as, aph);
          mbar_wait(d2e + 8 * as, aph ^ 1);
          tc_fence_after();
          const uint32_t d2 = tmem_base + 64 + as * 64;
          const uint64_t da = make_desc<64>(sH + as * kHBytes);
          const uint64_t db = make_desc<64>(sW2);
#pragma unroll
          for (int k = 0; k < 2; ++k) tc_mma(d2, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, k != 0);
          tc_commit(he + 8 * as);
          tc_commit(d2f + 8 * as);
        }
      }
    }
  } else if (warp == 2) {
    // residual tiles (raw x) on their own producer, so the operand ring never waits for the last epilogue
    if (lane == 0) {
      uint32_t rb = 0, rph = 0;                          // buffer / phase kept as explicit counters (see note in group B)
      const unsigned long long pol = l2_policy(g.stream_loads ? L2_EVICT_FIRST : L2_DEFAULT);
      for (int i = 0; i < n_my; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int b = tile / g.tiles_t, t0 = (tile % g.tiles_t) * 128;
        mbar_wait(re + 8 * rb, rph ^ 1);
        mbar_expect_tx(rf + 8 * rb, kResBytes);
        tma_load_3d(sRes + rb * kResBytes, &tm_res, rf + 8 * rb, 0, t0, b, pol);
        if (++rb == kResBufs) { rb = 0; rph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue group A: hidden activations -> shared-memory A operand of the second MMA
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    for (int i = 0; i < n_my; ++i) {
      const uint32_t as = i & 1, aph = (i >> 1) & 1;
      mbar_wait(d1f + 8 * as, aph);
      tc_fence_after();
      uint32_t raw[32];
      tc_ld32(tmem_base + as * 32 + ((uint32_t)(quad * 32) << 16), raw);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d1e + 8 * as);
      float v[32];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 bb = lds4(sConst + 16u * k);
        v[4 * k] = elu1(__uint_as_float(raw[4 * k]) + bb.x);
        v[4 * k + 1] = elu1(__uint_as_float(raw[4 * k + 1]) + bb.y);
        v[4 * k + 2] = elu1(__uint_as_float(raw[4 * k + 2]) + bb.z);
        v[4 * k + 3] = elu1(__uint_as_float(raw[4 * k + 3]) + bb.w);
      }
      mbar_wait(he + 8 * as, aph ^ 1);                   // MMA #2 of tile i-2 has finished reading this buffer
      const uint32_t row = sH + as * kHBytes + (uint32_t)r * 64u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t chunk = (uint32_t)c ^ (uint32_t)((r >> 1) & 3);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + chunk * 16u),
                     "r"(pack_bf16(v[8 * c], v[8 * c + 1])), "r"(pack_bf16(v[8 * c + 2], v[8 * c + 3])),
                     "r"(pack_bf16(v[8 * c + 4], v[8 * c + 5])), "r"(pack_bf16(v[8 * c + 6], v[8 * c + 7])) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(hf + 8 * as);
    }
  } else if (warp >= 8) {
    // ---- epilogue group B (two warps per TMEM lane quadrant, 32 channels each): residual, ELU, the three
    //      partial dot products of the output convolution; the `half == 0` warps combine and store
    const int quad = warp & 3, half = (warp - 8) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t sB2 = sConst + 128u + 128u * half;              // b2[32 half ..]
    const uint32_t sWf = sConst + 384u + 128u * half;              // wf[j][32 half ..], j stride 256 B
    const float bf = cst[288];
    // NOTE: the residual buffer index is an explicit counter, not i % kResBufs: with the modulo form ptxas 12.9
    // strength-reduced `rf + 8 * (i % 3)` into an induction register that it also used as the base of the px
    // stores below, so they drifted by 8 bytes per tile (seen in SASS; PTX was correct).
    uint32_t rb = 0, rph = 0;
    for (int i = 0; i < n_my; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int b = tile / g.tiles_t, tt = tile % g.tiles_t, t0 = tt * 128;
      const uint32_t as = i & 1, aph = (i >> 1) & 1;
      mbar_wait(d2f + 8 * as, aph);
      tc_fence_after();
      uint32_t raw[32];
      tc_ld32(tmem_base + 64 + as * 64 + 32 * half + ((uint32_t)(quad * 32) << 16), raw);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d2e + 8 * as);
      mbar_wait(rf + 8 * rb, rph);
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
      const uint32_t rrow = sRes + rb * kResBytes + (uint32_t)r * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                      // 8 channels per 16-byte chunk
        const uint32_t chunk = (uint32_t)(4 * half + c) ^ (uint32_t)(r & 7);
        uint4 q;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                     : "r"(rrow + chunk * 16u));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
        float y[8];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float4 bb = lds4(sB2 + 32u * c + 16u * k);
          const float2 x0 = __bfloat1622float2(h[2 * k]), x1 = __bfloat1622float2(h[2 * k + 1]);
          y[4 * k] = elu1(__uint_as_float(raw[8 * c + 4 * k]) + bb.x + x0.x);
          y[4 * k + 1] = elu1(__uint_as_float(raw[8 * c + 4 * k + 1]) + bb.y + x0.y);
          y[4 * k + 2] = elu1(__uint_as_float(raw[8 * c + 4 * k + 2]) + bb.z + x1.x);
          y[4 * k + 3] = elu1(__uint_as_float(raw[8 * c + 4 * k + 3]) + bb.w + x1.y);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float4 w0 = lds4(sWf + 32u * c + 16u * k), w1 = lds4(sWf + 256u + 32u * c + 16u * k),
                       w2 = lds4(sWf + 512u + 32u * c + 16u * k);
          p0 = fmaf(y[4 * k], w0.x, p0); p0 = fmaf(y[4 * k + 1], w0.y, p0); p0 = fmaf(y[4 * k + 2], w0.z, p0); p0 = fmaf(y[4 * k + 3], w0.w, p0);
          p1 = fmaf(y[4 * k], w1.x, p1); p1 = fmaf(y[4 * k + 1], w1.y, p1); p1 = fmaf(y[4 * k + 2], w1.z, p1); p1 = fmaf(y[4 * k + 3], w1.w, p1);
          p2 = fmaf(y[4 * k], w2.x, p2); p2 = fmaf(y[4 * k + 1], w2.y, p2); p2 = fmaf(y[4 * k + 2], w2.z, p2); p2 = fmaf(y[4 * k + 3], w2.w, p2);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(re + 8 * rb);
      if (++rb == kResBufs) { rb = 0; rph ^= 1; }
      float* pxa = px + as * 768 + half * 384;           // [tile parity][half][3][128]
      pxa[r] = p0;
      pxa[128 + r] = p1;
      pxa[256 + r] = p2;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (half == 0) {
        const float* pa = px + as * 768;
        const float* pb = pa + 384;
        float out = p2 + pb[256 + r] + bf;
        if (r >= 2) out += pa[r - 2] + pb[r - 2] + pa[128 + r - 1] + pb[128 + r - 1];
        else if (r == 1) out += pa[128] + pb[128];
        g.audio[b * g.audio_bs + t0 + r] = out;
        if (g.pcm) g.pcm[b * g.audio_bs + t0 + r] = pcm16_of(out);
        if (r == 127) {
          float* bd = g.bnd + ((long long)b * (g.tiles_t + 1) + tt + 1) * 4;
          bd[0] = pa[126] + pb[126]; bd[1] = pa[127] + pb[127]; bd[2] = pa[255] + pb[255];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// first two samples of every tile: add the partial products that live in the previous tile / previous frame,
// then carry this frame's last ones into slot 0
__global__ void sn_tail_fix_kernel(float* __restrict__ audio, long long audio_bs, float* __restrict__ bnd, int nb, int tiles_t) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb * tiles_t) return;
  const int b = i / tiles_t, k = i % tiles_t;
  float* slot = bnd + ((long long)b * (tiles_t + 1) + k) * 4;
  const float f0 = slot[0], f1 = slot[1], f2 = slot[2];
  float* a = audio + b * audio_bs + (long long)k * 128;
  a[0] += f0 + f2;
  a[1] += f1;
  if (k == 0) {
    const float* last = bnd + ((long long)b * (tiles_t + 1) + tiles_t) * 4;
    slot[0] = last[0]; slot[1] = last[1]; slot[2] = last[2];
  }
}

bool g_tail_attr = false;

}  // namespace

bool sn_tail_plan(SnTail* p, const __nv_bfloat16* xe, const __nv_bfloat16* xraw, int nb, int T, int C, int hidden,
                  int taps, int fin_taps, const __nv_bfloat16* w1, const __nv_bfloat16* w2) {
  p->valid = false;
  if (!gemm_tc_available() || C != 64 || hidden != 32 || taps != 3 || fin_taps != 3 || T % 128 != 0 || nb < 1) return false;
  if (!g_tail_attr) {
    if (cudaFuncSetAttribute(sn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    g_tail_attr = true;
  }
  bool ok = true;
  {
    const unsigned long long dims[3] = {64ull, (unsigned long long)(T + 2), (unsigned long long)nb};
    const unsigned long long str[2] = {128ull, (unsigned long long)(T + 2) * 128ull};
    const unsigned box[3] = {64u, 128u, 1u};
    ok = ok && tc_encode_bf16(&p->tm_a, xe, 3, dims, str, box, 64);
  }
  {
    const unsigned long long dims[3] = {64ull, (unsigned long long)T, (unsigned long long)nb};
    const unsigned long long str[2] = {128ull, (unsigned long long)T * 128ull};
    const unsigned box[3] = {64u, 128u, 1u};
    ok = ok && tc_encode_bf16(&p->tm_res, xraw, 3, dims, str, box, 64);
  }
  {
    const unsigned long long dims[2] = {192ull, 32ull};
    const unsigned long long str[1] = {384ull};
    const unsigned box[2] = {64u, 32u};
    ok = ok && tc_encode_bf16(&p->tm_w1, w1, 2, dims, str, box, 64);
  }
  {
    const unsigned long long dims[2] = {32ull, 64ull};
    const unsigned long long str[1] = {64ull};
    const unsigned box[2] = {32u, 64u};
    ok = ok && tc_encode_bf16(&p->tm_w2, w2, 2, dims, str, box, 32);
  }
  p->nb = nb; p->T = T;
  p->valid = ok;
  return ok;
}

void sn_tail_launch(const SnTail& p, cudaStream_t s, bool with_fix) {
  TailArgs a;
  a.nb = p.nb; a.T = p.T; a.tiles_t = p.T / 128; a.total_tiles = p.nb * a.tiles_t;
  a.b1 = p.b1; a.b2 = p.b2; a.wf = p.wf; a.bf = p.bf;
  a.audio = p.audio; a.audio_bs = p.audio_bs; a.bnd = p.bnd; a.pcm = p.pcm;
  { static const int sl = [] { const char* v = getenv("PTTS_SN_STREAM"); return (v && v[0] == '0') ? 0 : 1; }(); a.stream_loads = sl; }
  {
    const double rows = (double)p.nb * p.T;
    ProfScope ps("sn_tail", nullptr, 2.0 * rows * (192.0 * 32 + 32.0 * 64 + 192.0), rows * (64 * 2 * 2 + 4), s);
    launch_k(sn_tail_kernel, dim3((unsigned)std::min(a.total_tiles, gemm_tc_grid_cap() > 0 ? std::min(148, gemm_tc_grid_cap()) : 148)), dim3(kTailThreads), (size_t)kTailSmem, s,
             p.tm_a, p.tm_w1, p.tm_w2, p.tm_res, a);
    ++g_launches;
  }
  if (with_fix) {
    ProfScope ps("sn_tail_fix", nullptr, 0, 0, s);
    launch_k(sn_tail_fix_kernel, dim3((unsigned)((a.total_tiles + 127) / 128)), dim3(128), 0, s, p.audio, p.audio_bs, p.bnd,
             p.nb, a.tiles_t);
    ++g_launches;
  }
}

}  // namespace ptts
