// tcgen05 / TMEM / TMA multi-tap GEMM for sm_100a.  See gemm_tc.cuh for the operator definition.
//
// Hardware mapping
//   * operands: TMA (cp.async.bulk.tensor) loads 128 x BK (A, 3-D box = sequences x time x channels) and
//     BN x BK (W) bf16 tiles into a 4-stage shared-memory ring, 128-byte (BK=64) or 64-byte (BK=32) swizzle;
//   * math: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) with shared
//     memory matrix descriptors; the fp32 accumulator lives in BN TMEM columns;
//   * pipeline: full/empty mbarriers between the TMA warp and the MMA warp (tcgen05.commit frees a stage),
//     one tmem_full mbarrier towards the four epilogue warps;
//   * epilogue: tcgen05.ld 32x32b.x32 (thread = accumulator row), fused bias / activation / AdaLN gate /
//     LayerScale / residual, fp32 and/or bf16 stores (the bf16 copy feeds the next GEMM's TMA).
#include "gemm_tc.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace ptts {
namespace {

constexpr int kThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps

struct KArgs {
  int nb, T, taps, C, N;
  int box_t, box_b, tiles_t;
  int splits;                   // gridDim.z; split z covers k-iterations [z*ips, (z+1)*ips)
  long long split_stride;       // elements between the partial planes of split_ws
  float* split_ws;
  TcEpilogue e;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major operand, rows of ROW_BYTES (= swizzle span) bytes
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t layout = (ROW_BYTES == 128) ? 2ull : 4ull;   // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint64_t sbo = (8 * ROW_BYTES) >> 4;                  // stride between 8-row groups
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// activation on 32 accumulator values with the selector hoisted out of the element loop (uniform branch)
__device__ __forceinline__ void act32(float (&v)[32], int act) {
  if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752440f));
  } else if (act == ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __fdividef(v[i], 1.0f + __expf(-v[i]));
  } else if (act == ACT_ELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.0f ? v[i] : __expf(v[i]) - 1.0f;   // output is rounded to bf16
  }
}

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                           const __grid_constant__ CUtensorMap tm_b, const KArgs g) {
  constexpr int ROW_BYTES = BK * 2;
  constexpr int A_BYTES = 128 * ROW_BYTES, B_BYTES = BN * ROW_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * A_BYTES;
  const uint32_t bars = sB + STAGES * B_BYTES;
  // bars: full[STAGES], empty[STAGES], tmem_full ; then the TMEM base address word
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull = bars + 16 * STAGES, tptr = tfull + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int tile_b = blockIdx.y / g.tiles_t, tile_t = blockIdx.y % g.tiles_t;
  const int b0 = tile_b * g.box_b, t0 = tile_t * g.box_t;
  const int kc_per_tap = g.C / BK;
  const int iters_all = g.taps * kc_per_tap;
  const int ips = iters_all / g.splits;
  const int it0 = blockIdx.z * ips;
  const int iters = ips;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_acc) : "r"(tptr));
  // barriers, TMEM and descriptors are set up; everything below touches the predecessor's output
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        const int git = it0 + it;
        const int tap = git / kc_per_tap, kc = git - tap * kc_per_tap;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
        tma_load_3d(sA + s * A_BYTES, &tm_a, full0 + 8 * s, kc * BK, t0 + tap, b0);
        tma_load_2d(sB + s * B_BYTES, &tm_b, full0 + 8 * s, tap * g.C + kc * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        const uint64_t da = make_desc<ROW_BYTES>(sA + s * A_BYTES);
        const uint64_t db = make_desc<ROW_BYTES>(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          tc_mma(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
        tc_commit(empty0 + 8 * s);    // frees the stage once these MMAs have read it
      }
      tc_commit(tfull);               // accumulator complete
    }
  } else {
    // ---- epilogue: warps 2..9; TMEM lane quadrant = warp % 4, two warps per quadrant split the columns ----
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int CHUNKS = BN / 32;
    const int r = quad * 32 + lane;                       // accumulator row within the tile
    const int b = b0 + r / g.box_t, t = t0 + r % g.box_t;
    const bool row_ok = (b < g.nb) && (t < g.T);
    const TcEpilogue& e = g.e;
    const float oscale = e.out_scale == 0.f ? 1.f : e.out_scale;
    mbar_wait(tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int ch = half; ch < CHUNKS; ch += 2) {
      const int cb = ch * 32;
      uint32_t raw[32];
      __syncwarp();
      tc_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, raw);
      if (row_ok && g.splits > 1) {
        // split-K: raw fp32 partial; the consumer (LayerNorm) adds the planes to the residual stream
        float4* yp = reinterpret_cast<float4*>(g.split_ws + blockIdx.z * g.split_stride +
                                               ((long long)b * g.T + t) * g.N + n0 + cb);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          yp[i] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                              __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
      } else if (row_ok) {
        const int n = n0 + cb;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
        if (e.bias) {
          const float4* bp = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = __ldg(bp + i);
            v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
        }
        act32(v, e.act);
        if (oscale != 1.f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= oscale;
        }
        if (e.col_scale) {
          const float4* cp = reinterpret_cast<const float4*>(e.col_scale + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = __ldg(cp + i);
            v[4 * i] *= q.x; v[4 * i + 1] *= q.y; v[4 * i + 2] *= q.z; v[4 * i + 3] *= q.w;
          }
        }
        if (e.row_gate) {
          const float4* gp = reinterpret_cast<const float4*>(e.row_gate + b * e.gate_bs + t * e.gate_rs + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = gp[i];
            v[4 * i] *= q.x; v[4 * i + 1] *= q.y; v[4 * i + 2] *= q.z; v[4 * i + 3] *= q.w;
          }
        }
        if (e.res32) {
          const float4* rp = reinterpret_cast<const float4*>(e.res32 + b * e.res32_bs + t * e.res32_rs + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = rp[i];
            v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
          }
        }
        if (e.res16) {
          const uint4* rp = reinterpret_cast<const uint4*>(e.res16 + b * e.res16_bs + t * e.res16_rs + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 q = rp[i];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y;
            }
          }
        }
        if (e.y32) {
          float4* yp = reinterpret_cast<float4*>(e.y32 + b * e.y32_bs + t * e.y32_rs + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) yp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (e.yraw16) {
          uint4* yp = reinterpret_cast<uint4*>(e.yraw16 + b * e.yraw16_bs + t * e.yraw16_rs + n);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            yp[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        }
        if (e.y16) {
          act32(v, e.y16_act);
          uint4* yp = reinterpret_cast<uint4*>(e.y16 + b * e.y16_bs + t * e.y16_rs + n);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            yp[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, int BK, int STAGES>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * (128 * BK * 2 + BN * BK * 2) + 1024 + 16 * STAGES + 32;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
bool g_init_done = false;

template <int BN, int BK, int ST>
void set_attr1() {
  if (smem_bytes<BN, BK, ST>() <= 227 * 1024)
    cudaFuncSetAttribute(gemm_tc_kernel<BN, BK, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<BN, BK, ST>());
}
template <int BN, int BK>
void set_attr() {
  set_attr1<BN, BK, 2>(); set_attr1<BN, BK, 4>(); set_attr1<BN, BK, 6>(); set_attr1<BN, BK, 8>();
}

bool encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, int bk) {
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (bk == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                        strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_sync();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

}  // namespace

void gemm_tc_init() {
  if (g_init_done) return;
  g_init_done = true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = (EncodeFn)fn;
  else
    cudaGetLastError();
  set_attr<128, 64>(); set_attr<64, 64>(); set_attr<32, 64>();
  set_attr<128, 32>(); set_attr<64, 32>(); set_attr<32, 32>();
}

bool gemm_tc_available() { return g_encode != nullptr; }

bool gemm_tc_plan(TcGemm* g, const __nv_bfloat16* a, long long a_bs, long long a_rs, int nb, int T, int taps, int C,
                  const __nv_bfloat16* w, int N, const char* tag, int max_splits) {
  g->valid = false;
  if (!g_encode) return false;
  int bk = 0;
  if (C % 64 == 0) bk = 64;
  else if (C % 32 == 0) bk = 32;
  else return false;
  if (N % 32 != 0 || (a_rs % 8) != 0 || (a_bs % 8) != 0) return false;
  // M tiling: 128 rows = box_b sequences x box_t time steps
  int box_t, box_b;
  if (nb == 1) { box_t = 128; box_b = 1; }
  else {
    box_t = 1;
    while (box_t < 128 && T % (box_t * 2) == 0) box_t *= 2;
    box_b = 128 / box_t;
  }
  const int tiles_t = (T + box_t - 1) / box_t;
  const int tiles_b = (nb + box_b - 1) / box_b;
  const long long m_tiles = (long long)tiles_t * tiles_b;
  // Tile / split / ring-depth choice by a small cost model (microseconds).  A TMA ring turn costs ~2 us of
  // latency, so an SM fills shared memory at min(120 GB/s, resident CTAs x stages x stage bytes / 2 us); all SMs
  // together are limited to ~5 TB/s of L2 -> SM traffic; every tile pays a fixed prologue and an epilogue that
  // overlaps only with co-resident CTAs.
  const int iters = taps * (C / bk);
  int bn = 0, best_splits = 1, best_stages = 2;
  double best = 1e30;
  for (int cand : {128, 64, 32}) {
    if (N % cand) continue;
    for (int sp = 1; sp <= max_splits; sp *= 2) {
      if (iters % sp || (sp > 1 && iters / sp < 2)) continue;
      const int it = iters / sp;
      const double stage_bytes = 128.0 * bk * 2 + (double)cand * bk * 2;
      const double ctas = (double)m_tiles * (N / cand) * sp;
      for (int st : {2, 4, 6, 8}) {
        if (st > 2 && st > it + 1) continue;
        if (st * stage_bytes + 2048 > 220.0 * 1024) continue;
        int per_sm = (int)(220.0 * 1024 / (st * stage_bytes + 2048));
        per_sm = per_sm > 4 ? 4 : per_sm;
        const double per_sm_ctas = std::ceil(ctas / 148.0);
        const double resident = std::min<double>(per_sm, per_sm_ctas);
        const double rate = std::min(120e3, resident * st * stage_bytes / 2.0);          // bytes / us / SM
        const double t_main = per_sm_ctas * it * stage_bytes / rate;
        const double t_epi = (0.5 + 0.02 * cand) * per_sm_ctas / resident;               // exposed epilogue + prologue
        const double t_agg = ctas * it * stage_bytes / 5e6;
        const double t = std::max(t_main, t_agg) + t_epi + 2.0 + (sp > 1 ? 0.3 : 0.0);
        if (t < best) { best = t; bn = cand; best_splits = sp; best_stages = st; }
      }
    }
  }
  if (!bn) return false;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)(T + taps - 1), (cuuint64_t)nb};
    const cuuint64_t str[2] = {(cuuint64_t)a_rs * 2, (cuuint64_t)(nb == 1 ? (long long)(T + taps - 1) * a_rs : a_bs) * 2};
    const cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_t, (cuuint32_t)box_b};
    if (!encode(&g->tm_a, a, 3, dims, str, box, bk)) return false;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)taps * C, (cuuint64_t)N};
    const cuuint64_t str[1] = {(cuuint64_t)taps * C * 2};
    const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
    if (!encode(&g->tm_b, w, 2, dims, str, box, bk)) return false;
  }
  g->nb = nb; g->T = T; g->taps = taps; g->C = C; g->N = N;
  g->box_t = box_t; g->box_b = box_b; g->bn = bn; g->bk = bk;
  g->stages = best_stages; g->splits = best_splits; g->split_ws = nullptr;
  g->e = TcEpilogue{};
  g->tag = tag;
  g->valid = true;
  static const bool verbose = [] { const char* v = getenv("PTTS_TC_VERBOSE"); return v && v[0] == '1'; }();
  if (verbose)
    fprintf(stderr, "[gemm_tc] %-10s nb=%d T=%d taps=%d C=%d N=%d -> box %dx%d bn=%d bk=%d stages=%d splits=%d est=%.1fus\n",
            tag ? tag : "?", nb, T, taps, C, N, box_b, box_t, bn, bk, best_stages, best_splits, best);
  return true;
}

void gemm_tc_launch(const TcGemm& g, cudaStream_t s) {
  KArgs a;
  a.nb = g.nb; a.T = g.T; a.taps = g.taps; a.C = g.C; a.N = g.N;
  a.box_t = g.box_t; a.box_b = g.box_b;
  a.tiles_t = (g.T + g.box_t - 1) / g.box_t;
  a.e = g.e;
  a.splits = g.splits; a.split_ws = g.split_ws; a.split_stride = (long long)g.nb * g.T * g.N;
  const int tiles_b = (g.nb + g.box_b - 1) / g.box_b;
  dim3 grid(g.N / g.bn, a.tiles_t * tiles_b, g.splits), block(kThreads);
  const double flops = 2.0 * g.nb * g.T * (double)g.N * g.taps * g.C;
  const double bytes = (double)g.N * g.taps * g.C * 2 + (double)g.nb * (g.T + g.taps - 1) * g.C * 2 +
                       (double)g.nb * g.T * g.N * ((g.e.y32 ? 4 : 0) + (g.e.y16 ? 2 : 0) + (g.e.yraw16 ? 2 : 0) +
                                                   (g.e.res32 ? 4 : 0) + (g.e.res16 ? 2 : 0));
  ProfScope ps("gemm_tc", g.tag, flops, bytes, s);
#define PTTS_TC(BN_, BK_)                                                                                       \
  do {                                                                                                          \
    switch (g.stages) {                                                                                         \
      case 8: launch_k(gemm_tc_kernel<BN_, BK_, 8>, grid, block, smem_bytes<BN_, BK_, 8>(), s, g.tm_a, g.tm_b, a); break; \
      case 6: launch_k(gemm_tc_kernel<BN_, BK_, 6>, grid, block, smem_bytes<BN_, BK_, 6>(), s, g.tm_a, g.tm_b, a); break; \
      case 4: launch_k(gemm_tc_kernel<BN_, BK_, 4>, grid, block, smem_bytes<BN_, BK_, 4>(), s, g.tm_a, g.tm_b, a); break; \
      default: launch_k(gemm_tc_kernel<BN_, BK_, 2>, grid, block, smem_bytes<BN_, BK_, 2>(), s, g.tm_a, g.tm_b, a); break; \
    }                                                                                                           \
  } while (0)
  if (g.bk == 64) {
    if (g.bn == 128) PTTS_TC(128, 64);
    else if (g.bn == 64) PTTS_TC(64, 64);
    else PTTS_TC(32, 64);
  } else {
    if (g.bn == 128) PTTS_TC(128, 32);
    else if (g.bn == 64) PTTS_TC(64, 32);
    else PTTS_TC(32, 32);
  }
#undef PTTS_TC
  ++g_launches;
}

void launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t s) {
  ProfScope ps("f32_to_bf16", nullptr, 0, 6.0 * n, s);
  launch_k(f32_to_bf16_kernel, dim3((int)std::min<long long>((n + 255) / 256, 4096)), dim3(256), 0, s, src, dst, n);
  ++g_launches;
}

int gemm_tc_debug(const LinearParams& p, bool bf16_storage, cudaStream_t s) {
  if (!bf16_storage || !g_encode) return -1;
  const long long na = (long long)p.nb * (p.T + p.taps - 1) * p.C;
  __nv_bfloat16* a16 = nullptr;
  if (cudaMalloc((void**)&a16, na * 2) != cudaSuccess) return -1;
  launch_f32_to_bf16(p.A, a16, na, s);
  TcGemm g;
  int rc = 0;
  if (!gemm_tc_plan(&g, a16, p.a_bs, p.a_rs, p.nb, p.T, p.taps, p.C, (const __nv_bfloat16*)p.W, p.N, "debug")) rc = -1;
  if (rc == 0) {
    g.e.bias = p.bias;
    g.e.act = p.act;
    g.e.y32 = p.Y; g.e.y32_bs = p.y_bs; g.e.y32_rs = p.y_rs;
    gemm_tc_launch(g, s);
  }
  cudaStreamSynchronize(s);
  cudaFree(a16);
  return rc;
}

}  // namespace ptts
