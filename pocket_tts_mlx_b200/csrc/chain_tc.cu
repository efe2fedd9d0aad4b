// Cluster chain kernel (see chain_tc.cuh): a dependent chain of small-M GEMMs with fused LayerNorm in one launch.
//
// One cluster of NC CTAs per 128-row tile of the batch.  Inside a CTA the roles are those of gemm_tc.cu:
//   warp 0    : TMA producer.  The A operand of a step ([128 rows x K <= 512] bf16, the same for all CTAs of the cluster)
//               is RESIDENT for the step: its eight [128 x 64] blocks are fetched once per cluster -- CTA b % NC loads
//               block b and multicasts it into every CTA's shared memory.  This CTA's own [bn x 64] W blocks stream
//               through a CTA-local ring; W blocks of the next step are issued before the step barrier (weights do not
//               depend on activations), the A blocks after it.  (A first version streamed A through a cluster-synchronous
//               ring, one cluster-wide round trip per 16 KB block: 274 us for the flow head instead of 133.)
//   warp 1    : tcgen05.mma issuer (M = 128, N = bn <= 64, two TMEM accumulator stages); after the last MMA that reads
//               the resident A it multicasts a commit to every CTA's "a_free" barrier;
//   warps 2-9 : epilogue: TMEM -> registers -> bias / activation / gate / residual -> global memory (the next step's
//               A operand, fetched through L2 by TMA).
// Steps are separated by a cluster-scope mbarrier ("opdone": every epilogue warp of every CTA arrives on every CTA's
// barrier after its global stores; the producers wait on it before they fetch the next step's A blocks).
// LayerNorm needs whole rows, which are spread over the cluster: each CTA computes (mean, M2) of its slice of a row,
// writes the pair into every CTA's shared memory (st.shared::cluster), a second cluster-scope mbarrier ("statsdone")
// orders that, and every thread merges the NC partials (Chan's parallel variance formula) for its own row.
#include "chain_tc.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "gemm_tc.cuh"
#include "tc_device.cuh"

namespace ptts {
namespace {

constexpr int kChStages = 8;                    // CTA-local ring of W blocks
constexpr uint32_t kChA = 16384;                // one [128 x 64] bf16 block of the resident A operand
constexpr int kChABlocks = 8;                   // K <= 512 per chunk
constexpr uint32_t kChW = 8192;                 // one [bn <= 64][64] W block
constexpr int kChThreads = 320;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_remote(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_addr, uint32_t cta) {
  const uint32_t r = map_remote(local_addr, cta);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// (n, mean, M2) of a set merged with another (Chan et al.)
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb == 0.f) return;
  const float nt = n + nb;
  const float d = mb - mean;
  mean += d * (nb / nt);
  m2 += m2b + d * d * (n * nb / nt);
  n = nt;
}

template <int NC>
__global__ void __launch_bounds__(kChThreads, 1) chain_kernel(const ChainOp* __restrict__ ops, const int n_ops, const int M,
                                                              long long* __restrict__ dbg) {
  // dbg (optional): [n_ops][8] clock64 stamps of CTA 0: producer {step barrier passed, A issued}, MMA {A resident, last
  // commit}, epilogue warp 2 {first accumulator ready, stores done, row statistics ready}
#define CH_STAMP(slot) do { if (dbg && blockIdx.x == 0) dbg[oi * 8 + (slot)] = clock64(); } while (0)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = base + kChABlocks * kChA;                     // W ring behind the resident A blocks
  const uint32_t stats = sW + kChStages * kChW;                     // [NC][128] float2 (mean, M2) of every CTA's row slices
  const uint32_t hb = stats + NC * 1024u;                           // [128] float4: the other column half's partial
  const uint32_t bars = hb + 2048u;
  const uint32_t full0 = bars, empty0 = bars + 8 * kChStages, tfull0 = empty0 + 8 * kChStages, tempty0 = tfull0 + 16,
                 opdone = tempty0 + 16, statsdone = opdone + 8, a_full = statsdone + 8, a_free = a_full + 8, tptr = a_free + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int m0 = (int)(blockIdx.x / NC) * 128;
  constexpr uint16_t kAll = (uint16_t)((1u << NC) - 1u);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kChStages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_free, NC);                       // one multicast commit from every CTA of the cluster
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 8);
    }
    mbar_init(opdone, NC * 8);                   // every epilogue warp of every CTA
    mbar_init(statsdone, NC * 4);                // the four row-quadrant warps that publish a CTA's partials
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                            // every CTA's barriers exist before anybody multicasts or arrives remotely
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr));
  pdl_sync();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t git = 0, cc = 0;                 // W stages issued, A chunks loaded
      for (int oi = 0; oi < n_ops; ++oi) {
        const ChainOp* op = ops + oi;
        const int bn = op->bn, nkb = op->K / 64, slice = op->N / NC;
        const int tiles = slice >= bn ? slice / bn : 1;
        const int n_base = slice >= bn ? (int)rank * slice : 0;
        const uint32_t wbytes = (uint32_t)bn * 128u;
        const int n_chunks = (nkb + kChABlocks - 1) / kChABlocks;      // > 1 only for single-tile steps (host-checked)
        for (int ch = 0; ch < n_chunks; ++ch, ++cc) {
          const int kb0 = ch * kChABlocks, nkc = min(kChABlocks, nkb - kb0);
          const int total = tiles * nkc;
          const int pre = total < kChStages ? total : kChStages;
          auto issue_w = [&](int i) {
            const uint32_t g = git + (uint32_t)i, s = g % kChStages, ph = (g / kChStages) & 1;
            const int t = i / nkc, kb = kb0 + i - t * nkc;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_expect_tx(full0 + 8 * s, wbytes);
            tma_load_2d(sW + s * kChW, &op->tm_w, full0 + 8 * s, kb * 64, n_base + t * bn);
          };
          for (int i = 0; i < pre; ++i) issue_w(i);        // weights first: they do not depend on the previous step
          if (cc > 0) mbar_wait(a_free, (cc - 1) & 1u);     // every CTA's MMAs have finished with the resident A
          if (ch == 0 && oi > 0) {
            mbar_wait_cluster(opdone, (uint32_t)(oi - 1) & 1u);      // the previous step's outputs are in global memory
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (ch == 0) CH_STAMP(0);
          mbar_expect_tx(a_full, (uint32_t)nkc * kChA);
          for (int b = 0; b < nkc; ++b)
            if ((uint32_t)b % NC == rank) tma_load_2d_mc(base + (uint32_t)b * kChA, &op->tm_a, a_full, (kb0 + b) * 64, m0, kAll);
          if (ch == 0) CH_STAMP(1);
          for (int i = pre; i < total; ++i) issue_w(i);
          git += (uint32_t)total;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t git = 0, tcount = 0, cc = 0;
      for (int oi = 0; oi < n_ops; ++oi) {
        const ChainOp* op = ops + oi;
        const int bn = op->bn, nkb = op->K / 64, slice = op->N / NC;
        const int tiles = slice >= bn ? slice / bn : 1;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int n_chunks = (nkb + kChABlocks - 1) / kChABlocks;
        for (int ch = 0; ch < n_chunks; ++ch, ++cc) {
          const int nkc = min(kChABlocks, nkb - ch * kChABlocks);
          mbar_wait(a_full, cc & 1u);
          tc_fence_after();
          if (ch == 0) CH_STAMP(2);
          for (int t = 0; t < tiles; ++t) {
            const uint32_t tc_ = tcount + (n_chunks == 1 ? (uint32_t)t : 0u);
            const uint32_t as = tc_ & 1, aph = (tc_ >> 1) & 1;
            if (ch == 0) {
              mbar_wait(tempty0 + 8 * as, aph ^ 1);
              tc_fence_after();
            }
            const uint32_t tmem_acc = tmem_base + as * 64;
            for (int kb = 0; kb < nkc; ++kb, ++git) {
              const uint32_t s = git % kChStages, ph = (git / kChStages) & 1;
              mbar_wait(full0 + 8 * s, ph);
              tc_fence_after();
              const uint64_t da = make_desc<128>(base + (uint32_t)kb * kChA);
              const uint64_t db = make_desc<128>(sW + s * kChW);
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ch | kb | k) != 0);
              tc_commit(empty0 + 8 * s);
            }
            if (ch == n_chunks - 1) tc_commit(tfull0 + 8 * as);
          }
          tc_commit_mc(a_free, kAll);              // this CTA is done reading the resident A blocks
          CH_STAMP(3);
        }
        tcount += (uint32_t)tiles;
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int row = m0 + r;
    const bool row_ok = row < M;
    uint32_t tcount = 0, ln_count = 0;
    for (int oi = 0; oi < n_ops; ++oi) {
      const ChainOp& op = ops[oi];
      const int bn = op.bn, slice = op.N / NC;
      const int tiles = slice >= bn ? slice / bn : 1;
      const int n_base = slice >= bn ? (int)rank * slice : 0;
      const bool st_ok = row_ok && (slice >= bn || rank == 0);      // a step narrower than the cluster: rank 0 stores
      const int chunks = bn / 32;
      for (int t = 0; t < tiles; ++t, ++tcount) {
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(tfull0 + 8 * as, aph);
        tc_fence_after();
        if (t == 0 && warp == 2 && lane == 0) CH_STAMP(4);
        const uint32_t tmem_acc = tmem_base + as * 64 + ((uint32_t)(quad * 32) << 16);
        const int n_tile = n_base + t * bn;
        if (op.kind == CH_RES_LN) {
          // ---- residual stream update + LayerNorm over the whole row (one tile per CTA by construction)
          float v[2][32];
          const int mine = chunks > half ? (chunks - half + 1) / 2 : 0;      // 32-column chunks of this thread: half, half + 2
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c >= mine) break;
            const int ch = half + 2 * c;
            uint32_t raw[32];
            __syncwarp();
            tc_ld32(tmem_acc + (uint32_t)(ch * 32), raw);
            const int n = n_tile + ch * 32;
            float (&w)[32] = v[c];
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __uint_as_float(raw[i]);
            if (op.bias) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(op.bias + n) + i);
                w[4 * i] += q.x; w[4 * i + 1] += q.y; w[4 * i + 2] += q.z; w[4 * i + 3] += q.w;
              }
            }
            if (op.gate && row_ok) {
              const float4* gp = reinterpret_cast<const float4*>(op.gate + (long long)row * op.gate_rs + n);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 q = __ldcg(gp + i);      // written by ANOTHER CTA earlier in this launch: not through L1
                w[4 * i] *= q.x; w[4 * i + 1] *= q.y; w[4 * i + 2] *= q.z; w[4 * i + 3] *= q.w;
              }
            }
            if (!op.x_init && row_ok) {
              const float4* xp = reinterpret_cast<const float4*>(op.x + (long long)row * op.x_rs + n);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 q = xp[i];
                w[4 * i] += q.x; w[4 * i + 1] += q.y; w[4 * i + 2] += q.z; w[4 * i + 3] += q.w;
              }
            }
            if (row_ok) {
              float4* xo = reinterpret_cast<float4*>(op.x + (long long)row * op.x_rs + n);
#pragma unroll
              for (int i = 0; i < 8; ++i) xo[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] = 0.f;
            }
          }
          // the accumulator is in registers: hand the TMEM stage back before the cluster-wide part
          tc_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * as) : "memory");
          if (op.ln_on) {
            float nl = (float)(32 * mine), mean = 0.f, m2 = 0.f;
            if (mine > 0) {
              float sum = 0.f;
#pragma unroll
              for (int c = 0; c < 2; ++c)
                if (c < mine) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) sum += v[c][i];
                }
              mean = sum / nl;
#pragma unroll
              for (int c = 0; c < 2; ++c)
                if (c < mine) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) { const float d = v[c][i] - mean; m2 = fmaf(d, d, m2); }
                }
            }
            if (half == 1) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hb + (uint32_t)r * 16u), "f"(mean), "f"(m2), "f"(nl), "f"(0.f) : "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) {
              float om, o2, on, pad;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(om), "=f"(o2), "=f"(on), "=f"(pad) : "r"(hb + (uint32_t)r * 16u));
              chan_merge(nl, mean, m2, on, om, o2);
              const uint32_t mine_addr = stats + (rank * 128u + (uint32_t)r) * 8u;
#pragma unroll 1
              for (uint32_t c = 0; c < (uint32_t)NC; ++c)
                asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(map_remote(mine_addr, c)), "f"(mean), "f"(m2) : "memory");
              __syncwarp();
              if (lane < NC) mbar_arrive_remote(statsdone, (uint32_t)lane);      // lane c tells CTA c
            }
            mbar_wait_cluster(statsdone, ln_count & 1u);
            ++ln_count;
            if (warp == 2 && lane == 0) CH_STAMP(6);
            float nt = 0.f, mt = 0.f, m2t = 0.f;
#pragma unroll 1
            for (uint32_t c = 0; c < (uint32_t)NC; ++c) {
              float pm, p2;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pm), "=f"(p2) : "r"(stats + (c * 128u + (uint32_t)r) * 8u));
              chan_merge(nt, mt, m2t, (float)slice, pm, p2);
            }
            const float rstd = rsqrtf(m2t / (float)op.N + op.ln_eps);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c >= mine) break;
              const int n = n_tile + (half + 2 * c) * 32;
              float (&w)[32] = v[c];
#pragma unroll
              for (int i = 0; i < 32; ++i) w[i] = (w[i] - mt) * rstd;
              if (op.ln_w) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 gq = __ldg(reinterpret_cast<const float4*>(op.ln_w + n) + i);
                  const float4 bq = __ldg(reinterpret_cast<const float4*>(op.ln_b + n) + i);
                  w[4 * i] = fmaf(w[4 * i], gq.x, bq.x); w[4 * i + 1] = fmaf(w[4 * i + 1], gq.y, bq.y);
                  w[4 * i + 2] = fmaf(w[4 * i + 2], gq.z, bq.z); w[4 * i + 3] = fmaf(w[4 * i + 3], gq.w, bq.w);
                }
              }
              if (op.mod_scale && row_ok) {
                const float4* sp = reinterpret_cast<const float4*>(op.mod_scale + (long long)row * op.mod_rs + n);
                const float4* hp = reinterpret_cast<const float4*>(op.mod_shift + (long long)row * op.mod_rs + n);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 sq = __ldcg(sp + i), hq = __ldcg(hp + i);
                  w[4 * i] = fmaf(w[4 * i], 1.f + sq.x, hq.x); w[4 * i + 1] = fmaf(w[4 * i + 1], 1.f + sq.y, hq.y);
                  w[4 * i + 2] = fmaf(w[4 * i + 2], 1.f + sq.z, hq.z); w[4 * i + 3] = fmaf(w[4 * i + 3], 1.f + sq.w, hq.w);
                }
              }
              if (row_ok) {
                uint4* yp = reinterpret_cast<uint4*>(op.h16 + (long long)row * op.h_rs + n);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  yp[i] = make_uint4(pack_bf16(w[8 * i], w[8 * i + 1]), pack_bf16(w[8 * i + 2], w[8 * i + 3]),
                                     pack_bf16(w[8 * i + 4], w[8 * i + 5]), pack_bf16(w[8 * i + 6], w[8 * i + 7]));
              }
            }
          }
          continue;
        }
        for (int ch = half; ch < chunks; ch += 2) {
          uint32_t raw[32];
          __syncwarp();
          tc_ld32(tmem_acc + (uint32_t)(ch * 32), raw);
          const int n = n_tile + ch * 32;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (op.bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(op.bias + n) + i);
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (!st_ok) continue;
          if (op.kind == CH_STORE16) {
            act32(v, op.act);
            uint4* yp = reinterpret_cast<uint4*>(op.y16 + (long long)row * op.y_rs + n);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              yp[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                 pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          } else if (op.kind == CH_STORE32) {
            float4* yp = reinterpret_cast<float4*>(op.y32 + (long long)row * op.y_rs + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) yp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {   // CH_FIN
            const float4* lp = reinterpret_cast<const float4*>(op.lat_in + (long long)row * op.lat_rs + n);
            float4* lo = reinterpret_cast<float4*>(op.lat_out + (long long)row * op.lat_rs + n);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = lp[i];
              v[4 * i] = fmaf(v[4 * i], op.out_scale, q.x); v[4 * i + 1] = fmaf(v[4 * i + 1], op.out_scale, q.y);
              v[4 * i + 2] = fmaf(v[4 * i + 2], op.out_scale, q.z); v[4 * i + 3] = fmaf(v[4 * i + 3], op.out_scale, q.w);
              lo[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (op.lat16) {
              uint4* yp = reinterpret_cast<uint4*>(op.lat16 + (long long)row * op.lat16_rs + n);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                yp[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty0 + 8 * as) : "memory");
      }
      // this warp's global stores of the step are done: tell every CTA of the cluster (their producers fetch them by TMA)
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      if (warp == 2 && lane == 0) CH_STAMP(5);
      if (lane < NC && oi + 1 < n_ops) mbar_arrive_remote(opdone, (uint32_t)lane);      // lane c tells CTA c
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody leaves while a peer may still multicast, store or arrive into this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

template <int NC>
constexpr size_t chain_smem() {
  return (size_t)kChABlocks * kChA + (size_t)kChStages * kChW + NC * 1024 + 2048 + 256 + 1024;
}

int g_nc = -1;

template <int NC>
bool chain_setup() {
  if (cudaFuncSetAttribute(chain_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem<NC>()) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (NC > 8 && cudaFuncSetAttribute(chain_kernel<NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(NC * 2); cfg.blockDim = dim3(kChThreads); cfg.dynamicSmemBytes = chain_smem<NC>();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, chain_kernel<NC>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return n >= 2;
}

}  // namespace

bool chain_available() { return chain_cluster_size() > 0; }

int chain_cluster_size() {
  if (g_nc >= 0) return g_nc;
  g_nc = 0;
  if (!gemm_tc_available()) return 0;
  const char* v = getenv("PTTS_CHAIN_NC");
  const int want = v ? atoi(v) : 16;
  if (want <= 0) return 0;
  if (want >= 16 && chain_setup<16>()) g_nc = 16;
  else if (chain_setup<8>()) g_nc = 8;
  return g_nc;
}

bool chain_encode_a(CUtensorMap* tm, const __nv_bfloat16* a, int M, int K, long long row_stride) {
  const unsigned long long dims[2] = {(unsigned long long)K, (unsigned long long)M};
  const unsigned long long str[1] = {(unsigned long long)row_stride * 2};
  const unsigned box[2] = {64u, 128u};
  return tc_encode_bf16(tm, a, 2, dims, str, box, 64);
}

bool chain_encode_w(CUtensorMap* tm, const __nv_bfloat16* w, int N, int K, int bn) {
  const unsigned long long dims[2] = {(unsigned long long)K, (unsigned long long)N};
  const unsigned long long str[1] = {(unsigned long long)K * 2};
  const unsigned box[2] = {64u, (unsigned)bn};
  return tc_encode_bf16(tm, w, 2, dims, str, box, 64);
}

int chain_pick_bn(int N, int nc) {
  if (N % nc) return (N % 32 == 0 && N < nc * 32) ? 32 : 0;
  const int slice = N / nc;
  for (int bn : {64, 32})
    if (slice % bn == 0) return bn;
  return (N % 32 == 0 && slice < 32) ? 32 : 0;
}

bool chain_step_ok(int N, int K, int nc) {
  const int bn = chain_pick_bn(N, nc);
  if (bn == 0 || K % 64) return false;
  const int slice = N / nc, tiles = slice >= bn ? slice / bn : 1;
  return K <= 64 * kChABlocks || tiles == 1;
}

void chain_launch(const ChainOp* d_ops, int n_ops, int M, int nc, const char* tag, double flops, double bytes, cudaStream_t s) {
  if (n_ops <= 0 || M <= 0) return;
  ProfScope ps("chain", tag, flops, bytes, s);
  cudaLaunchConfig_t cfg{};
  const int m_tiles = (M + 127) / 128;
  cfg.gridDim = dim3((unsigned)(nc * m_tiles)); cfg.blockDim = dim3(kChThreads); cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl_on ? 2 : 1;
  // PTTS_CHAIN_DBG=1: per-step clock stamps of CTA 0, printed after a synchronisation (debugging / profiling only)
  static const bool dbg_on = [] { const char* v = getenv("PTTS_CHAIN_DBG"); return v && v[0] == '1'; }();
  static long long* d_dbg = nullptr;
  if (dbg_on && !d_dbg) cudaMalloc((void**)&d_dbg, 64 * 8 * sizeof(long long));
  long long* dbg = (dbg_on && n_ops <= 64) ? d_dbg : nullptr;
  if (dbg) cudaMemsetAsync(dbg, 0, 64 * 8 * sizeof(long long), s);
  if (nc == 16) {
    cfg.dynamicSmemBytes = chain_smem<16>();
    cudaLaunchKernelEx(&cfg, chain_kernel<16>, d_ops, n_ops, M, dbg);
  } else {
    cfg.dynamicSmemBytes = chain_smem<8>();
    cudaLaunchKernelEx(&cfg, chain_kernel<8>, d_ops, n_ops, M, dbg);
  }
  ++g_launches;
  if (dbg) {
    long long h[64 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost);
    const long long t0 = h[1] ? h[1] : h[0];
    fprintf(stderr, "[chain %s] M=%d nc=%d  (cycles since the first A issue)\n  op  barrier  a_issue  a_ready  mma_done  acc_ready  stats  stored\n", tag ? tag : "?", M, nc);
    for (int i = 0; i < n_ops; ++i) {
      const long long* r = h + i * 8;
      auto rel = [&](long long v) { return v ? (long long)(v - t0) : -1LL; };
      fprintf(stderr, "  %2d %8lld %8lld %8lld %9lld %10lld %6lld %7lld\n", i, rel(r[0]), rel(r[1]), rel(r[2]), rel(r[3]), rel(r[4]), rel(r[6]), rel(r[5]));
    }
  }
}

}  // namespace ptts
