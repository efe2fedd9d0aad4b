// Cluster "chain" kernel: a whole dependent chain of small-M GEMMs in ONE launch (see chain_tc.cu).
//
// The FlowLM decode step and the flow head at batch <= 512 are chains of GEMMs over the same few hundred rows
// (M = batch), each too small to fill the machine and each costing 5-9 us as its own launch whatever its size
// (launch + TMEM/barrier set-up + TMA -> MMA -> TMEM -> epilogue fill and drain).  The rows of such a chain are
// independent of each other, so a thread-block CLUSTER can own a 128-row tile for the whole chain: its CTAs split the
// N axis of every GEMM, the A operand of a step is streamed once per cluster with TMA multicast, the output of a step
// goes to global memory (L2) as the next step's A operand, and steps are separated by cluster-scope mbarrier
// hand-shakes instead of kernel boundaries.  LayerNorm (plain or AdaLN-modulated) needs whole rows: it is fused into
// the epilogue of the GEMM that produces the residual stream, with the row statistics exchanged through distributed
// shared memory.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace ptts {

enum ChainEpi : int {
  CH_STORE16 = 0,   // y16 = act(acc + bias)                                   (bf16: next step's A operand)
  CH_STORE32 = 1,   // y32 = acc + bias                                        (fp32: AdaLN modulation table)
  CH_RES_LN = 2,    // x = [x +] gate * (acc + bias); h16 = mod(LN(x))          (residual stream + fused LayerNorm)
  CH_FIN = 3,       // lat_out = lat_in + out_scale * (acc + bias)              (Euler update of the latent, N = latent_dim)
};

struct alignas(128) ChainOp {
  CUtensorMap tm_a;          // A [M][K] bf16, box {64, 128 rows}, 128-byte swizzle
  CUtensorMap tm_w;          // W [N][K] bf16, box {64, bn rows}
  int K, N, bn, kind;
  int act;                   // CH_STORE16: Act
  float out_scale;           // CH_FIN
  const float* bias;         // [N] or null
  __nv_bfloat16* y16; float* y32; long long y_rs;
  // CH_RES_LN
  float* x; long long x_rs; int x_init;                  // x_init: x = gate * (acc + bias), no residual read
  const float* gate; long long gate_rs;                   // per (row, col) or null
  int ln_on; const float *ln_w, *ln_b; float ln_eps;      // ln_w / ln_b null: no affine
  const float *mod_scale, *mod_shift; long long mod_rs;   // h = LN(x) * (1 + scale) + shift, or null
  __nv_bfloat16* h16; long long h_rs;
  // CH_FIN
  const float* lat_in; float* lat_out; long long lat_rs; __nv_bfloat16* lat16; long long lat16_rs;
};

bool chain_available();
// largest cluster size (16 or 8) the device can co-schedule for the chain kernel; 0 when unavailable
int chain_cluster_size();
bool chain_encode_a(CUtensorMap* tm, const __nv_bfloat16* a, int M, int K, long long row_stride);
bool chain_encode_w(CUtensorMap* tm, const __nv_bfloat16* w, int N, int K, int bn);
// N tile (64 or 32) for a step of width N split over nc CTAs (0: not divisible).  Steps with K > 512 keep their
// accumulator across the K chunks and must be a single tile per CTA: chain_step_ok checks both.
int chain_pick_bn(int N, int nc);
bool chain_step_ok(int N, int K, int nc);
// ops: DEVICE array of n_ops ChainOp (128-byte aligned); M rows; nc = cluster size the ops were planned for
void chain_launch(const ChainOp* d_ops, int n_ops, int M, int nc, const char* tag, double flops, double bytes, cudaStream_t s);

}  // namespace ptts
