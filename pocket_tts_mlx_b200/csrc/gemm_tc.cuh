// tcgen05 / TMEM / TMA multi-tap GEMM (bf16 operands, fp32 accumulation in tensor memory).
//
//   D[b,t,n] = sum_{j<taps} sum_{c<C} A[b, t+j, c] * W[n, j*C + c]
//
// A is a bf16 activation buffer [nb][T+taps-1][C] (the first taps-1 rows of every sequence are the carried
// streaming state) described by a 3-D TMA tensor map, so tap j of a causal conv / polyphase transposed conv
// is just the same box fetched at time coordinate t0+j; plain Linears use taps = 1.  W is [N][taps*C] bf16
// (K-major), a 2-D tensor map.  One CTA computes a 128 x BN tile: warp 0 issues TMA into a 4-stage
// 128B-swizzled ring, warp 1 issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator, warps 2-5 drain
// it with tcgen05.ld and apply the fused epilogue.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace ptts {

struct TcEpilogue {
  const float* bias;          // [N] or null
  int act;                    // Act on (acc + bias)
  float out_scale;            // 0 means 1
  const float* row_gate; long long gate_bs, gate_rs;     // [b,t,n] multiplier or null
  const float* col_scale;     // [N] or null
  const float* res32; long long res32_bs, res32_rs;      // fp32 residual or null
  const __nv_bfloat16* res16; long long res16_bs, res16_rs;
  float* y32; long long y32_bs, y32_rs;                  // fp32 output or null
  __nv_bfloat16* y16; long long y16_bs, y16_rs; int y16_act;   // bf16 output of act2(v) or null
  __nv_bfloat16* yraw16; long long yraw16_bs, yraw16_rs;       // bf16 copy of v or null
  // FlowLM qkv projection (N = 3*H*64, columns q | k | v): RoPE and the KV-cache append happen in the epilogue.
  // q is rotated and stored fp32 to q_rot [M][H*64]; k is rotated and, like v, written as bf16 into the row's page
  // slot.  rope_cs [M][64] holds cos[32] | sin[32] of each row's position (rope_table_kernel).  Replaces
  // flow_rope_append_kernel (pocket_tts_mlx/modules/attention.py:145-148,50-61; rope.py:9-42).
  const float* rope_cs;
  float* q_rot;
  __nv_bfloat16* kv_layer;                 // pool + layer * layer_stride
  const int *kv_row_seq, *kv_row_pos, *kv_page_table;
  int kv_max_pages, kv_heads;
  long long kv_page_stride;
  // kv_ring > 0: Mimi ring cache instead of pages (modules/attention.py:67-105): row (b, t) sits at position
  // kv_row_pos[b] + t, slot = position % kv_ring of [b][head][slot][64]; V lives kv_v_offset elements after K
  int kv_ring;
  long long kv_v_offset;
};

struct TcGemm {
  CUtensorMap tm_a, tm_b;
  CUtensorMap tm_y16, tm_yraw16;   // output maps for the TMA-store epilogue (valid when tma_store)
  CUtensorMap tm_res;              // bf16 residual map (valid when res_tma)
  bool tma_store = false;
  bool res_tma = false;
  int nb, T, taps, C, N;
  int box_t, box_b;           // 128-row M tile = box_b sequences x box_t time steps
  int bn, bk;                 // N tile (32/64/128), K chunk in elements (64 or 32)
  int stages;                 // smem ring depth (2..8)
  int stg_tiles;              // 16 KB staging tiles for TMA-store outputs
  int persist;                // 1: persistent CTAs with two TMEM accumulator stages; 0: one tile per CTA
  int splits;                 // split-K factor; > 1: raw fp32 partials go to split_ws[z][nb*T][N], epilogue skipped
  float* split_ws;
  int a_policy = 0, w_policy = 0;   // L2Policy (tc_device.cuh) of the A / W operand loads, set by the engine per call site
  const void* pf_ptr = nullptr; long long pf_bytes = 0;   // weights of the next GEMM of a latency-bound chain (L2 prefetch), or null
  TcEpilogue e;
  const char* tag;
  bool valid = false;
};

void gemm_tc_init();
bool gemm_tc_available();
// Describe one GEMM call site.  a: bf16 [nb][T+taps-1][C] with strides (a_bs, a_rs) in elements;
// w: bf16 [N][taps*C].  Returns false when the shape is unsupported (caller keeps the SIMT path).
bool gemm_tc_plan(TcGemm* g, const __nv_bfloat16* a, long long a_bs, long long a_rs, int nb, int T, int taps, int C,
                  const __nv_bfloat16* w, int N, const char* tag, int max_splits = 1, int n_bf16_out = 0);
// Call after the epilogue pointers are set: bf16 outputs of N-tiles >= 64 leave through a swizzled
// shared-memory tile and cp.async.bulk.tensor stores instead of per-thread row stores.
void gemm_tc_bind_outputs(TcGemm* g);
void gemm_tc_launch(const TcGemm& g, cudaStream_t s);
// SM partition between the branches of the pipelined frame graph: while cap > 0 the persistent kernels (GEMM, SEANet tail)
// launch at most `cap` CTAs, leaving the other SMs to the kernels of the other branch.  Set around the launches of a
// branch at capture time (single-threaded per context).
void gemm_tc_set_grid_cap(int cap);
int gemm_tc_grid_cap();
// fp32-in / fp32-out debug entry used by ptts_debug_linear(path=3); returns < 0 when unsupported.
int gemm_tc_debug(const LinearParams& p, bool bf16_storage, cudaStream_t s);

// Kernel-level benchmark (ptts_debug_gemm_bench): median microseconds of `reps` launches on synthetic bf16
// operands with L2 flushed between launches.  force = {bn, stages, splits, persist} or all zeros for the planner's
// own choice; epi: number of bf16 outputs (0 = one fp32 output), +4 adds a bf16 residual read.
int gemm_tc_bench(int nb, int T, int taps, int C, int N, int epi, const int force[4], int reps, float* us,
                  int chosen[4], cudaStream_t s);

// bf16 tiled tensor map; swizzle_elems = 64 (128-byte swizzle) or 32 (64-byte swizzle) = inner box extent
bool tc_encode_bf16(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                    const unsigned long long* strides_bytes, const unsigned* box, int swizzle_elems);

// Fused SEANet tail (seanet_tail.cu): the last residual block and the output convolution in one kernel,
//   y = x + W2 . ELU(W1 (*) ELU(x) + b1) + b2 ;  audio[t] = sum_j wf[j] . ELU(y[t-2+j]) + bf
// for C = 64 channels, hidden 32, 3-tap causal convs, T % 128 == 0.  `xe` = ELU(x) behind 2 carried state rows
// ([nb][T+2][64] bf16), `xraw` = x ([nb][T][64] bf16); `bnd` ([nb][T/128+1][4] fp32, zero at stream start) carries
// the three partial products that cross tile / frame boundaries.
struct SnTail {
  CUtensorMap tm_a, tm_w1, tm_w2, tm_res;
  int nb = 0, T = 0;
  const float *b1 = nullptr, *b2 = nullptr, *wf = nullptr, *bf = nullptr;
  float* audio = nullptr; long long audio_bs = 0;
  float* bnd = nullptr;
  short* pcm = nullptr;       // optional int16 PCM output (same strides as audio)
  bool valid = false;
};
bool sn_tail_plan(SnTail* p, const __nv_bfloat16* xe, const __nv_bfloat16* xraw, int nb, int T, int C, int hidden,
                  int taps, int fin_taps, const __nv_bfloat16* w1, const __nv_bfloat16* w2);
// with_fix = false: the caller finishes the first two samples of every tile itself (state_shift_kernel does it in the
// Mimi decoder's end-of-frame launch)
void sn_tail_launch(const SnTail& p, cudaStream_t s, bool with_fix = true);

// helpers used by the bf16 pipeline
void launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t s);

}  // namespace ptts
