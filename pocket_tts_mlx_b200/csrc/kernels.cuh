// Kernel-side declarations shared by the engine and the .cu translation units.
// Everything here is device plumbing for the hot path of SURVEY.md section 8; no host-side fallbacks.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace ptts {

constexpr int kHeadDim = 64;      // both transformers use 64-wide heads (b6369a24.yaml)
constexpr int kPageTokens = 32;   // FlowLM KV page: 32 tokens x H x 64

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_ELU = 3 };

// Y[b,t,n] = epi( sum_{j<taps} sum_{c<C} pro(A[b, t+j, c]) * W[n, j*C + c] )
//   epi(v) = res[b,t,n] + col_scale[n] * row_gate[b,t,n] * out_scale * act(v + bias[n])
// A rows of sequence b start at A + b*a_bs, consecutive time rows are a_rs apart.  The first taps-1
// rows of each sequence are carried streaming state (previous inputs), so a k-tap causal conv, a
// polyphase transposed conv (taps = 2, N = stride*C_out) and a plain Linear (taps = 1) are one operator.
struct LinearParams {
  const float* A; long long a_bs, a_rs;
  int nb, T, taps, C;
  const void* W; int w_bf16;                 // [N][taps*C], bf16 or fp32 storage
  int N;
  const float* bias;                          // [N] or null
  int a_pro;                                  // Act applied to A on load (ELU in front of SEANet convs)
  int act;                                    // Act applied to the accumulator (+bias)
  float out_scale;
  const float* row_gate; long long gate_bs, gate_rs;   // null or [b,t,n] multiplier (AdaLN gate)
  const float* col_scale;                     // null or [N] (LayerScale)
  const float* res; long long res_bs, res_rs; // null or residual with Y's row indexing
  float* Y; long long y_bs, y_rs;
  const char* tag;                            // call-site label for the profiler (may be null)
  // GEMV path only: the weights are read once per frame and are too many to stay in L2 (FlowLM layers at batch 1): load
  // them with the streaming (evict-first) hint so that they do not evict the flow head's and the Mimi decoder's
  int w_stream;
  // GEMV path only: LayerNorm (optionally AdaLN-modulated) fused into the row staging -- every CTA holds the full
  // input rows in shared memory anyway, so the separate norm launch disappears from the batch-1 chain
  int ln_on; const float* ln_w; const float* ln_b; float ln_eps;
  const float* ln_scale; const float* ln_shift; long long ln_mod_rs;   // per-row modulation vectors or null
  // GEMV path, FlowLM qkv at <= 4 rows (two output features per warp = one RoPE pair): rotate q / k and append k, v to
  // the paged cache in the epilogue instead of a separate flow_rope_append launch.  Y is not written.
  int rope_on; float* q_rot; void* kv_layer; int kv_bf16;
  const int *kv_row_seq, *kv_row_pos, *kv_page_table; int kv_max_pages, kv_heads; long long kv_page_stride;
  const float* rope_freqs;
};
bool linear_gemv_rope_supported(const LinearParams& p);   // GEMV path with a (2i, 2i+1) pair per warp

inline double linear_flops(const LinearParams& p) { return 2.0 * p.nb * p.T * (double)p.N * p.taps * p.C; }
inline double linear_bytes(const LinearParams& p) {
  const double w = (double)p.N * p.taps * p.C * (p.w_bf16 ? 2 : 4);
  const double a = (double)p.nb * (p.T + p.taps - 1) * p.C * 4;
  const double y = (double)p.nb * p.T * p.N * 4 * (p.res ? 2 : 1);
  return w + a + y;
}

void launch_linear_tile(const LinearParams& p, cudaStream_t s);   // any M; SIMT 64x64 tiles
void launch_linear_gemv(const LinearParams& p, cudaStream_t s);
// y[M][N] = x[M][K] W[N][K]^T + b for K = 32 (bf16 weights); false when the shape is not covered
bool launch_small_k_linear(const __nv_bfloat16* W, const float* bias, const float* X, float* Y, int M, int K, int N,
                           const char* tag, cudaStream_t s);   // M = nb*T <= 16; weight-streaming
bool linear_gemv_supported(const LinearParams& p);

// y = LN(x; w, b, eps) [* (1 + scale) + shift]; rows of width C; scale/shift are per-row vectors
// (AdaLN modulation, modules/mlp.py:11-13,95-100) or null.
struct NormParams {
  const float* X; long long x_bs, x_rs;
  int nb, T, C;
  const float* w; const float* b; float eps;
  const float* scale; const float* shift; long long mod_rs;   // row stride of the modulation buffer
  float* Y; long long y_bs, y_rs;
  __nv_bfloat16* Y16;            // when set, the result is written here as bf16 (same strides) instead of Y
  // split-K partial planes [acc_n][rows][C] of the producing GEMM: x += sum of planes (written back to X)
  const float* acc; int acc_n; long long acc_stride;
};
void launch_layernorm(const NormParams& p, cudaStream_t s);

// ---- FlowLM attention over the paged KV pool ------------------------------------------------------
// Pool layout: [layer][page][k|v][head][slot(32)][64], element bf16 or fp32.
struct FlowAttnParams {
  const float* qkv;            // [M][3*H*64]  (q | k | v), pre-RoPE
  float* q_rot;                // [M][H*64] scratch
  float* out;                  // [M][H*64]
  __nv_bfloat16* out16;        // when set, bf16 output instead (A operand of the next tcgen05 GEMM)
  void* pool; int kv_bf16; long long layer_stride, page_stride;   // strides in elements
  int layer;
  const int* row_seq;          // [M] sequence of each row, or null => row m belongs to sequence m
  const int* row_pos;          // [M] absolute position of each row (decode: the sequence length)
  const int* page_table; int max_pages;   // [n_seq][max_pages]
  int M, H;
  const float* freqs;          // [32] RoPE frequencies (fp32, computed like modules/rope.py:17-18)
  long long total_keys;        // host-side sum over rows of (row_pos+1), for the profiler's byte count
  int splits; float* part;     // split-KV for small batches: partials [M][H][splits][66] merged by a second kernel
                               // (a last-slice-merges variant with an arrival counter was measured: +10 us per frame)
  // cascade: all rows share the first prefix_len keys (the voice prompt).  A tensor-core kernel computes that part
  // once per (row, head) from the voice's own pages into prefix_part [M][H][66]; the per-sequence kernel then
  // starts at key prefix_len and merges the partial.
  int prefix_len; const int* prefix_pages; float* prefix_part;
  // prefill (many rows per sequence): rows of sequence s are seq_row0[s] .. seq_row0[s+1]-1 at positions
  // seq_pos0[s] + t; set by the text / voice prefill so that the tensor-core prefill kernel can be used
  const int* seq_row0; const int* seq_pos0; int n_seq, max_rows_per_seq;
  // host pointer to two CUtensorMaps over the bf16 pool viewed as [page][k|v][head][slot][64] (128-byte swizzle):
  // [0] box = one head's K and V slices of a whole page (8 KB), [1] box = 8 key slots of K or V (1 KB); or null.
  // Lets the decode attention stream K/V with TMA (flow_attention_stream_kernel).
  const void* kv_tmap;
  // debugging (PTTS_ATTN_DBG=1): [layer][2] {earliest CTA start, latest CTA end} in globaltimer nanoseconds, or null
  unsigned long long* tstamp;
  // folded cascade (stream kernel only): this layer's {claim counter, flag per (16 rows, head) tile} block, zero when the
  // launch starts, and the block of the layer launched next (zeroed by this launch).  Null: the prefix partial comes
  // from flow_prefix_attention_kernel (its own launch).
  int* pflags; int* pflags_next;
  int kv_evict_first;      // stream kernel: the private K/V boxes carry an L2 evict-first policy
};
void launch_flow_prefix_attention(const FlowAttnParams& p, cudaStream_t s);
// causal attention of whole prefill chunks on tensor cores (bf16 KV, bf16 output); false when not applicable
bool launch_flow_prefill_attention(const FlowAttnParams& p, cudaStream_t s);
// Mimi encoder (voice cloning): RoPE in place on q,k of qkv [T][3*H*64], then windowed causal attention -> out [T][H*64]
void launch_enc_attention(float* qkv, float* out, const float* freqs, int T, int H, int context, cudaStream_t s);
void launch_enc_conv0(const float* xpad, const float* w, const float* b, float* y, long long T, int N, int k, cudaStream_t s);
void launch_replicate_row(float* dst, const float* src, int n, int C, cudaStream_t s);
void launch_flow_rope_append(const FlowAttnParams& p, cudaStream_t s);
void launch_rope_table(const int* row_pos, const float* freqs, float* table, int M, int T, cudaStream_t s);
void launch_flow_attention(const FlowAttnParams& p, cudaStream_t s);
// true when launch_flow_attention(p) would run the persistent TMA-stream kernel in the geometry that can fold the
// cascade prefix in (p.pflags)
bool flow_attention_streams(const FlowAttnParams& p);
// device buffer [64][2] of the attention time stamps (allocated on first use), or null when PTTS_ATTN_DBG is off
unsigned long long* flow_attention_dbg_buffer();

// ---- Mimi ring-buffer attention --------------------------------------------------------------------
// Ring layout: [layer][k|v][seq][head][slot(context)][64].
struct MimiAttnParams {
  const float* qkv;            // [B*T][3*H*64]
  float* q_rot;                // [B*T][H*64]
  float* out;                  // [B*T][H*64]
  __nv_bfloat16* out16;
  void* ring; int kv_bf16; long long layer_stride, kv_stride;
  int layer;
  const int* offset;           // [B] absolute position of this chunk's first step (= end_offset)
  int B, T, H, context;
  const float* freqs;
};
void launch_mimi_rope_ring(const MimiAttnParams& p, cudaStream_t s);
void launch_mimi_attention(const MimiAttnParams& p, cudaStream_t s);

// ---- small fused ops --------------------------------------------------------------------------------
// x[b] = W_in (bos_flag[b] ? bos : prev[b])            (models/flow_lm.py:93-94)
void launch_input_rows(const float* w_in, const float* bos, const float* prev, const int* bos_flag,
                       float* x, int B, int D, int L, cudaStream_t s);
// rows[m] = table[ids[m]]                               (conditioners/text.py:43-45)
void launch_embed_rows(const void* table, int table_bf16, const int* ids, float* rows, int M, int D,
                       cudaStream_t s);
// c = LN(x[row_of[b]]); logit = w_eos . c + b_eos      (models/flow_lm.py:120,100)
void launch_final_norm_eos(const float* x, const int* row_of, const float* ln_w, const float* ln_b,
                           const float* w_eos, const float* b_eos, float* c, __nv_bfloat16* c16, float* logit,
                           int B, int D, const float* acc, int acc_n, long long acc_stride, cudaStream_t s,
                           // optional: the flow head's start noise of each row is prepared by the same launch
                           const float* nz = nullptr, float* x0 = nullptr, int nL = 0, float nstd = 1.f,
                           float nclamp = -1.f, int use_philox = 0, const unsigned long long* counter = nullptr,
                           // optional bf16 copy of x0 as rows of 64 (A operand of the chain kernel's input projection)
                           __nv_bfloat16* x0_16 = nullptr);
// x0 = clip(sqrt(temp) * z); z from the host buffer or a Philox4x32-10 + Box-Muller stream
void launch_noise_prep(const float* z, float* x0, int n, float std, float clamp, int use_philox,
                       const unsigned long long* counter, cudaStream_t s);
// z = Wq (lat*std + mean); up[b,t,c] = wu[c,t] z[c] + wu[c,S+t] zprev[b,c]; zprev = z
void launch_quant_upsample(const float* lat, const float* emb_std, const float* emb_mean, const float* wq,
                           const float* wu, float* zprev, float* out, long long out_bs, int B, int L, int C,
                           int S, cudaStream_t s);
// audio[b,t] = bias + sum_{j<taps} sum_c elu(x~[b,t+j,c]) w[j*C+c]      (SEANet last conv, N = 1)
// pcm (optional, same indexing as audio): the sample as 16-bit PCM, clip(v, -1, 1) * 32767 truncated toward zero --
// the conversion of the reference's StreamingWAVWriter.write_pcm_data (data/audio.py:64-70) done where the sample is made
void launch_final_conv(const float* x, long long x_bs, const float* w, const float* bias, float* audio,
                       long long audio_bs, int B, int T, int C, int taps, cudaStream_t s, short* pcm = nullptr);
// same with a bf16 input that already went through ELU in the producer's epilogue
void launch_final_conv16(const __nv_bfloat16* x, long long x_bs, const float* w, const float* bias, float* audio,
                         long long audio_bs, int B, int T, int C, int taps, cudaStream_t s, short* pcm = nullptr);
// carried conv state: move the last `rows` time rows of each sequence's buffer to its front
struct ShiftEntry { void* buf; long long bs; int T, rows, C, esz; };   // bs in elements, esz = bytes/element
void launch_state_shift(const ShiftEntry* entries_dev, int n_entries, int B, cudaStream_t s, float* audio = nullptr,
                        long long audio_bs = 0, float* bnd = nullptr, int tiles_t = 0, int* mimi_offset = nullptr,
                        int inc_mimi = 0, short* pcm = nullptr);
// seq_len += inc_len; bos_flag = 0; mimi_offset += inc_mimi; philox counter += 1
void launch_advance(int* seq_len, int* bos_flag, int* mimi_offset, unsigned long long* counter, int B,
                    int inc_len, int inc_mimi, cudaStream_t s, const int* active = nullptr);
// y = a * x + y   (Euler update x += v / n)
void launch_axpy(const float* x, float* y, float a, int n, cudaStream_t s);
void launch_copy_pages(void* pool, int kv_bf16, long long layer_stride, long long page_stride, int n_layers,
                       const int* src_pages, const int* dst_pages, int n_pairs, cudaStream_t s);
void launch_fill_u32(unsigned int* dst, unsigned int v, long long n, cudaStream_t s);
// continuous batching: the streaming state of n_slots sequences back to a template (or to zero when tpl is null) in one
// launch.  A piece is one state buffer: sequence b's part is bytes long at base + b * stride, its image in the template
// at tpl + tpl_off (bytes and addresses are multiples of 4).
struct StatePieceDev { char* base; unsigned long long stride, bytes, tpl_off; };
void launch_restore_state(const StatePieceDev* pieces, int n_pieces, const int* slots, int n_slots, const char* tpl, cudaStream_t s);
void launch_gather_frame(const float* lat_all, float* lat, int B, int F, int L, const int* frame_idx,
                         cudaStream_t s);
void launch_scatter_audio(const float* audio, float* audio_all, int B, int F, int n, const int* frame_idx,
                          cudaStream_t s);
void launch_inc(int* v, int inc, cudaStream_t s);

// launch accounting (ptts_launch_count) and the eager per-kernel profiler (ptts_batch_profile_step):
// every launcher opens a ProfScope; when profiling is on it brackets the launch with CUDA events on the
// launching stream and books the ALGORITHMIC flops / bytes of that launch under "kernel:tag".
extern std::atomic<long long> g_launches;   // process-wide: contexts may be driven from different threads
extern bool g_prof_on;
struct ProfScope {
  int slot = -1;
  bool nvtx = false;      // PTTS_NVTX=1: an NVTX range "kernel:tag" around the launch (ncu --nvtx --nvtx-include, Nsight Systems)
  cudaStream_t s;
  ProfScope(const char* kernel, const char* tag, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
};
void prof_start();
// "name,launches,ms,flops,bytes\n" per kernel:tag, sorted by time; stops profiling
const char* prof_report();

// ---- programmatic dependent launch -----------------------------------------------------------------------
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it may be scheduled while
// its predecessor in the stream / graph is still draining.  Kernels call pdl_trigger() first (lets THEIR
// successor be scheduled early) and pdl_wait() before the first access to global memory that the predecessor
// may have written (griddepcontrol.wait returns once the predecessor grid has completed and flushed).
extern std::atomic<bool> g_pdl_on;   // PTTS_NO_PDL=1 turns the launch attribute off (the device calls are then no-ops)
extern std::atomic<bool> g_launch_prio_on;
extern thread_local int g_launch_prio;
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

template <typename... KP, typename... Args>
inline void launch_k(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (g_pdl_on) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (g_launch_prio_on) {          // PTTS_PRIO=2: explicit per-launch priority (kept by a captured kernel node)
    at[n].id = cudaLaunchAttributePriority;
    at[n].val.priority = g_launch_prio;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);
}

// ---- device helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_SILU: return v / (1.0f + __expf(-v));
    case ACT_ELU: return v > 0.0f ? v : expm1f(v);
    default: return v;
  }
}

// float sample -> 16-bit PCM exactly as NumPy does it in the reference: (clip(v, -1, 1) * 32767).astype(int16)
__device__ __forceinline__ short pcm16_of(float v) {
  return (short)__float2int_rz(fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void epilogue_store(const LinearParams& p, int b, int t, int n, float v) {
  if (p.bias) v += p.bias[n];
  v = act_apply(v, p.act) * p.out_scale;
  if (p.row_gate) v *= p.row_gate[b * p.gate_bs + t * p.gate_rs + n];
  if (p.col_scale) v *= p.col_scale[n];
  if (p.res) v += p.res[b * p.res_bs + t * p.res_rs + n];
  p.Y[b * p.y_bs + t * p.y_rs + n] = v;
}

}  // namespace ptts
