/*
 * ptts.h -- C ABI of libptts_b200.so: pocket-tts streaming generation on NVIDIA B200 (sm_100a).
 *
 * The reference (jishnuvenugopal/pocket-tts-mlx) has no FFI: its hot path is Python over the MLX
 * tensor runtime.  This header is the boundary a host runtime binds instead of MLX for that path;
 * every entry point names the reference code it replaces (paths relative to
 * /root/reference/pocket_tts_mlx/).  The Python facade `pocket_tts_mlx_b200.TTSModel` binds it with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative ptts_status; ptts_last_error() gives the
 *     thread-local message of the last failure;
 *   - all pointers are HOST pointers to plain arrays whose lifetime spans the call; the library owns
 *     every device allocation;
 *   - one ptts_ctx per GPU; a ctx (and the objects made from it) must be used by one thread at a
 *     time; different ctxs may be used concurrently from different threads or processes;
 *   - there is no CPU fallback: without a CUDA device ptts_ctx_create fails with PTTS_ERR_CUDA.
 */
#ifndef PTTS_H_
#define PTTS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTTS_ABI_VERSION 1

typedef enum {
  PTTS_OK = 0,
  PTTS_ERR_INVALID = -1,   /* bad argument / shape / unknown name */
  PTTS_ERR_CUDA = -2,      /* CUDA runtime or driver failure (includes "no device") */
  PTTS_ERR_STATE = -3,     /* call order violated (e.g. step before prefill) */
  PTTS_ERR_NOMEM = -4,     /* KV page pool or device memory exhausted */
  PTTS_ERR_MISSING = -5    /* a required checkpoint tensor was never loaded */
} ptts_status;

typedef enum { PTTS_BF16 = 0, PTTS_FP32 = 1 } ptts_precision;
typedef enum { PTTS_DT_F32 = 0, PTTS_DT_BF16 = 1, PTTS_DT_F16 = 2 } ptts_dtype;

/* Model geometry + sampling knobs.  Field meanings follow config/b6369a24.yaml and the kwargs of
 * TTSModel.load_model (models/tts_model.py:202-221). */
typedef struct {
  /* FlowLM backbone (models/flow_lm.py, modules/mimi_transformer.py:104-114) */
  int32_t d_model, n_heads, n_layers, ffn_dim, n_bins, latent_dim;
  float max_period;
  /* flow head (modules/mlp.py:122-168) */
  int32_t flow_dim, flow_depth;
  /* Mimi decoder transformer (modules/mimi_transformer.py:123-171) */
  int32_t mimi_d, mimi_heads, mimi_layers, mimi_ffn, mimi_context;
  float mimi_max_period;
  /* SEANet decoder (modules/seanet.py:111-170) */
  int32_t seanet_dim, n_filters, n_ratios, ratios[8];
  int32_t kernel_size, res_kernel_size, last_kernel_size, compress;
  int32_t upsample_stride;          /* 16: encoder frame rate / frame rate (models/mimi.py:48-52) */
  /* sampling (default_parameters.py:3-10) */
  float temp;
  int32_t lsd_decode_steps;
  float noise_clamp;                /* < 0 or NaN: no clamp */
  float eos_threshold;
  /* runtime */
  int32_t precision;                /* ptts_precision: storage of weights + KV cache */
  int64_t kv_pool_tokens;           /* capacity of the paged FlowLM KV pool, in tokens */
  int32_t max_batch;                /* largest batch a ptts_batch may hold */
  int32_t reserved[8];
} ptts_config;

typedef struct ptts_ctx ptts_ctx;
typedef struct ptts_batch ptts_batch;

int32_t ptts_abi_version(void);
const char* ptts_last_error(void);
int32_t ptts_device_count(void);       /* 0 when no CUDA device/driver is usable */

/* ---- context + weights --------------------------------------------------------------------
 * Replaces TTSModel._from_pydantic_config_with_weights (models/tts_model.py:96-200): module tree
 * construction, the safetensors key walk and the conv weight transposes (:175-186). */
int32_t ptts_ctx_create(int32_t device, const ptts_config* cfg, ptts_ctx** out);
void ptts_ctx_destroy(ptts_ctx* ctx);
/* One checkpoint tensor in CHECKPOINT (PyTorch) layout under its checkpoint key, e.g.
 * "flow_lm.transformer.layers.0.self_attn.in_proj.weight".  Unknown names are ignored (return 1),
 * like the reference's skipped keys (:171-173,190-192). */
int32_t ptts_load_weight(ptts_ctx* ctx, const char* name, int32_t dtype, int32_t ndim,
                         const int64_t* shape, const void* data);
/* Re-pack (conv taps -> GEMM K axis, transposed convs -> polyphase), fold constants (time
 * embeddings of the LSD schedule, modules/mlp.py:53-74), convert to the storage precision, upload. */
int32_t ptts_finalize_weights(ptts_ctx* ctx);
/* After ptts_finalize_weights: the flow_lm.* / mimi.* keys that were loaded but that nothing on the path consumed, one
 * per line (empty string when every key was used).  The reference only counts such keys ("skipped",
 * models/tts_model.py:171-173,190-192); a loader uses this to validate its key map against a real checkpoint. */
const char* ptts_unused_weights(ptts_ctx* ctx);

/* Voice cloning.  ptts_has_voice_cloning: 1 when the loaded checkpoint carried the Mimi encoder
 * (mimi.encoder.*, mimi.encoder_transformer.*, mimi.downsample.*, flow_lm.speaker_proj_weight), like the reference's
 * `has_voice_cloning` (models/tts_model.py:77,145-151).  ptts_encode_audio turns a mono waveform at the model sample
 * rate into the conditioning that ptts_voice_create prefills: MimiModel.encode_to_latent (models/mimi.py:77-85: zero
 * padding to whole frames, SEANet encoder, encoder transformer, stride-16 downsample) followed by the speaker
 * projection (models/tts_model.py:271-276).  out_cond holds max_frames x d_model floats; *n_frames = ceil(n / 1920). */
int32_t ptts_has_voice_cloning(ptts_ctx* ctx);
int32_t ptts_encode_audio(ptts_ctx* ctx, const float* audio, int64_t n_samples, float* out_cond, int32_t max_frames,
                          int32_t* n_frames);

/* ---- voice prompt -------------------------------------------------------------------------
 * Replaces get_state_for_audio_prompt's prefill (models/tts_model.py:510-518): runs cond
 * [n_frames, d_model] through the backbone and keeps the KV pages as an immutable prefix that any
 * number of sequences share.  Returns the voice id (>= 0).
 * ptts_voice_destroy releases the prefix pages (the reference's state dict is garbage-collected).  Batch slots hold
 * a reference on their voice: destroying a voice that live slots still attend only retires the id, its pages return
 * to the pool when the last such slot is re-initialised or its batch destroyed. */
int32_t ptts_voice_create(ptts_ctx* ctx, const float* cond, int32_t n_frames);
int32_t ptts_voice_destroy(ptts_ctx* ctx, int32_t voice_id);
int32_t ptts_voice_length(ptts_ctx* ctx, int32_t voice_id);

/* ---- a batch of sequences generated in lock-step -------------------------------------------
 * Replaces the state plumbing of _generate_audio_stream_short_text (models/tts_model.py:372-384):
 * state copy, _expand_kv_cache, init_states(mimi), with per-sequence lengths instead of the
 * reference's single batch-1 `current_end`.  max_len[b] = voice + text + frames upper bound. */
int32_t ptts_batch_create(ptts_ctx* ctx, int32_t n_seq, const int32_t* voice_ids,
                          const int32_t* max_len, ptts_batch** out);
void ptts_batch_destroy(ptts_batch* batch);
/* Text prefill (tts_model.py:388-391): ids of sequence b are ids[offsets[b]..offsets[b+1]). */
int32_t ptts_batch_prefill_text(ptts_batch* batch, const int32_t* ids, const int32_t* offsets);
/* _warmup_mimi_decoder (tts_model.py:464-476): n_frames zero latents decoded and discarded. */
int32_t ptts_batch_warmup_mimi(ptts_batch* batch, int32_t n_frames);
/* One frame for every sequence = one replay of the per-frame CUDA graph (tts_model.py:404-426):
 * FlowLM step over the KV cache -> EOS logit -> flow head (all LSD steps) -> Mimi decode.
 *   noise      [n_seq, lsd_steps?1:1, latent_dim] raw N(0,1) draws (scaled by sqrt(temp) and clamped
 *              on the device, models/flow_lm.py:103-109), or NULL to use the device Philox stream;
 *   out_latent [n_seq, latent_dim] or NULL;  out_eos_logit [n_seq] or NULL;
 *   out_audio  [n_seq, 1920] or NULL (NULL also skips the device->host copy, not the decode). */
int32_t ptts_batch_step(ptts_batch* batch, const float* noise, float* out_latent,
                        float* out_eos_logit, float* out_audio);
/* Zero-copy variant of ptts_batch_step: the caller writes the noise into, and reads the results from, the
 * library's pinned staging buffers (noise [n_seq, latent_dim], latent [n_seq, latent_dim], eos_logit [n_seq],
 * audio [n_seq, 1920]); the pointers stay valid for the life of the batch, the contents until the next step. */
int32_t ptts_batch_host_buffers(ptts_batch* batch, float** noise, float** latent, float** eos_logit, float** audio);
int32_t ptts_batch_step_staged(ptts_batch* batch);
/* Teacher forcing for parity tests: overwrite the latent that the next step feeds back. */
int32_t ptts_batch_set_prev_latent(ptts_batch* batch, const float* latent);
/* Asynchronous staged steps (either mode): frames alternate between two sets of pinned staging buffers, so the
 * host can write the noise of frame t+1 and enqueue it (ptts_batch_step_staged_async returns the set it used,
 * = frame index & 1) while frame t is still running, then ptts_batch_staged_wait(set) before reading that frame's
 * latents / EOS logits / audio (the previous frame's in pipelined mode) from the set.  Enable with
 * ptts_batch_set_async_staging(batch, 1) before the first frame (after ptts_batch_set_pipelined, if that is used).
 * While it is on, ptts_batch_step / ptts_batch_step_staged (which only know set 0) return PTTS_ERR_STATE. */
int32_t ptts_batch_set_async_staging(ptts_batch* batch, int32_t on);
int32_t ptts_batch_host_buffers_set(ptts_batch* batch, int32_t set, float** noise, float** latent, float** eos_logit,
                                    float** audio);
int32_t ptts_batch_step_staged_async(ptts_batch* batch, int32_t* set_out);
int32_t ptts_batch_staged_wait(ptts_batch* batch, int32_t set);

/* Same step without the host round trip: results stay on the device (used by bench `value`). */
int32_t ptts_batch_step_device(ptts_batch* batch);
/* Continuous batching (the reference decodes one utterance at a time, models/tts_model.py:346-361):
 * a slot whose utterance has ended is re-initialised for the next one while the other sequences keep decoding.
 * ptts_batch_reset_seq returns the slot's private KV pages, attaches `voice_id`'s prefix (max_len must fit the page
 * budget the batch was created with), rewinds length / BOS flag and restores the slot's Mimi streaming state to the
 * one right after ptts_batch_warmup_mimi (= init_states + _warmup_mimi_decoder, models/tts_model.py:378-383,464-476).
 * Follow it with ptts_batch_prefill_text where every other sequence has an empty token range.
 * ptts_batch_set_active(slot, 0) parks a slot: it is still computed with the batch but stops growing its KV cache
 * (stream-ordered, allowed in every mode).  In pipelined mode the frame graph launched after ptts_batch_reset_seq(s)
 * still decodes the PREVIOUS utterance's last latent for the slot (that frame's audio for the slot is to be ignored)
 * and the slot's Mimi state is restored right after it; the new utterance's audio starts one step later, as always in
 * pipelined mode.  A batch whose sequences all share one voice attends the shared prefix once (cascade attention); the first
 * slot that switches to another voice turns that off for the batch (the frame graphs are re-captured). */
int32_t ptts_batch_reset_seq(ptts_batch* batch, int32_t slot, int32_t voice_id, int32_t max_len);
/* the same for n slots with one stream synchronisation */
int32_t ptts_batch_reset_seqs(ptts_batch* batch, int32_t n, const int32_t* slots, const int32_t* voice_ids,
                              const int32_t* max_lens);
int32_t ptts_batch_set_active(ptts_batch* batch, int32_t slot, int32_t active);

/* Throughput mode.  FlowLM step t only needs latent t-1, and so does the Mimi decode of frame t-1, so the two
 * run as concurrent branches of one CUDA graph.  After ptts_batch_set_pipelined(batch, 1) (before the first
 * frame) every step returns latent t and EOS logit t but the AUDIO OF FRAME t-1 (zeros at t = 0);
 * ptts_batch_flush decodes the last frame.  Results are identical to the sequential mode, one frame later. */
int32_t ptts_batch_set_pipelined(ptts_batch* batch, int32_t on);
int32_t ptts_batch_flush(ptts_batch* batch, float* out_audio);
/* 16-bit PCM output (replaces StreamingWAVWriter.write_pcm_data's host conversion, data/audio.py:64-70): after
 * ptts_batch_set_pcm16(batch, 1) (before the first frame) the kernels that produce a frame's final samples also store
 * them as int16 = trunc(clip(v, -1, 1) * 32767), and every host step copies THOSE out instead of the fp32 samples
 * (half the device->host bytes).  The frame's PCM lands in the pinned buffer ptts_batch_host_pcm returns for the
 * staging set the step used (set 0 for ptts_batch_step / _step_staged / _flush), [n_seq, 1920] int16; pass
 * out_audio = NULL to ptts_batch_step / ptts_batch_flush while it is on. */
int32_t ptts_batch_set_pcm16(ptts_batch* batch, int32_t on);
int32_t ptts_batch_host_pcm(ptts_batch* batch, int32_t set, int16_t** pcm);
int32_t ptts_batch_seed(ptts_batch* batch, uint64_t seed);
int32_t ptts_batch_lengths(ptts_batch* batch, int32_t* out_len);

/* Mimi decode only (models/mimi.py:70-75 driven frame by frame as in tts_model.py:415-419):
 * latents [n_seq, n_frames, latent_dim] -> audio [n_seq, n_frames*1920], continuing the batch's
 * Mimi streaming state.  audio may be NULL (device-resident timing).  With audio set, the waveforms are copied out in
 * chunks of 16 frames while later frames are still being decoded (any host memory; no extra pass at the end). */
int32_t ptts_batch_mimi_decode(ptts_batch* batch, const float* latents, int32_t n_frames, float* audio);

/* ---- timing + introspection -------------------------------------------------------------- */
int32_t ptts_sync(ptts_ctx* ctx);
/* CUDA events on the library's own stream: begin/end bracket a region, elapsed in milliseconds. */
int32_t ptts_timer_begin(ptts_ctx* ctx);
int32_t ptts_timer_end(ptts_ctx* ctx, float* elapsed_ms);
/* Kernel launches issued by this library since the counter was last reset (graph replays count
 * their kernel nodes). */
int64_t ptts_launch_count(ptts_ctx* ctx, int32_t reset);
/* One eager (non-graph) frame with every kernel launch bracketed by CUDA events on the library's
 * stream.  *report points at "kernel:tag,launches,ms,flops,bytes\n" lines (algorithmic flops/bytes of the
 * launches, sorted by time) valid until the next call.  Advances the batch by one frame. */
int32_t ptts_batch_profile_step(ptts_batch* batch, const char** report);
/* In-graph timing of the frame's sections (each captured as its own CUDA graph and replayed): ms[0] FlowLM
 * backbone, ms[1] EOS + flow head, ms[2] Mimi transformer, ms[3] SEANet decoder, ms[4] whole frame, ms[5] / ms[6] the
 * two Mimi sections with their persistent kernels capped at the SM share they get in the pipelined graph.  Perturbs
 * the batch's streaming state (profiling only).  Returns the number of entries written. */
int32_t ptts_batch_profile_sections(ptts_batch* batch, float* ms, int32_t cap);
/* Write an L2-sized scratch buffer (flushes L2 between timed iterations). */
int32_t ptts_flush_l2(ptts_ctx* ctx);
/* Stand-alone entry to the multi-tap linear operator, for kernel-level parity tests:
 * Y[b,t,n] = sum_{j<taps} sum_c A[b,t+j,c] * W[n, j*C+c] (+bias[n]); path 0 auto, 1 SIMT tile,
 * 2 small-M GEMV, 3 tcgen05 (bf16 storage only). */
int32_t ptts_debug_linear(ptts_ctx* ctx, int32_t path, int32_t n_b, int32_t n_t, int32_t taps,
                          int32_t c_in, int32_t n_out, const float* a, const float* w,
                          const float* bias, float* y);

/* Debugging aid (PTTS_ATTN_DBG=1 in the environment): {earliest CTA start, latest CTA end} in globaltimer nanoseconds of
 * the decode-attention launch of every FlowLM layer since the previous call (the stamps are reset by the call), i.e.
 * how long the HBM-bound attention really takes INSIDE a pipelined frame graph, next to the Mimi branch.  out holds
 * 2 * max_layers values; returns the number of layers written, 0 when the switch is off. */
int32_t ptts_debug_attention_stamps(ptts_ctx* ctx, uint64_t* out, int32_t max_layers);

/* Stand-alone run of the cluster chain kernel (one launch for a dependent chain of small-M GEMMs with fused
 * LayerNorm; csrc/chain_tc.cu) on a miniature flow head, for kernel-level parity tests:
 *   sy = silu(a0 W0^T + b0); ada = sy Wa^T + ba = shift | scale | gate; x1 = a1 Wi^T + bi;
 *   h = LN(x1; lnw, lnb)(1 + scale) + shift; u = silu(h W1^T + b1); x1 += gate * (u W2^T + b2); h = LN(x1);
 *   lat = lat_in + out_scale (h Wf^T + bf).
 * a0 [M, K0], a1 [M, 64], W0 [D, K0], Wa [3D, D], Wi [D, 64], W1 / W2 [D, D], Wf [32, D]; operands are rounded to
 * bf16 on upload.  Returns the cluster size used (8 or 16). */
int32_t ptts_debug_chain(ptts_ctx* ctx, int32_t M, int32_t D, int32_t K0, const float* a0, const float* a1, const float* w0,
                         const float* b0, const float* wa, const float* ba, const float* wi, const float* bi,
                         const float* lnw, const float* lnb, const float* w1, const float* b1, const float* w2,
                         const float* b2, const float* wf, const float* bf, const float* lat_in, float out_scale,
                         float* out_x1, float* out_h, float* out_lat, float* out_ada);

/* Kernel-level benchmark of the tcgen05 multi-tap GEMM on synthetic bf16 operands (L2 flushed between
 * launches): median microseconds over `reps`.  force = {N tile, ring stages, split-K, persistent} or NULL for
 * the planner's choice (returned in chosen[4]); epi = number of bf16 outputs (0: one fp32 output), +4 adds a
 * bf16 residual read.  reps < 0: |reps| warm back-to-back launches on the stream, average per launch;
 * reps <= -1000: the same |reps|-1000 launches replayed from a captured CUDA graph (the in-frame cost). */
int32_t ptts_debug_gemm_bench(ptts_ctx* ctx, int32_t n_b, int32_t n_t, int32_t taps, int32_t c_in, int32_t n_out,
                              int32_t epi, const int32_t* force, int32_t reps, float* us, int32_t* chosen);

#ifdef __cplusplus
}
#endif
#endif /* PTTS_H_ */
