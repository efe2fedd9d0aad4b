// libptts_b200 engine: context, weight re-packing, paged KV pool, voices, lock-step batches, per-frame
// CUDA graphs, and the extern "C" surface declared in include/ptts.h.
//
// Host-side structure replaced (reference, /root/reference/pocket_tts_mlx/):
//   models/tts_model.py:96-200   module tree + weight walk        -> Ctx::finalize (repack + upload)
//   models/tts_model.py:484-518  voice prefill                    -> voice_create (immutable KV pages)
//   models/tts_model.py:363-428  per-chunk state + frame loop     -> Batch (+ step graph)
//   modules/stateful_module.py   dict-of-dicts streaming state    -> device buffers owned by Batch
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/ptts.h"
#include "chain_tc.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

using namespace ptts;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(PTTS_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                  cudaGetErrorString(e_));                                                         \
  } while (0)

#define RET(call)                      \
  do {                                 \
    int r_ = (call);                   \
    if (r_ < 0) return r_;             \
  } while (0)

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
float h2f(uint16_t h) {   // IEEE half -> float
  uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023, u;
  if (e == 0) {
    if (m == 0) u = s << 31;
    else { e = 1; while (!(m & 1024)) { m <<= 1; --e; } m &= 1023; u = (s << 31) | ((e + 112) << 23) | (m << 13); }
  } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
  else u = (s << 31) | ((e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct LinW {
  void* w = nullptr;        // [N][K] storage precision
  __nv_bfloat16* w16 = nullptr;   // bf16 copy for the tensor-core path (== w in bf16 mode)
  float* bias = nullptr;    // [N] fp32 or null
  int bf16 = 0, N = 0, K = 0;
};

struct Voice {
  int len = 0;
  std::vector<int> pages;
  bool alive = false;
  // Live batch slots attached to this voice share its full prefix pages: ptts_voice_destroy on a voice that is still
  // in use only marks it (doomed); the pages go back to the pool when the last slot lets go.
  int refs = 0;
  bool doomed = false;
};

struct FlowWork {
  int cap = 0;
  bool tc = false;            // bf16 operands + tcgen05 GEMMs (M > 16 rows in bf16 mode)
  float *x = nullptr, *h = nullptr, *qkv = nullptr, *qrot = nullptr, *att = nullptr, *ff = nullptr;
  __nv_bfloat16 *h16 = nullptr, *att16 = nullptr, *ff16 = nullptr;   // bf16 operands for the tensor-core path
  int plan_M = 0;
  std::vector<TcGemm> plans;  // 4 per layer: qkv, out, ff1, ff2
  float *ws_out = nullptr, *ws_ff2 = nullptr;   // split-K partial planes [8][M][D] of out-proj / ffn2
  float* attn_part = nullptr;                   // split-KV attention partials [M][H][8][66] (decode at small batch)
  // cascade attention (decode of a batch whose sequences all share one voice prefix)
  int prefix_len = 0; int* d_prefix_pages = nullptr; float* prefix_part = nullptr;
  int* pflags = nullptr; int pflags_stride = 0;     // folded cascade: [n_layers][1 + tiles] claim counter + tile flags
  int pend_n = 0;             // planes of the last ffn2 still to be added to x (consumed by the next norm)
  float* rope_cs = nullptr;   // [M][64] cos | sin of each row's position (fused RoPE epilogue of the qkv GEMM)
  // set by the prefill callers around flow_layers: row ranges / start positions per sequence (device arrays), which
  // let whole chunks go through the tensor-core prefill attention instead of the per-row decode kernel
  const int* seq_row0 = nullptr; const int* seq_pos0 = nullptr; int n_seq = 0, max_rows_per_seq = 0;
  // interleaved pipelined frame: called right before / right after the attention kernels of layer i are launched
  std::function<void(int)> pre_attn, post_attn;
};

}  // namespace

struct ptts_batch;
struct ptts_ctx {
  int device = 0;
  ptts_config cfg{};
  cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2 carries the Mimi branch of the pipelined frame graph
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_attn[16] = {}, ev_slice[16] = {};     // interleaved pipelined frame: attention i launched / Mimi slice i done
  bool finalized = false;
  bool bf16 = true;
  bool force_simt = false;     // PTTS_FORCE_SIMT=1: keep the CUDA-core GEMMs (A/B comparisons)
  std::unordered_map<std::string, HostTensor> host;
  std::unordered_set<std::string> used;      // checkpoint keys finalize() consumed
  std::string unused_report;                  // loaded flow_lm.* / mimi.* keys nothing consumed, one per line
  std::vector<void*> allocs;

  struct FlowLayer { float *ln1w, *ln1b, *ln2w, *ln2b; LinW qkv, out, ff1, ff2; };
  std::vector<FlowLayer> fl;
  void* embed = nullptr;
  float *w_in = nullptr, *bos = nullptr, *emb_std = nullptr, *emb_mean = nullptr;
  float *outn_w = nullptr, *outn_b = nullptr, *eos_w = nullptr, *eos_b = nullptr;
  LinW cond, ada_all, in_proj, fin;
  __nv_bfloat16* in_proj_pad = nullptr;     // [flow_dim][64] bf16: input_proj with K zero-padded to one 64-wide k-block (chain kernel)
  std::vector<float*> cond_bias_step;
  struct ResBlk { float *lnw, *lnb; LinW m1, m2; };
  std::vector<ResBlk> rb;
  int n_ada = 0;
  float *wq = nullptr, *wu = nullptr;
  struct MimiLayer { float *ln1w, *ln1b, *ln2w, *ln2b, *ls1, *ls2; LinW qkv, out, ff1, ff2; };
  std::vector<MimiLayer> ml;
  LinW conv0;
  struct Stage { LinW ct, r3, r1; int stride, c_in, c_out, hidden; };
  std::vector<Stage> stages;
  float *fin_w = nullptr, *fin_b = nullptr;
  int fin_taps = 0, fin_c = 0;
  float *freqs_flow = nullptr, *freqs_mimi = nullptr;
  // voice cloning (Mimi encode side; loaded when the checkpoint carries mimi.encoder.*), fp32 weights
  bool has_encoder = false;
  float *enc0_w = nullptr, *enc0_b = nullptr;              // first conv: [n_filters][k], 1 input channel
  struct EncStage { int stride, c_in; LinW r3, r1, down; };
  std::vector<EncStage> enc_stages;
  LinW enc_last, enc_down, speaker_proj;
  std::vector<MimiLayer> el;

  void* pool = nullptr;
  long long n_pages = 0, page_stride = 0, layer_stride = 0;
  CUtensorMap kv_tmap[2];              // bf16 pool as [page][k|v][head][slot][64]: whole-page and 8-slot boxes (decode attention)
  bool kv_tmap_ok = false;
  std::vector<int> free_pages;
  std::vector<Voice> voices;
  FlowWork prefill_work;
  std::vector<ptts_batch*> parked;     // destroyed batches whose device arenas (and graphs) are recycled
  void* l2_scratch = nullptr;
  size_t l2_bytes = 0;
  std::string prof_names;
  int prio_hi = 0, prio_lo = 0;

  int dalloc(void** p, size_t bytes) {
    CU(cudaMalloc(p, bytes ? bytes : 16));
    allocs.push_back(*p);
    return 0;
  }
};

namespace {

using Ctx = ptts_ctx;

// ---- weight staging ---------------------------------------------------------------------------------------
const HostTensor* find(Ctx& c, const std::string& name) {
  auto it = c.host.find(name);
  if (it == c.host.end()) return nullptr;
  c.used.insert(name);
  return &it->second;
}

int need(Ctx& c, const std::string& name, std::initializer_list<int64_t> shape, const HostTensor** out) {
  const HostTensor* t = find(c, name);
  if (!t) return fail(PTTS_ERR_MISSING, "checkpoint tensor '%s' was not loaded", name.c_str());
  std::vector<int64_t> want(shape);
  if (t->shape != want) {
    std::string got, exp;
    for (auto s : t->shape) got += std::to_string(s) + ",";
    for (auto s : want) exp += std::to_string(s) + ",";
    return fail(PTTS_ERR_INVALID, "tensor '%s' has shape [%s] but [%s] is required", name.c_str(), got.c_str(),
                exp.c_str());
  }
  *out = t;
  return 0;
}

int upload_f32(Ctx& c, const float* src, size_t n, float** dst) {
  RET(c.dalloc((void**)dst, n * sizeof(float)));
  CU(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int upload_vec(Ctx& c, const std::string& name, int64_t n, float** dst) {
  const HostTensor* t;
  RET(need(c, name, {n}, &t));
  return upload_f32(c, t->data.data(), (size_t)n, dst);
}

// matrix [N][K] in storage precision (+ bf16 copy when storage is fp32 is NOT made: fp32 mode is SIMT-only)
bool g_upload_fp32 = false;   // set while the voice-cloning encoder is loaded: it always runs in fp32
int upload_mat(Ctx& c, const std::vector<float>& w, int N, int K, LinW* out) {
  const bool bf = c.bf16 && !g_upload_fp32;
  out->N = N; out->K = K; out->bf16 = bf ? 1 : 0;
  if (bf) {
    std::vector<uint16_t> h((size_t)N * K);
    for (size_t i = 0; i < h.size(); ++i) h[i] = f2bf(w[i]);
    RET(c.dalloc(&out->w, h.size() * 2));
    CU(cudaMemcpy(out->w, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    out->w16 = (__nv_bfloat16*)out->w;
  } else {
    RET(c.dalloc(&out->w, w.size() * 4));
    CU(cudaMemcpy(out->w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  }
  return 0;
}

int load_linear(Ctx& c, const std::string& prefix, int N, int K, bool bias, LinW* out) {
  const HostTensor* t;
  RET(need(c, prefix + ".weight", {N, K}, &t));
  RET(upload_mat(c, t->data, N, K, out));
  if (bias) RET(upload_vec(c, prefix + ".bias", N, &out->bias));
  return 0;
}

// Conv1d checkpoint weight (out,in,k) -> [out][j*in + c]
int load_conv(Ctx& c, const std::string& prefix, int c_out, int c_in, int k, LinW* out) {
  const HostTensor* t;
  RET(need(c, prefix + ".weight", {c_out, c_in, k}, &t));
  std::vector<float> w((size_t)c_out * k * c_in);
  for (int o = 0; o < c_out; ++o)
    for (int ci = 0; ci < c_in; ++ci)
      for (int j = 0; j < k; ++j)
        w[((size_t)o * k + j) * c_in + ci] = t->data[((size_t)o * c_in + ci) * k + j];
  RET(upload_mat(c, w, c_out, k * c_in, out));
  return upload_vec(c, prefix + ".bias", c_out, &out->bias);
}

// ConvTranspose1d checkpoint weight (in,out,2s), stride s -> polyphase [r*out+co][tap*in + ci]:
// tap 0 multiplies the previous input column (kernel index r+s), tap 1 the current one (kernel index r).
int load_convtr(Ctx& c, const std::string& prefix, int c_in, int c_out, int s, LinW* out) {
  const HostTensor *t, *b;
  RET(need(c, prefix + ".weight", {c_in, c_out, 2 * s}, &t));
  RET(need(c, prefix + ".bias", {c_out}, &b));
  const int N = s * c_out, K = 2 * c_in;
  std::vector<float> w((size_t)N * K);
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < c_out; ++co)
      for (int ci = 0; ci < c_in; ++ci) {
        const size_t n = (size_t)r * c_out + co;
        w[n * K + ci] = t->data[((size_t)ci * c_out + co) * 2 * s + r + s];
        w[n * K + c_in + ci] = t->data[((size_t)ci * c_out + co) * 2 * s + r];
      }
  RET(upload_mat(c, w, N, K, out));
  std::vector<float> bias((size_t)N);
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < c_out; ++co) bias[(size_t)r * c_out + co] = b->data[co];
  return upload_f32(c, bias.data(), bias.size(), &out->bias);
}

// time-embedding constant of one LSD step: (emb(s,0)+emb(t,1))/2, modules/mlp.py:53-74,161-163
int time_embedding(Ctx& c, int j, double tau, std::vector<double>* out) {
  const int fd = c.cfg.flow_dim;
  const std::string p = "flow_lm.flow_net.time_embed." + std::to_string(j) + ".mlp";
  const HostTensor *w1, *b1, *w2, *b2, *al;
  RET(need(c, p + ".0.weight", {fd, 256}, &w1));
  RET(need(c, p + ".0.bias", {fd}, &b1));
  RET(need(c, p + ".2.weight", {fd, fd}, &w2));
  RET(need(c, p + ".2.bias", {fd}, &b2));
  RET(need(c, p + ".3.alpha", {fd}, &al));
  std::vector<double> e(256), z1(fd), z2(fd);
  for (int i = 0; i < 128; ++i) {
    const double f = (double)(float)std::exp(-std::log(10000.0) * i / 128.0);
    const double arg = (double)(float)((float)tau * (float)f);
    e[i] = std::cos(arg);
    e[128 + i] = std::sin(arg);
  }
  for (int o = 0; o < fd; ++o) {
    double a = b1->data[o];
    for (int i = 0; i < 256; ++i) a += (double)w1->data[(size_t)o * 256 + i] * e[i];
    z1[o] = a / (1.0 + std::exp(-a));
  }
  double mean = 0;
  for (int o = 0; o < fd; ++o) {
    double a = b2->data[o];
    for (int i = 0; i < fd; ++i) a += (double)w2->data[(size_t)o * fd + i] * z1[i];
    z2[o] = a;
    mean += a;
  }
  mean /= fd;
  double var = 0;
  for (int o = 0; o < fd; ++o) var += (z2[o] - mean) * (z2[o] - mean);
  var /= (fd - 1);
  out->resize(fd);
  for (int o = 0; o < fd; ++o) (*out)[o] = z2[o] * ((double)al->data[o] / std::sqrt(1e-5 + var));
  return 0;
}

int finalize(Ctx& c) {
  const ptts_config& g = c.cfg;
  const int D = g.d_model, FF = g.ffn_dim, L = g.latent_dim, fd = g.flow_dim;
  // FlowLM
  c.fl.resize(g.n_layers);
  for (int i = 0; i < g.n_layers; ++i) {
    const std::string p = "flow_lm.transformer.layers." + std::to_string(i);
    auto& l = c.fl[i];
    RET(upload_vec(c, p + ".norm1.weight", D, &l.ln1w));
    RET(upload_vec(c, p + ".norm1.bias", D, &l.ln1b));
    RET(upload_vec(c, p + ".norm2.weight", D, &l.ln2w));
    RET(upload_vec(c, p + ".norm2.bias", D, &l.ln2b));
    RET(load_linear(c, p + ".self_attn.in_proj", 3 * D, D, false, &l.qkv));
    RET(load_linear(c, p + ".self_attn.out_proj", D, D, false, &l.out));
    RET(load_linear(c, p + ".linear1", FF, D, false, &l.ff1));
    RET(load_linear(c, p + ".linear2", D, FF, false, &l.ff2));
  }
  {
    const HostTensor* t;
    RET(need(c, "flow_lm.conditioner.embed.weight", {g.n_bins + 1, D}, &t));
    LinW tmp;
    RET(upload_mat(c, t->data, g.n_bins + 1, D, &tmp));
    c.embed = tmp.w;
    RET(need(c, "flow_lm.input_linear.weight", {D, L}, &t));
    std::vector<float> wt((size_t)D * L);                   // stored transposed [L][D] (coalesced in the kernel)
    for (int n = 0; n < D; ++n)
      for (int k = 0; k < L; ++k) wt[(size_t)k * D + n] = t->data[(size_t)n * L + k];
    RET(upload_f32(c, wt.data(), wt.size(), &c.w_in));
  }
  RET(upload_vec(c, "flow_lm.bos_emb", L, &c.bos));
  RET(upload_vec(c, "flow_lm.emb_std", L, &c.emb_std));
  RET(upload_vec(c, "flow_lm.emb_mean", L, &c.emb_mean));
  RET(upload_vec(c, "flow_lm.out_norm.weight", D, &c.outn_w));
  RET(upload_vec(c, "flow_lm.out_norm.bias", D, &c.outn_b));
  {
    const HostTensor* t;
    RET(need(c, "flow_lm.out_eos.weight", {1, D}, &t));
    RET(upload_f32(c, t->data.data(), D, &c.eos_w));
    RET(upload_vec(c, "flow_lm.out_eos.bias", 1, &c.eos_b));
  }
  // flow head
  const std::string fp = "flow_lm.flow_net";
  RET(load_linear(c, fp + ".cond_embed", fd, D, true, &c.cond));
  RET(load_linear(c, fp + ".input_proj", fd, L, true, &c.in_proj));
  if (c.bf16 && L <= 64) {
    const HostTensor* t;
    RET(need(c, fp + ".input_proj.weight", {fd, L}, &t));
    std::vector<uint16_t> h((size_t)fd * 64, 0);
    for (int n = 0; n < fd; ++n)
      for (int k = 0; k < L; ++k) h[(size_t)n * 64 + k] = f2bf(t->data[(size_t)n * L + k]);
    RET(c.dalloc((void**)&c.in_proj_pad, h.size() * 2));
    CU(cudaMemcpy(c.in_proj_pad, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  RET(load_linear(c, fp + ".final_layer.linear", L, fd, true, &c.fin));
  c.rb.resize(g.flow_depth);
  c.n_ada = g.flow_depth * 3 * fd + 2 * fd;
  std::vector<float> ada_w((size_t)c.n_ada * fd), ada_b((size_t)c.n_ada);
  for (int i = 0; i < g.flow_depth; ++i) {
    const std::string p = fp + ".res_blocks." + std::to_string(i);
    RET(upload_vec(c, p + ".in_ln.weight", fd, &c.rb[i].lnw));
    RET(upload_vec(c, p + ".in_ln.bias", fd, &c.rb[i].lnb));
    RET(load_linear(c, p + ".mlp.0", fd, fd, true, &c.rb[i].m1));
    RET(load_linear(c, p + ".mlp.2", fd, fd, true, &c.rb[i].m2));
    const HostTensor *w, *b;
    RET(need(c, p + ".adaLN_modulation.1.weight", {3 * fd, fd}, &w));
    RET(need(c, p + ".adaLN_modulation.1.bias", {3 * fd}, &b));
    memcpy(&ada_w[(size_t)i * 3 * fd * fd], w->data.data(), w->data.size() * 4);
    memcpy(&ada_b[(size_t)i * 3 * fd], b->data.data(), b->data.size() * 4);
  }
  {
    const HostTensor *w, *b;
    RET(need(c, fp + ".final_layer.adaLN_modulation.1.weight", {2 * fd, fd}, &w));
    RET(need(c, fp + ".final_layer.adaLN_modulation.1.bias", {2 * fd}, &b));
    memcpy(&ada_w[(size_t)g.flow_depth * 3 * fd * fd], w->data.data(), w->data.size() * 4);
    memcpy(&ada_b[(size_t)g.flow_depth * 3 * fd], b->data.data(), b->data.size() * 4);
    RET(upload_mat(c, ada_w, c.n_ada, fd, &c.ada_all));
    RET(upload_f32(c, ada_b.data(), ada_b.size(), &c.ada_all.bias));
  }
  {
    const HostTensor* cb;
    RET(need(c, fp + ".cond_embed.bias", {fd}, &cb));
    const int n = g.lsd_decode_steps;
    c.cond_bias_step.resize(n);
    for (int i = 0; i < n; ++i) {
      std::vector<double> e0, e1;
      RET(time_embedding(c, 0, (double)i / n, &e0));
      RET(time_embedding(c, 1, (double)(i + 1) / n, &e1));
      std::vector<float> bias(fd);
      for (int o = 0; o < fd; ++o) bias[o] = (float)((e0[o] + e1[o]) * 0.5 + cb->data[o]);
      RET(upload_f32(c, bias.data(), fd, &c.cond_bias_step[i]));
    }
  }
  // Mimi
  const int MD = g.mimi_d, SD = g.seanet_dim, S = g.upsample_stride;
  {
    const HostTensor* t;
    RET(need(c, "mimi.quantizer.output_proj.weight", {SD, L, 1}, &t));
    std::vector<float> wt((size_t)SD * L);                  // [L][SD]
    for (int ch = 0; ch < SD; ++ch)
      for (int k = 0; k < L; ++k) wt[(size_t)k * SD + ch] = t->data[(size_t)ch * L + k];
    RET(upload_f32(c, wt.data(), wt.size(), &c.wq));
    RET(need(c, "mimi.upsample.convtr.convtr.weight", {SD, 1, 2 * S}, &t));
    std::vector<float> ut((size_t)SD * 2 * S);              // [2S][SD]
    for (int ch = 0; ch < SD; ++ch)
      for (int k = 0; k < 2 * S; ++k) ut[(size_t)k * SD + ch] = t->data[(size_t)ch * 2 * S + k];
    RET(upload_f32(c, ut.data(), ut.size(), &c.wu));
  }
  c.ml.resize(g.mimi_layers);
  for (int i = 0; i < g.mimi_layers; ++i) {
    const std::string p = "mimi.decoder_transformer.transformer.layers." + std::to_string(i);
    auto& l = c.ml[i];
    RET(upload_vec(c, p + ".norm1.weight", MD, &l.ln1w));
    RET(upload_vec(c, p + ".norm1.bias", MD, &l.ln1b));
    RET(upload_vec(c, p + ".norm2.weight", MD, &l.ln2w));
    RET(upload_vec(c, p + ".norm2.bias", MD, &l.ln2b));
    RET(upload_vec(c, p + ".layer_scale_1.scale", MD, &l.ls1));
    RET(upload_vec(c, p + ".layer_scale_2.scale", MD, &l.ls2));
    RET(load_linear(c, p + ".self_attn.in_proj", 3 * MD, MD, false, &l.qkv));
    RET(load_linear(c, p + ".self_attn.out_proj", MD, MD, false, &l.out));
    RET(load_linear(c, p + ".linear1", g.mimi_ffn, MD, false, &l.ff1));
    RET(load_linear(c, p + ".linear2", MD, g.mimi_ffn, false, &l.ff2));
  }
  // SEANet decoder: model.0 conv, then per ratio [ELU, convtr, resblock], then [ELU, conv]
  int mult = 1 << g.n_ratios, idx = 0;
  RET(load_conv(c, "mimi.decoder.model.0.conv", mult * g.n_filters, SD, g.kernel_size, &c.conv0));
  idx = 1;
  c.stages.resize(g.n_ratios);
  for (int r = 0; r < g.n_ratios; ++r) {
    auto& st = c.stages[r];
    st.stride = g.ratios[r];
    st.c_in = mult * g.n_filters;
    st.c_out = st.c_in / 2;
    st.hidden = st.c_out / g.compress;
    idx += 1;
    RET(load_convtr(c, "mimi.decoder.model." + std::to_string(idx) + ".convtr", st.c_in, st.c_out, st.stride, &st.ct));
    idx += 1;
    const std::string rp = "mimi.decoder.model." + std::to_string(idx) + ".block.";
    RET(load_conv(c, rp + "1.conv", st.hidden, st.c_out, g.res_kernel_size, &st.r3));
    RET(load_conv(c, rp + "3.conv", st.c_out, st.hidden, 1, &st.r1));
    idx += 1;
    mult /= 2;
  }
  idx += 1;
  {
    const HostTensor *t, *b;
    const std::string p = "mimi.decoder.model." + std::to_string(idx) + ".conv";
    RET(need(c, p + ".weight", {1, g.n_filters, g.last_kernel_size}, &t));
    RET(need(c, p + ".bias", {1}, &b));
    c.fin_taps = g.last_kernel_size;
    c.fin_c = g.n_filters;
    std::vector<float> w((size_t)c.fin_taps * c.fin_c);
    for (int ci = 0; ci < c.fin_c; ++ci)
      for (int j = 0; j < c.fin_taps; ++j) w[(size_t)j * c.fin_c + ci] = t->data[(size_t)ci * c.fin_taps + j];
    RET(upload_f32(c, w.data(), w.size(), &c.fin_w));
    RET(upload_f32(c, b->data.data(), 1, &c.fin_b));
  }
  // RoPE frequency tables exactly as modules/rope.py:17-18 computes them (fp32 exp of an fp32 product)
  auto freqs = [&](float period, float** dst) -> int {
    float f[32];
    const float scale = (float)(-std::log((double)period) * 2.0 / kHeadDim);
    for (int i = 0; i < 32; ++i) f[i] = expf((float)i * scale);
    return upload_f32(c, f, 32, dst);
  };
  RET(freqs(g.max_period, &c.freqs_flow));
  RET(freqs(g.mimi_max_period, &c.freqs_mimi));
  // KV pool
  c.page_stride = 2LL * g.n_heads * kPageTokens * kHeadDim;
  c.n_pages = (g.kv_pool_tokens + kPageTokens - 1) / kPageTokens;
  c.layer_stride = c.n_pages * c.page_stride;
  const size_t esz = c.bf16 ? 2 : 4;
  RET(c.dalloc(&c.pool, (size_t)g.n_layers * c.layer_stride * esz));
  // every byte of the pool is finite from the start: the tensor-core decode attention multiplies masked keys' V rows
  // by an exact 0
  CU(cudaMemset(c.pool, 0, (size_t)g.n_layers * c.layer_stride * esz));
  c.kv_tmap_ok = false;
  if (c.bf16 && gemm_tc_available()) {
    const unsigned long long H = (unsigned long long)g.n_heads, pages = (unsigned long long)g.n_layers * (unsigned long long)c.n_pages;
    const unsigned long long dims[5] = {(unsigned long long)kHeadDim, (unsigned long long)kPageTokens, H, 2ull, pages};
    const unsigned long long str[4] = {(unsigned long long)kHeadDim * 2, (unsigned long long)kPageTokens * kHeadDim * 2,
                                       H * kPageTokens * kHeadDim * 2, (unsigned long long)c.page_stride * 2};
    const unsigned box_page[5] = {(unsigned)kHeadDim, (unsigned)kPageTokens, 1u, 2u, 1u};
    const unsigned box_8[5] = {(unsigned)kHeadDim, 8u, 1u, 1u, 1u};
    c.kv_tmap_ok = pages < (1ull << 31) && tc_encode_bf16(&c.kv_tmap[0], c.pool, 5, dims, str, box_page, 64) &&
                   tc_encode_bf16(&c.kv_tmap[1], c.pool, 5, dims, str, box_8, 64);
  }
  c.free_pages.resize(c.n_pages);
  for (long long i = 0; i < c.n_pages; ++i) c.free_pages[i] = (int)(c.n_pages - 1 - i);
  // ---- voice cloning: SEANet encoder, encoder transformer, downsample, speaker projection (fp32) ----
  if (find(c, "mimi.encoder.model.0.conv.weight") && find(c, "mimi.downsample.conv.conv.weight") &&
      find(c, "flow_lm.speaker_proj_weight")) {
    g_upload_fp32 = true;
    auto done = [&](int rc) { g_upload_fp32 = false; return rc; };
    const HostTensor *t0, *b0;
    if (int rc = need(c, "mimi.encoder.model.0.conv.weight", {g.n_filters, 1, g.kernel_size}, &t0)) return done(rc);
    if (int rc = need(c, "mimi.encoder.model.0.conv.bias", {g.n_filters}, &b0)) return done(rc);
    if (int rc = upload_f32(c, t0->data.data(), t0->data.size(), &c.enc0_w)) return done(rc);
    if (int rc = upload_f32(c, b0->data.data(), b0->data.size(), &c.enc0_b)) return done(rc);
    int emult = 1, eidx = 1;
    c.enc_stages.resize(g.n_ratios);
    for (int r = 0; r < g.n_ratios; ++r) {
      auto& st = c.enc_stages[r];
      st.stride = g.ratios[g.n_ratios - 1 - r];          // the encoder walks the ratios in reverse
      st.c_in = emult * g.n_filters;
      const std::string rp = "mimi.encoder.model." + std::to_string(eidx) + ".block.";
      if (int rc = load_conv(c, rp + "1.conv", st.c_in / g.compress, st.c_in, g.res_kernel_size, &st.r3)) return done(rc);
      if (int rc = load_conv(c, rp + "3.conv", st.c_in, st.c_in / g.compress, 1, &st.r1)) return done(rc);
      eidx += 2;
      if (int rc = load_conv(c, "mimi.encoder.model." + std::to_string(eidx) + ".conv", 2 * st.c_in, st.c_in, 2 * st.stride,
                             &st.down)) return done(rc);
      eidx += 1;
      emult *= 2;
    }
    eidx += 1;
    if (int rc = load_conv(c, "mimi.encoder.model." + std::to_string(eidx) + ".conv", SD, emult * g.n_filters,
                           g.last_kernel_size, &c.enc_last)) return done(rc);
    c.el.resize(g.mimi_layers);
    for (int i = 0; i < g.mimi_layers; ++i) {
      const std::string p = "mimi.encoder_transformer.transformer.layers." + std::to_string(i);
      auto& l = c.el[i];
      int rc = 0;
      rc = rc ? rc : upload_vec(c, p + ".norm1.weight", MD, &l.ln1w);
      rc = rc ? rc : upload_vec(c, p + ".norm1.bias", MD, &l.ln1b);
      rc = rc ? rc : upload_vec(c, p + ".norm2.weight", MD, &l.ln2w);
      rc = rc ? rc : upload_vec(c, p + ".norm2.bias", MD, &l.ln2b);
      rc = rc ? rc : upload_vec(c, p + ".layer_scale_1.scale", MD, &l.ls1);
      rc = rc ? rc : upload_vec(c, p + ".layer_scale_2.scale", MD, &l.ls2);
      rc = rc ? rc : load_linear(c, p + ".self_attn.in_proj", 3 * MD, MD, false, &l.qkv);
      rc = rc ? rc : load_linear(c, p + ".self_attn.out_proj", MD, MD, false, &l.out);
      rc = rc ? rc : load_linear(c, p + ".linear1", g.mimi_ffn, MD, false, &l.ff1);
      rc = rc ? rc : load_linear(c, p + ".linear2", MD, g.mimi_ffn, false, &l.ff2);
      if (rc) return done(rc);
    }
    {
      // downsample: Conv1d (out,in,2s) without bias -> [out][j*in + c]
      const HostTensor* t;
      const int s2 = 2 * g.upsample_stride;
      if (int rc = need(c, "mimi.downsample.conv.conv.weight", {SD, SD, s2}, &t)) return done(rc);
      std::vector<float> w((size_t)SD * s2 * SD);
      for (int o = 0; o < SD; ++o)
        for (int ci = 0; ci < SD; ++ci)
          for (int j = 0; j < s2; ++j) w[((size_t)o * s2 + j) * SD + ci] = t->data[((size_t)o * SD + ci) * s2 + j];
      if (int rc = upload_mat(c, w, SD, s2 * SD, &c.enc_down)) return done(rc);
    }
    {
      const HostTensor* t;
      if (int rc = need(c, "flow_lm.speaker_proj_weight", {D, SD}, &t)) return done(rc);
      if (int rc = upload_mat(c, t->data, D, SD, &c.speaker_proj)) return done(rc);
    }
    g_upload_fp32 = false;
    c.has_encoder = true;
  }
  {
    // keys that were offered but that no part of the path reads (the reference counts them as "skipped",
    // models/tts_model.py:171-173,190-192): kept so that a loader can check its key map against a real checkpoint.
    // When the encoder is not loaded as a whole, its keys are expected leftovers and still listed.
    std::vector<std::string> left;
    for (auto& kv : c.host) if (!c.used.count(kv.first)) left.push_back(kv.first);
    std::sort(left.begin(), left.end());
    for (auto& n : left) { c.unused_report += n; c.unused_report += '\n'; }
  }
  c.host.clear();
  c.used.clear();
  c.finalized = true;
  return 0;
}

// ---- operator dispatch -----------------------------------------------------------------------------------
void rows_norm(Ctx& c, const float* X, int M, int C, const float* w, const float* b, float eps, float* Y,
               const float* scale, const float* shift, long long mod_rs, __nv_bfloat16* Y16, const float* acc, int acc_n);
void run_linear(Ctx& c, const LinW& w, LinearParams p) {
  p.W = w.w; p.w_bf16 = w.bf16; p.N = w.N;
  if (!p.bias) p.bias = w.bias;
  if (p.out_scale == 0.f) p.out_scale = 1.f;
  if (linear_gemv_supported(p)) launch_linear_gemv(p, c.stream);
  else launch_linear_tile(p, c.stream);
}

// LayerNorm (+ optional AdaLN modulation) followed by a Linear on `M` rows: one GEMV launch with the norm fused
// into its row staging when the small-M path applies, otherwise the norm kernel writes `scratch` first.
void run_norm_linear(Ctx& c, const LinW& w, const float* X, int M, int C, const float* lw, const float* lb, float eps,
                     const float* scale, const float* shift, long long mod_rs, float* scratch, LinearParams p) {
  p.A = X; p.a_bs = 0; p.a_rs = C; p.nb = 1; p.T = M; p.taps = 1; p.C = C;
  p.W = w.w; p.w_bf16 = w.bf16; p.N = w.N;
  if (linear_gemv_supported(p)) {
    p.ln_on = 1; p.ln_w = lw; p.ln_b = lb; p.ln_eps = eps; p.ln_scale = scale; p.ln_shift = shift; p.ln_mod_rs = mod_rs;
    run_linear(c, w, p);
  } else {
    rows_norm(c, X, M, C, lw, lb, eps, scratch, scale, shift, mod_rs, nullptr, nullptr, 0);
    p.A = scratch;
    run_linear(c, w, p);
  }
}

LinearParams rows_linear(const float* A, int M, int C, float* Y, int N, const char* tag = nullptr) {
  LinearParams p{};
  p.tag = tag;
  p.A = A; p.a_bs = 0; p.a_rs = C; p.nb = 1; p.T = M; p.taps = 1; p.C = C;
  p.Y = Y; p.y_bs = 0; p.y_rs = N;
  p.out_scale = 1.f;
  return p;
}

void rows_norm(Ctx& c, const float* X, int M, int C, const float* w, const float* b, float eps, float* Y,
               const float* scale, const float* shift, long long mod_rs, __nv_bfloat16* Y16, const float* acc, int acc_n) {
  NormParams n{};
  n.Y16 = Y16;
  n.acc = acc; n.acc_n = acc_n; n.acc_stride = (long long)M * C;
  n.X = X; n.x_bs = 0; n.x_rs = C; n.nb = 1; n.T = M; n.C = C;
  n.w = w; n.b = b; n.eps = eps; n.scale = scale; n.shift = shift; n.mod_rs = mod_rs;
  n.Y = Y; n.y_bs = 0; n.y_rs = C;
  launch_layernorm(n, c.stream);
}

bool want_tc(Ctx& c, int M) {
  static const int min_rows = [] { const char* v = getenv("PTTS_TC_MIN_ROWS"); return v ? atoi(v) : 16; }();
  return c.bf16 && gemm_tc_available() && M >= min_rows && !c.force_simt;
}

void free_flow_work(FlowWork& w) {
  void* ptrs[] = {w.x, w.h, w.qkv, w.qrot, w.att, w.ff, w.h16, w.att16, w.ff16, w.ws_out, w.ws_ff2, w.attn_part,
                  w.d_prefix_pages, w.prefix_part, w.pflags, w.rope_cs};
  for (void* p : ptrs) if (p) cudaFree(p);
  w = FlowWork{};
}

int alloc_flow_work(Ctx& c, FlowWork& w, int M) {
  const bool tc = want_tc(c, M);
  if (M <= w.cap && tc == w.tc) return 0;
  free_flow_work(w);
  const size_t D = c.cfg.d_model, FF = c.cfg.ffn_dim;
  CU(cudaMalloc((void**)&w.x, M * D * 4));
  CU(cudaMalloc((void**)&w.qkv, M * 3 * D * 4));
  CU(cudaMalloc((void**)&w.qrot, M * D * 4));
  if (M * c.cfg.n_heads < 148) CU(cudaMalloc((void**)&w.attn_part, (size_t)M * c.cfg.n_heads * 8 * 66 * 4));
  if (tc) {
    CU(cudaMalloc((void**)&w.h16, M * D * 2));
    CU(cudaMalloc((void**)&w.att16, M * D * 2));
    CU(cudaMalloc((void**)&w.ff16, M * FF * 2));
    CU(cudaMalloc((void**)&w.ws_out, 8 * M * D * 4));
    CU(cudaMalloc((void**)&w.ws_ff2, 8 * M * D * 4));
    CU(cudaMalloc((void**)&w.rope_cs, (size_t)M * 64 * 4));
  } else {
    CU(cudaMalloc((void**)&w.h, M * D * 4));
    CU(cudaMalloc((void**)&w.att, M * D * 4));
    CU(cudaMalloc((void**)&w.ff, M * FF * 4));
  }
  w.cap = M;
  w.tc = tc;
  w.plan_M = 0;
  return 0;
}

// L2 policy of a GEMM's operand loads by call site.  PTTS_L2_W / PTTS_L2_A: comma-separated "prefix=policy" (policy 0
// default, 1 evict-first, 2 evict-last), e.g. PTTS_L2_W="flow.=1,head.=2".
int l2_policy_for(const char* env, const char* dflt, const char* tag) {
  const char* v = getenv(env);
  std::string spec = v ? v : dflt;
  int best = 0; size_t best_len = 0;
  size_t pos = 0;
  while (pos < spec.size()) {
    size_t end = spec.find(',', pos);
    if (end == std::string::npos) end = spec.size();
    const std::string item = spec.substr(pos, end - pos);
    const size_t eq = item.find('=');
    if (eq != std::string::npos) {
      const std::string pre = item.substr(0, eq);
      if (tag && strncmp(tag, pre.c_str(), pre.size()) == 0 && pre.size() >= best_len) { best = atoi(item.c_str() + eq + 1); best_len = pre.size(); }
    }
    pos = end + 1;
  }
  return best;
}

bool plan_tc(TcGemm* g, const __nv_bfloat16* a, long long a_bs, long long a_rs, int nb, int T, int taps, int C,
             const LinW& w, const char* tag, int max_splits = 1, int n_bf16_out = 0) {
  if (!w.w16 || w.K != taps * C) return false;
  if (!gemm_tc_plan(g, a, a_bs, a_rs, nb, T, taps, C, w.w16, w.N, tag, max_splits, n_bf16_out)) return false;
  g->e.bias = w.bias;
  // defaults: the FlowLM layer weights of a decode step (144 MB, read once per frame) and the big Mimi / SEANet
  // activations (read once by their consumer) are evict-first
  g->w_policy = l2_policy_for("PTTS_L2_W", "flow.=1", tag);
  g->a_policy = l2_policy_for("PTTS_L2_A", "sn.=1,mimi.=1", tag);
  if ((long long)nb * T > 2048 && g->w_policy == 1) g->w_policy = 0;      // prefill: the weights are re-read by every M tile
  return true;
}

int build_flow_plans(Ctx& c, FlowWork& w, int M) {
  if (w.plan_M == M) return 0;
  const int D = c.cfg.d_model, FF = c.cfg.ffn_dim;
  w.plans.assign((size_t)c.cfg.n_layers * 4, TcGemm{});
  for (int i = 0; i < c.cfg.n_layers; ++i) {
    auto& l = c.fl[i];
    TcGemm* g = &w.plans[(size_t)i * 4];
    // out-proj and ffn2 only feed the residual stream, so they may run split-K: the planes are summed into x
    // by the next LayerNorm (deterministic order) instead of an epilogue
    bool ok = plan_tc(&g[0], w.h16, 0, D, 1, M, 1, D, l.qkv, "flow.qkv") &&
              plan_tc(&g[1], w.att16, 0, D, 1, M, 1, D, l.out, "flow.out", 8) &&
              plan_tc(&g[2], w.h16, 0, D, 1, M, 1, D, l.ff1, "flow.ff1", 1, 1) &&
              plan_tc(&g[3], w.ff16, 0, FF, 1, M, 1, FF, l.ff2, "flow.ff2", 8);
    if (!ok) return fail(PTTS_ERR_CUDA, "tcgen05 plan failed for FlowLM layer %d (M=%d)", i, M);
    g[0].e.y32 = w.qkv; g[0].e.y32_rs = 3 * D;
    g[1].e.res32 = w.x; g[1].e.res32_rs = D; g[1].e.y32 = w.x; g[1].e.y32_rs = D;
    g[1].split_ws = w.ws_out;
    g[2].e.act = ACT_GELU; g[2].e.y16 = w.ff16; g[2].e.y16_rs = FF;
    g[3].e.res32 = w.x; g[3].e.res32_rs = D; g[3].e.y32 = w.x; g[3].e.y32_rs = D;
    g[3].split_ws = w.ws_ff2;
    gemm_tc_bind_outputs(&g[2]);
  }
  // decode steps: every GEMM of the chain asks L2 for the weights of the next one (PTTS_TC_PREFETCH=0: off)
  static const bool tc_pf = [] { const char* v = getenv("PTTS_TC_PREFETCH"); return !(v && v[0] == '0'); }();
  if (tc_pf && M <= 2048) {
    auto wbytes = [](const LinW& l) { return (long long)l.N * l.K * 2; };
    for (int i = 0; i < c.cfg.n_layers; ++i) {
      auto& l = c.fl[i];
      TcGemm* g = &w.plans[(size_t)i * 4];
      g[0].pf_ptr = l.out.w16; g[0].pf_bytes = wbytes(l.out);
      g[1].pf_ptr = l.ff1.w16; g[1].pf_bytes = wbytes(l.ff1);
      g[2].pf_ptr = l.ff2.w16; g[2].pf_bytes = wbytes(l.ff2);
      if (i + 1 < c.cfg.n_layers) { g[3].pf_ptr = c.fl[i + 1].qkv.w16; g[3].pf_bytes = wbytes(c.fl[i + 1].qkv); }
    }
  }
  w.plan_M = M;
  return 0;
}

// The single-launch flow head (cluster chain kernel) is opt-in: measured at batch 256 it takes 199 us against 133 us for
// the chain of 23 launches it replaces (per step ~10 us of cluster barrier, multicast fetch and L2-latency-bound
// epilogue work against ~5.8 us per captured launch), see DESIGN.md.
bool chain_enabled() {
  const char* v = getenv("PTTS_CHAIN");
  return v && v[0] == '1' && chain_cluster_size() > 0;
}

// the 6 pre-LN layers over M rows that sit at (row_seq, row_pos) of their sequences
void flow_layers(Ctx& c, FlowWork& w, int M, const int* row_seq, const int* row_pos, const int* page_table,
                 int max_pages, long long total_keys) {
  const int D = c.cfg.d_model, FF = c.cfg.ffn_dim;
  if (w.tc) {
    int pend = 0;   // split-K planes of the previous ffn2 that the next norm must add to x
    // RoPE + KV append ride in the qkv GEMM's epilogue (PTTS_NO_ROPE_FUSE=1 keeps the separate kernel)
    static const bool fuse_rope = [] { const char* v = getenv("PTTS_NO_ROPE_FUSE"); return !(v && v[0] == '1'); }();
    if (fuse_rope) launch_rope_table(row_pos, c.freqs_flow, w.rope_cs, M, 1, c.stream);
    for (int i = 0; i < c.cfg.n_layers; ++i) {
      auto& l = c.fl[i];
      const TcGemm* g = &w.plans[(size_t)i * 4];
      rows_norm(c, w.x, M, D, l.ln1w, l.ln1b, 1e-5f, nullptr, nullptr, nullptr, 0, w.h16, w.ws_ff2, pend);
      if (fuse_rope) {
        TcGemm q = g[0];
        auto& e = q.e;
        e.y32 = nullptr;
        e.rope_cs = w.rope_cs; e.q_rot = w.qrot;
        e.kv_layer = reinterpret_cast<__nv_bfloat16*>(c.pool) + (long long)i * c.layer_stride;
        e.kv_row_seq = row_seq; e.kv_row_pos = row_pos; e.kv_page_table = page_table;
        e.kv_max_pages = max_pages; e.kv_heads = c.cfg.n_heads; e.kv_page_stride = c.page_stride;
        gemm_tc_launch(q, c.stream);
      } else {
        gemm_tc_launch(g[0], c.stream);
      }
      FlowAttnParams a{};
      a.qkv = w.qkv; a.q_rot = w.qrot; a.out16 = w.att16;
      a.pool = c.pool; a.kv_bf16 = c.bf16; a.layer_stride = c.layer_stride; a.page_stride = c.page_stride;
      a.layer = i; a.row_seq = row_seq; a.row_pos = row_pos; a.page_table = page_table; a.max_pages = max_pages;
      a.M = M; a.H = c.cfg.n_heads; a.freqs = c.freqs_flow; a.total_keys = total_keys;
      a.part = w.attn_part; a.splits = w.attn_part ? std::min(8, std::max(1, 296 / (M * c.cfg.n_heads))) : 1;
      a.kv_tmap = c.kv_tmap_ok ? c.kv_tmap : nullptr;
      a.tstamp = (!row_seq && i < 32) ? flow_attention_dbg_buffer() : nullptr;
      { static const int kvp = [] { const char* v = getenv("PTTS_KV_EVICT_FIRST"); return v ? atoi(v) : 1; }(); a.kv_evict_first = kvp; }
      if (!row_seq && w.prefix_len > 0 && a.splits == 1) {
        a.prefix_len = w.prefix_len; a.prefix_pages = w.d_prefix_pages; a.prefix_part = w.prefix_part;
      }
      bool folded = false;
      if (!fuse_rope) launch_flow_rope_append(a, c.stream);
      a.seq_row0 = w.seq_row0; a.seq_pos0 = w.seq_pos0; a.n_seq = w.n_seq; a.max_rows_per_seq = w.max_rows_per_seq;
      static const bool no_tc_prefill = [] { const char* v = getenv("PTTS_NO_TC_PREFILL_ATTN"); return v && v[0] == '1'; }();
      if (row_seq && !no_tc_prefill && launch_flow_prefill_attention(a, c.stream)) {
        // whole prefill chunks: FlashAttention-2 style mma.sync kernel (64 query rows per CTA)
      } else {
        if (w.pre_attn) w.pre_attn(i);
        // cascade: the shared-prefix partials come from the stream kernel itself when it runs (no launch of their own)
        // (opt-in, PTTS_FOLD=1, read when the step is recorded: it saves the launch but its tiles unbalance the CTAs'
        // item ranges, measured +0.6 % frame throughput with the attention launch 6 us longer -- DESIGN.md section 4b)
        const char* fold_env = getenv("PTTS_FOLD");
        if (a.prefix_len > 0 && w.pflags && fold_env && fold_env[0] == '1' && c.cfg.n_layers >= 2 && flow_attention_streams(a)) {
          a.pflags = w.pflags + (long long)i * w.pflags_stride;
          a.pflags_next = w.pflags + (long long)((i + 1) % c.cfg.n_layers) * w.pflags_stride;
          folded = true;
        }
        if (!folded) launch_flow_prefix_attention(a, c.stream);
        launch_flow_attention(a, c.stream);
        if (w.post_attn) w.post_attn(i);
      }
      gemm_tc_launch(g[1], c.stream);
      rows_norm(c, w.x, M, D, l.ln2w, l.ln2b, 1e-5f, nullptr, nullptr, nullptr, 0, w.h16, w.ws_out,
                g[1].splits > 1 ? g[1].splits : 0);
      gemm_tc_launch(g[2], c.stream);
      gemm_tc_launch(g[3], c.stream);
      pend = g[3].splits > 1 ? g[3].splits : 0;
    }
    w.pend_n = pend;
    return;
  }
  w.pend_n = 0;
  static const bool fuse_rope_gemv = [] { const char* v = getenv("PTTS_NO_ROPE_FUSE"); return !(v && v[0] == '1'); }();
  // The 144 MB of layer weights pass through once per frame and do not fit L2 next to anything else: they are loaded with
  // the streaming (evict-first) hint, which leaves the flow head's and the Mimi decoder's weights (40 MB) resident from
  // frame to frame: batch-1 frame 0.488 -> 0.400 ms.  Keeping the first layer or two resident as well is slower (0.418 /
  // 0.479), and so is asking L2 for the next launch's weights ahead of time (0.454).  PTTS_W_STREAM=0: plain loads.
  static const int w_stream = [] { const char* v = getenv("PTTS_W_STREAM"); return (v && v[0] == '0') ? 0 : 1; }();
  for (int i = 0; i < c.cfg.n_layers; ++i) {
    auto& l = c.fl[i];
    LinearParams q = rows_linear(w.h, M, D, w.qkv, 3 * D, "flow.qkv");
    q.w_stream = w_stream;
    bool roped = false;
    {
      // batch <= 4 (the latency path): RoPE + KV append ride in the GEMV epilogue, one launch less per layer
      LinearParams probe = q;
      probe.A = w.x; probe.a_bs = 0; probe.a_rs = D; probe.nb = 1; probe.T = M; probe.taps = 1; probe.C = D; probe.N = l.qkv.N;
      if (fuse_rope_gemv && linear_gemv_rope_supported(probe)) {
        roped = true;
        q.rope_on = 1; q.q_rot = w.qrot; q.kv_bf16 = c.bf16;
        q.kv_layer = (char*)c.pool + (size_t)i * c.layer_stride * (c.bf16 ? 2 : 4);
        q.kv_row_seq = row_seq; q.kv_row_pos = row_pos; q.kv_page_table = page_table; q.kv_max_pages = max_pages;
        q.kv_heads = c.cfg.n_heads; q.kv_page_stride = c.page_stride; q.rope_freqs = c.freqs_flow;
      }
    }
    run_norm_linear(c, l.qkv, w.x, M, D, l.ln1w, l.ln1b, 1e-5f, nullptr, nullptr, 0, w.h, q);
    FlowAttnParams a{};
    a.qkv = w.qkv; a.q_rot = w.qrot; a.out = w.att;
    a.pool = c.pool; a.kv_bf16 = c.bf16; a.layer_stride = c.layer_stride; a.page_stride = c.page_stride;
    a.layer = i; a.row_seq = row_seq; a.row_pos = row_pos; a.page_table = page_table; a.max_pages = max_pages;
    a.M = M; a.H = c.cfg.n_heads; a.freqs = c.freqs_flow; a.total_keys = total_keys;
    a.part = w.attn_part; a.splits = w.attn_part ? std::min(8, std::max(1, 296 / (M * c.cfg.n_heads))) : 1;
    if (!roped) launch_flow_rope_append(a, c.stream);
    launch_flow_attention(a, c.stream);
    LinearParams o = rows_linear(w.att, M, D, w.x, D, "flow.out");
    o.res = w.x; o.res_bs = 0; o.res_rs = D; o.w_stream = w_stream;
    run_linear(c, l.out, o);
    LinearParams f1 = rows_linear(w.h, M, D, w.ff, FF, "flow.ff1");
    f1.act = ACT_GELU; f1.w_stream = w_stream;
    run_norm_linear(c, l.ff1, w.x, M, D, l.ln2w, l.ln2b, 1e-5f, nullptr, nullptr, 0, w.h, f1);
    LinearParams f2 = rows_linear(w.ff, M, FF, w.x, D, "flow.ff2");
    f2.res = w.x; f2.res_bs = 0; f2.res_rs = D; f2.w_stream = w_stream;
    run_linear(c, l.ff2, f2);
  }
}

int take_pages(Ctx& c, int n, std::vector<int>* out) {
  if ((int)c.free_pages.size() < n)
    return fail(PTTS_ERR_NOMEM, "KV page pool exhausted: need %d pages, %zu free (kv_pool_tokens=%lld)", n,
                c.free_pages.size(), (long long)c.cfg.kv_pool_tokens);
  for (int i = 0; i < n; ++i) {
    out->push_back(c.free_pages.back());
    c.free_pages.pop_back();
  }
  return 0;
}

void voice_unref(Ctx& c, int id) {
  if (id < 0 || id >= (int)c.voices.size()) return;
  Voice& v = c.voices[id];
  if (v.refs > 0) --v.refs;
  if (v.refs == 0 && v.doomed) {
    for (int p : v.pages) c.free_pages.push_back(p);
    v = Voice{};
  }
}

}  // namespace

// ---- batch -----------------------------------------------------------------------------------------------
struct ptts_batch {
  Ctx* ctx = nullptr;
  int B = 0, max_pages = 0;
  std::vector<int> h_len, voice_ids, max_len;
  std::vector<std::vector<int>> slot_pages;     // private KV pages of every sequence slot
  std::vector<int> slot_voice;                  // voice each slot holds a reference on (-1: none)
  // continuous batching: parked slots (h_active = 0) stop growing; a slot can be re-initialised for a new utterance
  std::vector<int> h_active, h_page_table;
  int* d_active = nullptr;
  std::vector<ShiftEntry> h_shift;              // host copy of the streaming-conv state map
  void* mimi_tpl = nullptr;                     // Mimi state of one sequence right after the warm-up frames
  bool has_tpl = false;
  StatePieceDev* d_state_pieces = nullptr; int n_state_pieces = 0; int* d_restore_slots = nullptr;   // restore_mimi_slots
  std::vector<void*> allocs;
  std::vector<std::pair<void*, size_t>> zero_list;   // streaming state + scratch that a fresh batch starts zeroed
  int *d_cp_src = nullptr, *d_cp_dst = nullptr;
  bool prefilled = false;
  unsigned long long seed = 0x5eed5eedULL;
  int *d_page_table = nullptr, *d_len = nullptr, *d_bos = nullptr, *d_mimi_off = nullptr, *d_frame_idx = nullptr;
  unsigned long long* d_counter = nullptr;
  FlowWork fw;
  // head
  float *d_noise = nullptr, *d_x = nullptr, *d_latent = nullptr, *d_c = nullptr, *d_logit = nullptr;
  float *d_sy = nullptr, *d_ada = nullptr, *d_x1 = nullptr, *d_hh = nullptr, *d_u = nullptr, *d_v = nullptr;
  float* d_zero_lat = nullptr;
  // mimi
  float *d_zprev = nullptr, *d_c0 = nullptr, *d_mh = nullptr, *d_mqkv = nullptr, *d_mqrot = nullptr;
  float *d_matt = nullptr, *d_mff = nullptr, *d_audio = nullptr, *d_fin = nullptr;
  struct StageBuf { float *ct_in, *r_in, *hid; int T_in, T_out; };
  std::vector<StageBuf> sb;
  void* ring = nullptr;
  long long ring_layer_stride = 0, ring_kv_stride = 0;
  ShiftEntry* d_shift = nullptr;
  int n_shift = 0;
  // ---- tensor-core (bf16 operand) pipeline: chosen per part by the row count (want_tc)
  bool tc_head = false, tc_mimi = false;
  __nv_bfloat16 *d_c16 = nullptr, *d_sy16 = nullptr, *d_hh16 = nullptr, *d_u16 = nullptr;
  std::vector<TcGemm> g_cond, g_m1, g_m2;
  TcGemm g_ada, g_fin;
  // the whole flow head (all Euler steps) as ONE cluster chain launch: two op lists that differ in where the last
  // step writes the latent (d_latent / d_latent_b, the ping-pong buffers of the pipelined frame graph)
  ChainOp* d_head_ops[2] = {nullptr, nullptr};
  int n_head_ops = 0, head_nc = 0;
  __nv_bfloat16* d_x16 = nullptr;    // [B][64] bf16 copy of the head's current latent (zero beyond latent_dim)
  float* d_xm = nullptr;
  __nv_bfloat16 *d_mh16 = nullptr, *d_matt16 = nullptr, *d_mff16 = nullptr, *d_c0_16 = nullptr, *d_fin16 = nullptr;
  std::vector<TcGemm> g_mimi;
  TcGemm g_conv0;
  struct StageBuf16 { __nv_bfloat16 *ct_in, *r_in, *xraw, *hid; int T_in, T_out; TcGemm ct, r3, r1; };
  std::vector<StageBuf16> sb16;
  float* d_mrope_cs = nullptr;      // [B*T0][64] cos | sin of this frame's Mimi positions (fused qkv epilogue)
  SnTail sn_tail;                   // fused last resblock + output conv (valid: replaces sb16.back().r3/.r1 + final conv)
  float* d_bnd = nullptr;
  bool no_tail_env = false;         // PTTS_NO_SNTAIL at creation (a recycled arena must match the current setting)
  int T0 = 0;           // steps per frame at the SEANet input (upsample stride)
  int frame_samples = 0;
  // 16-bit PCM output: the kernels that produce the final samples also write them as int16, and the
  // frame's copy-out moves those (half the D2H bytes) instead of the fp32 samples
  bool pcm16 = false, graphs_pcm = false;
  short* d_pcm = nullptr;
  short *h_pcm = nullptr, *h2_pcm = nullptr;
  // pinned staging
  float *h_noise = nullptr, *h_latent = nullptr, *h_logit = nullptr, *h_audio = nullptr;
  // asynchronous staged steps (pipelined mode): odd frames use a second set of pinned buffers, so the host can fill
  // the noise of frame t+1 and enqueue it while frame t is still running; one completion event per set
  float *h2_noise = nullptr, *h2_latent = nullptr, *h2_logit = nullptr, *h2_audio = nullptr;
  bool async_staging = false, graphs_async = false;
  cudaEvent_t ev_set[2] = {nullptr, nullptr};
  // graphs: index = host_noise*2 + copy_out
  cudaGraphExec_t step_graph[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaGraphExec_t step_graph_alt = nullptr;      // host-I/O frame through the second staging set (async staged steps)
  long long step_graph_alt_launches = 0;
  long long step_graph_launches[4] = {0, 0, 0, 0};
  // pipelined mode: frame graph = { FlowLM step t } || { Mimi decode of latent t-1 } on two streams.
  // Latents ping-pong between d_latent (even frames) and d_latent_b (odd frames); index = parity*2 + host_io.
  bool pipelined = false;
  // slots re-initialised while the batch is in pipelined mode: the frame graph launched next still decodes the OLD
  // utterance's last latent for them, so their Mimi streaming state is restored right AFTER that graph
  std::vector<int> pending_mimi_reset;
  int cascade_len = 0;              // prefix length the captured graphs were built for
  long long frame_idx = 0;          // frames stepped since the last (re)initialisation
  float* d_latent_b = nullptr;
  cudaGraphExec_t pipe_graph[4] = {nullptr, nullptr, nullptr, nullptr};
  long long pipe_graph_launches[4] = {0, 0, 0, 0};
  std::map<int, std::pair<cudaGraphExec_t, long long>> mimi_graphs;   // keyed by n_frames
  float *d_lat_all = nullptr, *d_audio_all = nullptr;
  long long lat_all_cap = 0;
  float* h_chunk[2] = {nullptr, nullptr};       // pinned staging of ptts_batch_mimi_decode's chunked read-back

  int dalloc(void** p, size_t bytes) {
    CU(cudaMalloc(p, bytes ? bytes : 16));
    allocs.push_back(*p);
    return 0;
  }
};

namespace {

using Batch = ptts_batch;

void mimi_frame_tc(Batch& bt, const float* latent, int part = 0, bool advance = false, int seg_lo = 0, int seg_hi = 1 << 30);

// one Mimi frame for every sequence: latent [B][L] -> audio [B][frame_samples]
// advance: also move every sequence's ring offset on by one frame (folded into the end-of-frame launch)
void mimi_frame(Batch& bt, const float* latent, bool advance = false) {
  if (bt.tc_mimi) return mimi_frame_tc(bt, latent, 0, advance);
  Ctx& c = *bt.ctx;
  const ptts_config& g = c.cfg;
  const int B = bt.B, T = bt.T0, MD = g.mimi_d, SD = g.seanet_dim;
  const int k0 = g.kernel_size;
  const long long c0_bs = (long long)(T + k0 - 1) * SD;
  float* x = bt.d_c0 + (long long)(k0 - 1) * SD;     // residual stream lives in conv0's input rows
  launch_quant_upsample(latent, c.emb_std, c.emb_mean, c.wq, c.wu, bt.d_zprev, x, c0_bs, B, g.latent_dim, SD,
                        g.upsample_stride, c.stream);
  for (int i = 0; i < g.mimi_layers; ++i) {
    auto& l = c.ml[i];
    NormParams n{};
    n.X = x; n.x_bs = c0_bs; n.x_rs = MD; n.nb = B; n.T = T; n.C = MD;
    n.w = l.ln1w; n.b = l.ln1b; n.eps = 1e-5f;
    n.Y = bt.d_mh; n.y_bs = (long long)T * MD; n.y_rs = MD;
    launch_layernorm(n, c.stream);
    LinearParams q{};
    q.A = bt.d_mh; q.a_bs = (long long)T * MD; q.a_rs = MD; q.nb = B; q.T = T; q.taps = 1; q.C = MD;
    q.Y = bt.d_mqkv; q.y_bs = (long long)T * 3 * MD; q.y_rs = 3 * MD; q.out_scale = 1.f;
    q.tag = "mimi.qkv";
    run_linear(c, l.qkv, q);
    MimiAttnParams a{};
    a.qkv = bt.d_mqkv; a.q_rot = bt.d_mqrot; a.out = bt.d_matt;
    a.ring = bt.ring; a.kv_bf16 = c.bf16; a.layer_stride = bt.ring_layer_stride; a.kv_stride = bt.ring_kv_stride;
    a.layer = i; a.offset = bt.d_mimi_off; a.B = B; a.T = T; a.H = g.mimi_heads; a.context = g.mimi_context;
    a.freqs = c.freqs_mimi;
    launch_mimi_rope_ring(a, c.stream);
    launch_mimi_attention(a, c.stream);
    LinearParams o{};
    o.A = bt.d_matt; o.a_bs = (long long)T * MD; o.a_rs = MD; o.nb = B; o.T = T; o.taps = 1; o.C = MD;
    o.col_scale = l.ls1; o.res = x; o.res_bs = c0_bs; o.res_rs = MD;
    o.Y = x; o.y_bs = c0_bs; o.y_rs = MD; o.out_scale = 1.f;
    o.tag = "mimi.out";
    run_linear(c, l.out, o);
    n.w = l.ln2w; n.b = l.ln2b;
    launch_layernorm(n, c.stream);
    LinearParams f1 = q;
    f1.Y = bt.d_mff; f1.y_bs = (long long)T * g.mimi_ffn; f1.y_rs = g.mimi_ffn; f1.act = ACT_GELU;
    f1.tag = "mimi.ff1";
    run_linear(c, l.ff1, f1);
    LinearParams f2 = o;
    f2.tag = "mimi.ff2";
    f2.A = bt.d_mff; f2.a_bs = (long long)T * g.mimi_ffn; f2.a_rs = g.mimi_ffn; f2.C = g.mimi_ffn;
    f2.col_scale = l.ls2;
    run_linear(c, l.ff2, f2);
  }
  // SEANet: conv0 -> [ELU, convtr, resblock] x n -> ELU -> last conv
  {
    LinearParams p{};
    p.A = bt.d_c0; p.a_bs = c0_bs; p.a_rs = SD; p.nb = B; p.T = T; p.taps = k0; p.C = SD;
    auto& s0 = bt.sb[0];
    const int c1 = c.stages[0].c_in;
    p.Y = s0.ct_in + c1; p.y_bs = (long long)(T + 1) * c1; p.y_rs = c1; p.out_scale = 1.f;
    p.tag = "sn.conv0";
    run_linear(c, c.conv0, p);
  }
  for (size_t r = 0; r < c.stages.size(); ++r) {
    auto& st = c.stages[r];
    auto& b = bt.sb[r];
    const int rk = g.res_kernel_size;
    const long long r_bs = (long long)(b.T_out + rk - 1) * st.c_out;
    {  // ELU -> transposed conv as a 2-tap GEMM with N = stride*c_out; rows land time-major in r_in
      LinearParams p{};
      p.A = b.ct_in; p.a_bs = (long long)(b.T_in + 1) * st.c_in; p.a_rs = st.c_in;
      p.nb = B; p.T = b.T_in; p.taps = 2; p.C = st.c_in; p.a_pro = ACT_ELU;
      p.Y = b.r_in + (long long)(rk - 1) * st.c_out; p.y_bs = r_bs; p.y_rs = (long long)st.stride * st.c_out;
      p.out_scale = 1.f;
      static const char* kCt[] = {"sn.ct0", "sn.ct1", "sn.ct2", "sn.ct3", "sn.ct4", "sn.ct5", "sn.ct6", "sn.ct7"};
      p.tag = kCt[r];
      run_linear(c, st.ct, p);
    }
    {  // resblock: x + conv_k1(ELU(conv_k3(ELU(x))))
      LinearParams p{};
      p.A = b.r_in; p.a_bs = r_bs; p.a_rs = st.c_out; p.nb = B; p.T = b.T_out; p.taps = rk; p.C = st.c_out;
      p.a_pro = ACT_ELU;
      p.Y = b.hid; p.y_bs = (long long)b.T_out * st.hidden; p.y_rs = st.hidden; p.out_scale = 1.f;
      static const char* kR3[] = {"sn.r3_0", "sn.r3_1", "sn.r3_2", "sn.r3_3", "sn.r3_4", "sn.r3_5", "sn.r3_6", "sn.r3_7"};
      static const char* kR1[] = {"sn.r1_0", "sn.r1_1", "sn.r1_2", "sn.r1_3", "sn.r1_4", "sn.r1_5", "sn.r1_6", "sn.r1_7"};
      p.tag = kR3[r];
      run_linear(c, st.r3, p);
      LinearParams q{};
      q.A = b.hid; q.a_bs = (long long)b.T_out * st.hidden; q.a_rs = st.hidden; q.nb = B; q.T = b.T_out;
      q.taps = 1; q.C = st.hidden; q.a_pro = ACT_ELU;
      q.res = b.r_in + (long long)(rk - 1) * st.c_out; q.res_bs = r_bs; q.res_rs = st.c_out;
      q.out_scale = 1.f;
      if (r + 1 < c.stages.size()) {
        auto& nb = bt.sb[r + 1];
        q.Y = nb.ct_in + st.c_out; q.y_bs = (long long)(b.T_out + 1) * st.c_out; q.y_rs = st.c_out;
      } else {
        q.Y = bt.d_fin + (long long)(c.fin_taps - 1) * st.c_out;
        q.y_bs = (long long)(b.T_out + c.fin_taps - 1) * st.c_out; q.y_rs = st.c_out;
      }
      q.tag = kR1[r];
      run_linear(c, st.r1, q);
    }
  }
  launch_final_conv(bt.d_fin, (long long)(bt.frame_samples + c.fin_taps - 1) * c.fin_c, c.fin_w, c.fin_b, bt.d_audio,
                    bt.frame_samples, B, bt.frame_samples, c.fin_c, c.fin_taps, c.stream, bt.pcm16 ? bt.d_pcm : nullptr);
  launch_state_shift(bt.d_shift, bt.n_shift, B, c.stream, nullptr, 0, nullptr, 0, advance ? bt.d_mimi_off : nullptr, bt.T0);
}

// Mimi frame on the tensor-core path: bf16 operands everywhere, every conv / transposed conv / linear is a
// tcgen05 GEMM whose epilogue already writes the next GEMM's (ELU'd) bf16 input, incl. the carried state rows.
// seg_lo / seg_hi: only the launches with index in [seg_lo, seg_hi) are issued (the interleaved pipelined frame cuts the
// decoder into slices that it starts behind the FlowLM attention kernels); one index per logical step, whatever the
// configuration, see mimi_tc_launches().
void mimi_frame_tc(Batch& bt, const float* latent, int part, bool advance, int seg_lo, int seg_hi) {   // part: 0 all, 1 transformer, 2 SEANet
  Ctx& c = *bt.ctx;
  const ptts_config& g = c.cfg;
  const int B = bt.B, T = bt.T0, MD = g.mimi_d, SD = g.seanet_dim;
  int k = 0;
  auto in = [&]() { const bool t = k >= seg_lo && k < seg_hi; ++k; return t; };
  if (part != 2) {
  if (in())
    launch_quant_upsample(latent, c.emb_std, c.emb_mean, c.wq, c.wu, bt.d_zprev, bt.d_xm, (long long)T * SD, B,
                          g.latent_dim, SD, g.upsample_stride, c.stream);
  static const bool fuse_rope = [] { const char* v = getenv("PTTS_NO_ROPE_FUSE"); return !(v && v[0] == '1'); }();
  if (in() && fuse_rope) launch_rope_table(bt.d_mimi_off, c.freqs_mimi, bt.d_mrope_cs, B * T, T, c.stream);
  for (int i = 0; i < g.mimi_layers; ++i) {
    auto& l = c.ml[i];
    const TcGemm* gm = &bt.g_mimi[(size_t)i * 4];
    NormParams n{};
    n.X = bt.d_xm; n.x_bs = (long long)T * MD; n.x_rs = MD; n.nb = B; n.T = T; n.C = MD;
    n.w = l.ln1w; n.b = l.ln1b; n.eps = 1e-5f;
    n.Y16 = bt.d_mh16; n.y_bs = (long long)T * MD; n.y_rs = MD;
    if (in()) launch_layernorm(n, c.stream);
    if (in()) {
      if (fuse_rope) {   // RoPE + ring write in the qkv epilogue
        TcGemm q = gm[0];
        auto& e = q.e;
        e.y32 = nullptr;
        e.rope_cs = bt.d_mrope_cs; e.q_rot = bt.d_mqrot;
        e.kv_layer = reinterpret_cast<__nv_bfloat16*>(bt.ring) + (long long)i * bt.ring_layer_stride;
        e.kv_row_pos = bt.d_mimi_off; e.kv_heads = g.mimi_heads;
        e.kv_ring = g.mimi_context; e.kv_v_offset = bt.ring_kv_stride;
        gemm_tc_launch(q, c.stream);
      } else {
        gemm_tc_launch(gm[0], c.stream);
      }
    }
    if (in()) {
      MimiAttnParams a{};
      a.qkv = bt.d_mqkv; a.q_rot = bt.d_mqrot; a.out16 = bt.d_matt16;
      a.ring = bt.ring; a.kv_bf16 = c.bf16; a.layer_stride = bt.ring_layer_stride; a.kv_stride = bt.ring_kv_stride;
      a.layer = i; a.offset = bt.d_mimi_off; a.B = B; a.T = T; a.H = g.mimi_heads; a.context = g.mimi_context;
      a.freqs = c.freqs_mimi;
      if (!fuse_rope) launch_mimi_rope_ring(a, c.stream);
      launch_mimi_attention(a, c.stream);
    }
    if (in()) gemm_tc_launch(gm[1], c.stream);
    n.w = l.ln2w; n.b = l.ln2b;
    if (in()) launch_layernorm(n, c.stream);
    if (in()) gemm_tc_launch(gm[2], c.stream);
    if (in()) gemm_tc_launch(gm[3], c.stream);
  }
  }
  if (part == 1) return;
  if (in()) gemm_tc_launch(bt.g_conv0, c.stream);
  for (size_t r = 0; r < bt.sb16.size(); ++r) {
    auto& sb = bt.sb16[r];
    if (in()) gemm_tc_launch(sb.ct, c.stream);
    if (bt.sn_tail.valid && r + 1 == bt.sb16.size()) {
      if (in()) {
        SnTail tl = bt.sn_tail;
        tl.pcm = bt.pcm16 ? bt.d_pcm : nullptr;
        sn_tail_launch(tl, c.stream, false);             // its boundary fix-up rides in the state-shift launch below
      }
      break;
    }
    if (in()) gemm_tc_launch(sb.r3, c.stream);
    if (in()) gemm_tc_launch(sb.r1, c.stream);
  }
  if (!bt.sn_tail.valid) {
    if (in())
      launch_final_conv16(bt.d_fin16, (long long)(bt.frame_samples + c.fin_taps - 1) * c.fin_c, c.fin_w, c.fin_b,
                          bt.d_audio, bt.frame_samples, B, bt.frame_samples, c.fin_c, c.fin_taps, c.stream,
                          bt.pcm16 ? bt.d_pcm : nullptr);
  }
  if (in())
    launch_state_shift(bt.d_shift, bt.n_shift, B, c.stream, bt.d_audio, bt.frame_samples, bt.sn_tail.valid ? bt.d_bnd : nullptr,
                       bt.frame_samples / 128, advance ? bt.d_mimi_off : nullptr, bt.T0, bt.pcm16 ? bt.d_pcm : nullptr);
}

// number of launch indices mimi_frame_tc(part = 0) walks through
int mimi_tc_launches(Batch& bt) {
  const ptts_config& g = bt.ctx->cfg;
  const int R = (int)bt.sb16.size();
  const int sn = bt.sn_tail.valid ? (1 + 3 * (R - 1) + 2) : (1 + 3 * R + 1);
  return 2 + 7 * g.mimi_layers + sn + 1;
}

// flow head on the tensor-core path (M = B rows)
void flow_head_tc(Batch& bt, float* lat_out) {
  Ctx& c = *bt.ctx;
  const ptts_config& g = c.cfg;
  const int B = bt.B, L = g.latent_dim, fd = g.flow_dim;
  const int n = g.lsd_decode_steps;
  if (bt.n_head_ops > 0 && (lat_out == bt.d_latent || lat_out == bt.d_latent_b)) {
    const double fl = 2.0 * B * n * ((double)g.d_model * fd + (double)c.n_ada * fd + 64.0 * fd + 2.0 * g.flow_depth * fd * fd + (double)fd * L);
    chain_launch(bt.d_head_ops[lat_out == bt.d_latent ? 0 : 1], bt.n_head_ops, B, bt.head_nc, "flow.head", fl, fl / B, c.stream);
    return;
  }
  for (int i = 0; i < n; ++i) {
    gemm_tc_launch(bt.g_cond[i], c.stream);
    gemm_tc_launch(bt.g_ada, c.stream);
    if (!(c.in_proj.w16 && launch_small_k_linear(c.in_proj.w16, c.in_proj.bias, bt.d_x, bt.d_x1, B, L, fd, "head.in", c.stream)))
      run_linear(c, c.in_proj, rows_linear(bt.d_x, B, L, bt.d_x1, fd, "head.in"));
    for (int r = 0; r < g.flow_depth; ++r) {
      const float* ada = bt.d_ada + (long long)r * 3 * fd;
      rows_norm(c, bt.d_x1, B, fd, c.rb[r].lnw, c.rb[r].lnb, 1e-6f, nullptr, ada + fd, ada, c.n_ada, bt.d_hh16, nullptr, 0);
      gemm_tc_launch(bt.g_m1[r], c.stream);
      gemm_tc_launch(bt.g_m2[r], c.stream);
    }
    const float* adaf = bt.d_ada + (long long)g.flow_depth * 3 * fd;
    rows_norm(c, bt.d_x1, B, fd, nullptr, nullptr, 1e-6f, nullptr, adaf + fd, adaf, c.n_ada, bt.d_hh16, nullptr, 0);
    if (i == n - 1) {
      TcGemm f = bt.g_fin;          // x + v / n of the last step is the frame's latent
      f.e.y32 = lat_out;
      gemm_tc_launch(f, c.stream);
    } else {
      gemm_tc_launch(bt.g_fin, c.stream);
    }
  }
}

bool sn_tail_disabled() {
  const char* nt = getenv("PTTS_NO_SNTAIL");
  return nt && nt[0] == '1';
}

int build_batch_tc(Batch& t) {
  Ctx& c = *t.ctx;
  const ptts_config& g = c.cfg;
  const int B = t.B, D = g.d_model, L = g.latent_dim, fd = g.flow_dim;
  auto bz = [&](__nv_bfloat16** p, size_t n) -> int {
    RET(t.dalloc((void**)p, n * 2));
    t.zero_list.push_back({*p, n * 2});
    return 0;
  };
  if (t.tc_head) {
    RET(bz(&t.d_c16, (size_t)B * D));
    RET(bz(&t.d_sy16, (size_t)B * fd));
    RET(bz(&t.d_hh16, (size_t)B * fd));
    RET(bz(&t.d_u16, (size_t)B * fd));
    bool ok = true;
    t.g_cond.assign(g.lsd_decode_steps, TcGemm{});
    for (int i = 0; i < g.lsd_decode_steps && ok; ++i) {
      ok = plan_tc(&t.g_cond[i], t.d_c16, 0, D, 1, B, 1, D, c.cond, "head.cond", 1, 1);
      t.g_cond[i].e.bias = c.cond_bias_step[i];
      t.g_cond[i].e.act = ACT_SILU;
      t.g_cond[i].e.y16 = t.d_sy16; t.g_cond[i].e.y16_rs = fd;
    }
    ok = ok && plan_tc(&t.g_ada, t.d_sy16, 0, fd, 1, B, 1, fd, c.ada_all, "head.ada");
    t.g_ada.e.y32 = t.d_ada; t.g_ada.e.y32_rs = c.n_ada;
    t.g_m1.assign(g.flow_depth, TcGemm{});
    t.g_m2.assign(g.flow_depth, TcGemm{});
    for (int r = 0; r < g.flow_depth && ok; ++r) {
      ok = plan_tc(&t.g_m1[r], t.d_hh16, 0, fd, 1, B, 1, fd, c.rb[r].m1, "head.m1", 1, 1) &&
           plan_tc(&t.g_m2[r], t.d_u16, 0, fd, 1, B, 1, fd, c.rb[r].m2, "head.m2");
      t.g_m1[r].e.act = ACT_SILU; t.g_m1[r].e.y16 = t.d_u16; t.g_m1[r].e.y16_rs = fd;
      auto& e = t.g_m2[r].e;
      e.row_gate = t.d_ada + (long long)r * 3 * fd + 2 * fd; e.gate_rs = c.n_ada;
      e.res32 = t.d_x1; e.res32_rs = fd; e.y32 = t.d_x1; e.y32_rs = fd;
    }
    ok = ok && plan_tc(&t.g_fin, t.d_hh16, 0, fd, 1, B, 1, fd, c.fin, "head.fin");
    t.g_fin.e.out_scale = 1.0f / (float)g.lsd_decode_steps;
    t.g_fin.e.res32 = t.d_x; t.g_fin.e.res32_rs = L; t.g_fin.e.y32 = t.d_x; t.g_fin.e.y32_rs = L;
    if (!ok) return fail(PTTS_ERR_CUDA, "tcgen05 plan failed for the flow head (B=%d)", B);
    for (auto& gm : t.g_cond) gemm_tc_bind_outputs(&gm);
    for (auto& gm : t.g_m1) gemm_tc_bind_outputs(&gm);
    // ---- single-launch flow head (cluster chain kernel)
    t.n_head_ops = 0;
    const int nc = chain_enabled() ? chain_cluster_size() : 0;
    if (nc > 0 && B <= 512 && L == 32 && c.in_proj_pad && fd % nc == 0 && fd / nc <= 64 && chain_pick_bn(fd, nc) == fd / nc) {
      RET(bz(&t.d_x16, (size_t)B * 64));
      const int nst = g.lsd_decode_steps, per = 4 + 2 * g.flow_depth;
      std::vector<ChainOp> ops((size_t)nst * per);
      bool cok = true;
      auto gemm = [&](ChainOp& o, const __nv_bfloat16* a, int K, const __nv_bfloat16* w16, int N, const float* bias, int kind) {
        o.K = K; o.N = N; o.bn = chain_pick_bn(N, nc); o.kind = kind; o.bias = bias;
        cok = cok && w16 && chain_step_ok(N, K, nc) && chain_encode_a(&o.tm_a, a, B, K, K) && chain_encode_w(&o.tm_w, w16, N, K, o.bn);
      };
      for (int v = 0; v < 2 && cok; ++v) {
        memset(ops.data(), 0, ops.size() * sizeof(ChainOp));
        for (int i = 0; i < nst; ++i) {
          ChainOp* o = &ops[(size_t)i * per];
          gemm(o[0], t.d_c16, D, c.cond.w16, fd, c.cond_bias_step[i], CH_STORE16);          // silu(t_emb + cond_embed(c))
          o[0].act = ACT_SILU; o[0].y16 = t.d_sy16; o[0].y_rs = fd;
          gemm(o[1], t.d_sy16, fd, c.ada_all.w16, c.n_ada, c.ada_all.bias, CH_STORE32);     // every AdaLN modulation at once
          o[1].y32 = t.d_ada; o[1].y_rs = c.n_ada;
          gemm(o[2], t.d_x16, 64, c.in_proj_pad, fd, c.in_proj.bias, CH_RES_LN);            // x1 = input_proj(x); first in_ln
          o[2].x = t.d_x1; o[2].x_rs = fd; o[2].x_init = 1;
          for (int r = 0; r < g.flow_depth; ++r) {
            ChainOp& pre = (r == 0) ? o[2] : o[2 + 2 * r];                                   // the step whose epilogue normalises for block r
            const float* ada = t.d_ada + (long long)r * 3 * fd;
            pre.ln_on = 1; pre.ln_w = c.rb[r].lnw; pre.ln_b = c.rb[r].lnb; pre.ln_eps = 1e-6f;
            pre.mod_shift = ada; pre.mod_scale = ada + fd; pre.mod_rs = c.n_ada; pre.h16 = t.d_hh16; pre.h_rs = fd;
            gemm(o[3 + 2 * r], t.d_hh16, fd, c.rb[r].m1.w16, fd, c.rb[r].m1.bias, CH_STORE16);
            o[3 + 2 * r].act = ACT_SILU; o[3 + 2 * r].y16 = t.d_u16; o[3 + 2 * r].y_rs = fd;
            gemm(o[4 + 2 * r], t.d_u16, fd, c.rb[r].m2.w16, fd, c.rb[r].m2.bias, CH_RES_LN);
            o[4 + 2 * r].x = t.d_x1; o[4 + 2 * r].x_rs = fd; o[4 + 2 * r].gate = ada + 2 * fd; o[4 + 2 * r].gate_rs = c.n_ada;
          }
          {
            ChainOp& last = o[2 + 2 * g.flow_depth];                                         // m2 of the last block: final-layer norm
            const float* adaf = t.d_ada + (long long)g.flow_depth * 3 * fd;
            last.ln_on = 1; last.ln_w = nullptr; last.ln_b = nullptr; last.ln_eps = 1e-6f;
            last.mod_shift = adaf; last.mod_scale = adaf + fd; last.mod_rs = c.n_ada; last.h16 = t.d_hh16; last.h_rs = fd;
          }
          ChainOp& f = o[per - 1];
          gemm(f, t.d_hh16, fd, c.fin.w16, L, c.fin.bias, CH_FIN);                           // x += v / n
          f.out_scale = 1.0f / (float)nst; f.lat_in = t.d_x; f.lat_rs = L;
          f.lat_out = (i == nst - 1) ? (v == 0 ? t.d_latent : t.d_latent_b) : t.d_x;
          f.lat16 = (i == nst - 1) ? nullptr : t.d_x16; f.lat16_rs = 64;
        }
        if (!cok) break;
        if (!t.d_head_ops[v]) RET(t.dalloc((void**)&t.d_head_ops[v], ops.size() * sizeof(ChainOp)));
        CU(cudaMemcpyAsync(t.d_head_ops[v], ops.data(), ops.size() * sizeof(ChainOp), cudaMemcpyHostToDevice, c.stream));
        CU(cudaStreamSynchronize(c.stream));
      }
      if (cok) { t.n_head_ops = nst * per; t.head_nc = nc; }
    }
  }
  if (t.tc_mimi) {
    const int T = t.T0, MD = g.mimi_d, SD = g.seanet_dim, FFm = g.mimi_ffn, k0 = g.kernel_size, rk = g.res_kernel_size;
    RET(t.dalloc((void**)&t.d_xm, (size_t)B * T * MD * 4));
    t.zero_list.push_back({t.d_xm, (size_t)B * T * MD * 4});
    RET(bz(&t.d_mh16, (size_t)B * T * MD));
    RET(t.dalloc((void**)&t.d_mrope_cs, (size_t)B * T * 64 * 4));
    RET(bz(&t.d_matt16, (size_t)B * T * MD));
    RET(bz(&t.d_mff16, (size_t)B * T * FFm));
    RET(bz(&t.d_c0_16, (size_t)B * (T + k0 - 1) * SD));
    std::vector<ShiftEntry> sh;
    sh.push_back({t.d_c0_16, (long long)(T + k0 - 1) * SD, T, k0 - 1, SD, 2});
    bool ok = true;
    t.g_mimi.assign((size_t)g.mimi_layers * 4, TcGemm{});
    for (int i = 0; i < g.mimi_layers && ok; ++i) {
      auto& l = c.ml[i];
      TcGemm* gm = &t.g_mimi[(size_t)i * 4];
      ok = plan_tc(&gm[0], t.d_mh16, (long long)T * MD, MD, B, T, 1, MD, l.qkv, "mimi.qkv") &&
           plan_tc(&gm[1], t.d_matt16, (long long)T * MD, MD, B, T, 1, MD, l.out, "mimi.out") &&
           plan_tc(&gm[2], t.d_mh16, (long long)T * MD, MD, B, T, 1, MD, l.ff1, "mimi.ff1", 1, 1) &&
           plan_tc(&gm[3], t.d_mff16, (long long)T * FFm, FFm, B, T, 1, FFm, l.ff2, "mimi.ff2", 1, i == g.mimi_layers - 1 ? 1 : 0);
      gm[0].e.y32 = t.d_mqkv; gm[0].e.y32_bs = (long long)T * 3 * MD; gm[0].e.y32_rs = 3 * MD;
      for (int k : {1, 3}) {
        auto& e = gm[k].e;
        e.col_scale = (k == 1) ? l.ls1 : l.ls2;
        e.res32 = t.d_xm; e.res32_bs = (long long)T * MD; e.res32_rs = MD;
        e.y32 = t.d_xm; e.y32_bs = (long long)T * MD; e.y32_rs = MD;
      }
      gm[2].e.act = ACT_GELU; gm[2].e.y16 = t.d_mff16; gm[2].e.y16_bs = (long long)T * FFm; gm[2].e.y16_rs = FFm;
      if (i == g.mimi_layers - 1) {   // the transformer output is conv0's input: bf16 copy behind the 6 state rows
        gm[3].e.y16 = t.d_c0_16 + (long long)(k0 - 1) * SD;
        gm[3].e.y16_bs = (long long)(T + k0 - 1) * SD; gm[3].e.y16_rs = SD; gm[3].e.y16_act = ACT_NONE;
      }
    }
    t.sb16.resize(g.n_ratios);
    int Tin = T;
    for (int r = 0; r < g.n_ratios; ++r) {
      auto& st = c.stages[r];
      auto& b = t.sb16[r];
      b.T_in = Tin; b.T_out = Tin * st.stride;
      RET(bz(&b.ct_in, (size_t)B * (Tin + 1) * st.c_in));
      RET(bz(&b.r_in, (size_t)B * (b.T_out + rk - 1) * st.c_out));
      RET(bz(&b.xraw, (size_t)B * b.T_out * st.c_out));
      RET(bz(&b.hid, (size_t)B * b.T_out * st.hidden));
      sh.push_back({b.ct_in, (long long)(Tin + 1) * st.c_in, Tin, 1, st.c_in, 2});
      if (rk > 1) sh.push_back({b.r_in, (long long)(b.T_out + rk - 1) * st.c_out, b.T_out, rk - 1, st.c_out, 2});
      Tin = b.T_out;
    }
    RET(bz(&t.d_fin16, (size_t)B * (Tin + c.fin_taps - 1) * c.fin_c));
    if (c.fin_taps > 1) sh.push_back({t.d_fin16, (long long)(Tin + c.fin_taps - 1) * c.fin_c, Tin, c.fin_taps - 1, c.fin_c, 2});
    // conv0: ELU'd output lands behind the one state row of the first transposed conv
    ok = ok && plan_tc(&t.g_conv0, t.d_c0_16, (long long)(T + k0 - 1) * SD, SD, B, T, k0, SD, c.conv0, "sn.conv0", 1, 1);
    {
      auto& e = t.g_conv0.e;
      const int c1 = c.stages[0].c_in;
      e.y16 = t.sb16[0].ct_in + c1; e.y16_bs = (long long)(T + 1) * c1; e.y16_rs = c1; e.y16_act = ACT_ELU;
    }
    static const char* kCt[] = {"sn.ct0", "sn.ct1", "sn.ct2", "sn.ct3", "sn.ct4", "sn.ct5", "sn.ct6", "sn.ct7"};
    static const char* kR3[] = {"sn.r3_0", "sn.r3_1", "sn.r3_2", "sn.r3_3", "sn.r3_4", "sn.r3_5", "sn.r3_6", "sn.r3_7"};
    static const char* kR1[] = {"sn.r1_0", "sn.r1_1", "sn.r1_2", "sn.r1_3", "sn.r1_4", "sn.r1_5", "sn.r1_6", "sn.r1_7"};
    for (int r = 0; r < g.n_ratios && ok; ++r) {
      auto& st = c.stages[r];
      auto& b = t.sb16[r];
      const long long r_bs = (long long)(b.T_out + rk - 1) * st.c_out;
      ok = plan_tc(&b.ct, b.ct_in, (long long)(b.T_in + 1) * st.c_in, st.c_in, B, b.T_in, 2, st.c_in, st.ct, kCt[r], 1, 2) &&
           plan_tc(&b.r3, b.r_in, r_bs, st.c_out, B, b.T_out, rk, st.c_out, st.r3, kR3[r], 1, 1) &&
           plan_tc(&b.r1, b.hid, (long long)b.T_out * st.hidden, st.hidden, B, b.T_out, 1, st.hidden, st.r1, kR1[r], 1, 1);
      if (!ok) break;
      {  // transposed conv: x (raw, residual) and ELU(x) (resblock input, behind its state rows)
        auto& e = b.ct.e;
        e.y16 = b.r_in + (long long)(rk - 1) * st.c_out; e.y16_bs = r_bs; e.y16_rs = (long long)st.stride * st.c_out;
        e.y16_act = ACT_ELU;
        e.yraw16 = b.xraw; e.yraw16_bs = (long long)b.T_out * st.c_out; e.yraw16_rs = (long long)st.stride * st.c_out;
      }
      {
        auto& e = b.r3.e;
        e.y16 = b.hid; e.y16_bs = (long long)b.T_out * st.hidden; e.y16_rs = st.hidden; e.y16_act = ACT_ELU;
      }
      {  // x + conv_k1(...): ELU'd straight into the next layer's input
        auto& e = b.r1.e;
        e.res16 = b.xraw; e.res16_bs = (long long)b.T_out * st.c_out; e.res16_rs = st.c_out;
        e.y16_act = ACT_ELU;
        if (r + 1 < g.n_ratios) {
          e.y16 = t.sb16[r + 1].ct_in + st.c_out; e.y16_bs = (long long)(b.T_out + 1) * st.c_out; e.y16_rs = st.c_out;
        } else {
          e.y16 = t.d_fin16 + (long long)(c.fin_taps - 1) * st.c_out;
          e.y16_bs = (long long)(b.T_out + c.fin_taps - 1) * st.c_out; e.y16_rs = st.c_out;
        }
      }
    }
    if (!ok) return fail(PTTS_ERR_CUDA, "tcgen05 plan failed for the Mimi decoder (B=%d)", B);
    {
      // 64-channel tail: one fused kernel instead of conv_k3 / conv_k1 / output conv (PTTS_NO_SNTAIL=1 keeps them apart)
      const bool no_tail = sn_tail_disabled();       // read per batch so a test can compare both paths
      t.no_tail_env = no_tail;
      auto& st = c.stages[g.n_ratios - 1];
      auto& b = t.sb16[g.n_ratios - 1];
      if (!no_tail && c.fin_c == st.c_out && st.r3.w16 && st.r1.w16 &&
          sn_tail_plan(&t.sn_tail, b.r_in, b.xraw, B, b.T_out, st.c_out, st.hidden, rk, c.fin_taps, st.r3.w16, st.r1.w16)) {
        const size_t n_bnd = (size_t)B * (b.T_out / 128 + 1) * 4;
        RET(t.dalloc((void**)&t.d_bnd, n_bnd * 4));
        t.zero_list.push_back({t.d_bnd, n_bnd * 4});
        t.sn_tail.b1 = st.r3.bias; t.sn_tail.b2 = st.r1.bias; t.sn_tail.wf = c.fin_w; t.sn_tail.bf = c.fin_b;
        t.sn_tail.audio = t.d_audio; t.sn_tail.audio_bs = t.frame_samples; t.sn_tail.bnd = t.d_bnd;
      }
    }
    for (auto& gm : t.g_mimi) gemm_tc_bind_outputs(&gm);
    gemm_tc_bind_outputs(&t.g_conv0);
    for (auto& b : t.sb16) {
      gemm_tc_bind_outputs(&b.ct);
      gemm_tc_bind_outputs(&b.r3);
      gemm_tc_bind_outputs(&b.r1);
    }
    t.n_shift = (int)sh.size();
    t.h_shift = sh;
    RET(t.dalloc((void**)&t.d_shift, sh.size() * sizeof(ShiftEntry)));
    CU(cudaMemcpyAsync(t.d_shift, sh.data(), sh.size() * sizeof(ShiftEntry), cudaMemcpyHostToDevice, c.stream));
  }
  return 0;
}

// FlowLM decode step + EOS + flow head; leaves the new latent in d_latent
void flow_step(Batch& bt, bool host_noise, int part = 0, const float* lat_in = nullptr, float* lat_out = nullptr) {
  // part: 0 all, 1 backbone, 2 EOS + flow head; lat_in / lat_out default to the single latent buffer
  Ctx& c = *bt.ctx;
  if (!lat_in) lat_in = bt.d_latent;
  if (!lat_out) lat_out = bt.d_latent;
  const ptts_config& g = c.cfg;
  const int B = bt.B, D = g.d_model, L = g.latent_dim, fd = g.flow_dim;
  if (part != 2) {
    launch_input_rows(c.w_in, c.bos, lat_in, bt.d_bos, bt.fw.x, B, D, L, c.stream);
    long long total_keys = 0;
    for (int l : bt.h_len) total_keys += l + 1;
    flow_layers(c, bt.fw, B, nullptr, bt.d_len, bt.d_page_table, bt.max_pages, total_keys);
  }
  if (part == 1) return;
  // out_norm + EOS logit, and in the same launch the start noise x0 of the flow head (Philox or the host's draw)
  const float nclamp = (g.noise_clamp >= 0.f) ? g.noise_clamp : -1.f;
  if (L <= 128) {
    launch_final_norm_eos(bt.fw.x, nullptr, c.outn_w, c.outn_b, c.eos_w, c.eos_b, bt.d_c, bt.tc_head ? bt.d_c16 : nullptr,
                          bt.d_logit, B, D, bt.fw.ws_ff2, bt.fw.pend_n, (long long)B * D, c.stream,
                          bt.d_noise, bt.d_x, L, sqrtf(g.temp), nclamp, host_noise ? 0 : 1, bt.d_counter,
                          bt.n_head_ops > 0 ? bt.d_x16 : nullptr);
  } else {
    launch_final_norm_eos(bt.fw.x, nullptr, c.outn_w, c.outn_b, c.eos_w, c.eos_b, bt.d_c, bt.tc_head ? bt.d_c16 : nullptr,
                          bt.d_logit, B, D, bt.fw.ws_ff2, bt.fw.pend_n, (long long)B * D, c.stream);
    launch_noise_prep(bt.d_noise, bt.d_x, B * L, sqrtf(g.temp), nclamp, host_noise ? 0 : 1, bt.d_counter, c.stream);
  }
  const int n = g.lsd_decode_steps;
  if (bt.tc_head) {
    flow_head_tc(bt, lat_out);      // the last Euler step writes the new latent straight to lat_out
    return;
  }
  for (int i = 0; i < n; ++i) {
    LinearParams ce = rows_linear(bt.d_c, B, D, bt.d_sy, fd, "head.cond");
    ce.bias = c.cond_bias_step[i];
    ce.act = ACT_SILU;
    run_linear(c, c.cond, ce);                                     // silu(t_emb + cond_embed(c))
    run_linear(c, c.ada_all, rows_linear(bt.d_sy, B, fd, bt.d_ada, c.n_ada, "head.ada"));   // every AdaLN modulation at once
    run_linear(c, c.in_proj, rows_linear(bt.d_x, B, L, bt.d_x1, fd, "head.in"));
    for (int r = 0; r < g.flow_depth; ++r) {
      const float* ada = bt.d_ada + (long long)r * 3 * fd;
      LinearParams m1 = rows_linear(bt.d_hh, B, fd, bt.d_u, fd, "head.m1");
      m1.act = ACT_SILU;
      run_norm_linear(c, c.rb[r].m1, bt.d_x1, B, fd, c.rb[r].lnw, c.rb[r].lnb, 1e-6f, ada + fd, ada, c.n_ada, bt.d_hh, m1);
      LinearParams m2 = rows_linear(bt.d_u, B, fd, bt.d_x1, fd, "head.m2");
      m2.row_gate = ada + 2 * fd; m2.gate_bs = 0; m2.gate_rs = c.n_ada;
      m2.res = bt.d_x1; m2.res_bs = 0; m2.res_rs = fd;
      run_linear(c, c.rb[r].m2, m2);
    }
    const float* adaf = bt.d_ada + (long long)g.flow_depth * 3 * fd;
    LinearParams fo = rows_linear(bt.d_hh, B, fd, (i == n - 1) ? lat_out : bt.d_x, L, "head.fin");     // x += v / n
    fo.out_scale = 1.0f / (float)n;
    fo.res = bt.d_x; fo.res_bs = 0; fo.res_rs = L;
    run_norm_linear(c, c.fin, bt.d_x1, B, fd, nullptr, nullptr, 1e-6f, adaf + fd, adaf, c.n_ada, bt.d_hh, fo);
  }
}

void full_step(Batch& bt, bool host_noise, bool copy_out, bool alt = false) {
  Ctx& c = *bt.ctx;
  const int B = bt.B, L = c.cfg.latent_dim;
  // alt: second set of pinned staging buffers (odd frames of an async-staged batch)
  float *hn = alt ? bt.h2_noise : bt.h_noise, *hl = alt ? bt.h2_latent : bt.h_latent, *hg = alt ? bt.h2_logit : bt.h_logit,
        *ha = alt ? bt.h2_audio : bt.h_audio;
  short* hp = alt ? bt.h2_pcm : bt.h_pcm;
  if (host_noise)
    cudaMemcpyAsync(bt.d_noise, hn, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, c.stream);
  flow_step(bt, host_noise);
  mimi_frame(bt, bt.d_latent, true);
  launch_advance(bt.d_len, bt.d_bos, nullptr, bt.d_counter, B, 1, 0, c.stream, bt.d_active);
  if (copy_out) {
    cudaMemcpyAsync(hl, bt.d_latent, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(hg, bt.d_logit, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    if (bt.pcm16) cudaMemcpyAsync(hp, bt.d_pcm, (size_t)B * bt.frame_samples * sizeof(short), cudaMemcpyDeviceToHost, c.stream);
    else cudaMemcpyAsync(ha, bt.d_audio, (size_t)B * bt.frame_samples * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
  }
}

// Pipelined frame t >= 1 (parity = t & 1): FlowLM step t reads latent t-1 and writes latent t on the main stream
// while the Mimi decoder turns latent t-1 into audio on the second stream; host_io adds the noise upload and the
// result downloads (latent t, EOS logit t, audio t-1).
void pipelined_frame(Batch& bt, int parity, bool host_io) {
  Ctx& c = *bt.ctx;
  const int B = bt.B, L = c.cfg.latent_dim;
  float* lat_cur = parity ? bt.d_latent_b : bt.d_latent;
  float* lat_prev = parity ? bt.d_latent : bt.d_latent_b;
  cudaStream_t main = c.stream;
  cudaEventRecord(c.ev_fork, main);
  cudaStreamWaitEvent(c.stream2, c.ev_fork, 0);
  const bool alt = bt.async_staging && parity == 1;        // odd frames of an async-staged batch: second buffer set
  float *hn = alt ? bt.h2_noise : bt.h_noise, *hl = alt ? bt.h2_latent : bt.h_latent, *hg = alt ? bt.h2_logit : bt.h_logit,
        *ha = alt ? bt.h2_audio : bt.h_audio;
  short* hp = alt ? bt.h2_pcm : bt.h_pcm;
  // SM partition between the branches: the Mimi branch's persistent kernels take at most PTTS_MIMI_GRID SMs (default
  // 74 = one die's worth), so the latency-bound FlowLM chain always finds free SMs instead of queueing behind a
  // 148-CTA persistent GEMM.  Measured at batch 256: 17.9 k -> 19.2 k audio-s/s (60: 19.0 k, 88: 18.7 k, 100: 18.1 k).
  static const int mimi_grid = [] { const char* v = getenv("PTTS_MIMI_GRID"); return v ? atoi(v) : 74; }();
  // PTTS_INTERLEAVE: 0 = the two branches start together and run freely; 1 = the Mimi decoder is cut into one slice per
  // FlowLM layer and slice i starts only when the attention kernels of layer i have finished (the HBM-bound attention
  // gets the whole machine, the decoder fills the latency-bound GEMM chains between two attentions); 2 = additionally
  // attention i+1 waits for slice i (strict alternation)
  static const int interleave = [] { const char* v = getenv("PTTS_INTERLEAVE"); return v ? atoi(v) : 0; }();
  const int NL = c.cfg.n_layers;
  auto mimi_out = [&]() {
    if (!host_io) return;
    if (bt.pcm16) cudaMemcpyAsync(hp, bt.d_pcm, (size_t)B * bt.frame_samples * sizeof(short), cudaMemcpyDeviceToHost, c.stream2);
    else cudaMemcpyAsync(ha, bt.d_audio, (size_t)B * bt.frame_samples * sizeof(float), cudaMemcpyDeviceToHost, c.stream2);
  };
  if (interleave > 0 && bt.tc_mimi && bt.fw.tc && NL <= 16) {
    const int n_launch = mimi_tc_launches(bt);
    std::vector<int> cut(NL + 1);
    for (int i = 0; i <= NL; ++i) cut[i] = (int)((long long)i * n_launch / NL);
    if (NL == 6 && n_launch == 26) { const int hand[7] = {0, 6, 11, 16, 20, 23, 26}; cut.assign(hand, hand + 7); }   // ~equal time
    bt.fw.post_attn = [&](int i) {
      cudaEventRecord(c.ev_attn[i], main);
      cudaStreamWaitEvent(c.stream2, c.ev_attn[i], 0);
      c.stream = c.stream2;
      gemm_tc_set_grid_cap(mimi_grid);
      mimi_frame_tc(bt, lat_prev, 0, true, cut[i], cut[i + 1]);
      gemm_tc_set_grid_cap(0);
      if (i == NL - 1) mimi_out();
      c.stream = main;
      if (interleave > 1) cudaEventRecord(c.ev_slice[i], c.stream2);
    };
    bt.fw.pre_attn = [&](int i) {
      if (interleave > 1 && i > 0) cudaStreamWaitEvent(main, c.ev_slice[i - 1], 0);
    };
    if (host_io)
      cudaMemcpyAsync(bt.d_noise, hn, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, c.stream);
    flow_step(bt, host_io, 0, lat_prev, lat_cur);
    bt.fw.post_attn = nullptr;
    bt.fw.pre_attn = nullptr;
    cudaEventRecord(c.ev_join, c.stream2);
  } else {
    c.stream = c.stream2;                                   // every launcher below targets the Mimi branch
    gemm_tc_set_grid_cap(mimi_grid);
    g_launch_prio = c.prio_lo;
    mimi_frame(bt, lat_prev, true);
    gemm_tc_set_grid_cap(0);
    c.stream = main;
    mimi_out();
    cudaEventRecord(c.ev_join, c.stream2);
    if (host_io)
      cudaMemcpyAsync(bt.d_noise, hn, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, c.stream);
    g_launch_prio = c.prio_hi;
    flow_step(bt, host_io, 0, lat_prev, lat_cur);
    g_launch_prio = 0;
  }
  launch_advance(bt.d_len, bt.d_bos, nullptr, bt.d_counter, B, 1, 0, c.stream, bt.d_active);
  if (host_io) {
    cudaMemcpyAsync(hl, lat_cur, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpyAsync(hg, bt.d_logit, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, c.stream);
  }
  cudaStreamWaitEvent(main, c.ev_join, 0);
}

int restore_mimi_slot(Batch& t, int slot);
int restore_mimi_slots(Batch& t, const std::vector<int>& slots);

int run_pipelined_step(Batch& bt, bool host_io) {
  Ctx& c = *bt.ctx;
  const int B = bt.B, L = c.cfg.latent_dim;
  if (bt.frame_idx == 0) {
    // first frame: nothing to decode yet -> FlowLM step only (eager), writes latent 0 into the even buffer
    if (host_io)
      CU(cudaMemcpyAsync(bt.d_noise, bt.h_noise, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    flow_step(bt, host_io, 0, bt.d_latent, bt.d_latent);
    launch_advance(bt.d_len, bt.d_bos, nullptr, bt.d_counter, B, 1, 0, c.stream, bt.d_active);
    if (host_io) {
      CU(cudaMemcpyAsync(bt.h_latent, bt.d_latent, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
      CU(cudaMemcpyAsync(bt.h_logit, bt.d_logit, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
      if (bt.pcm16) {
        memset(bt.h_pcm, 0, (size_t)B * bt.frame_samples * sizeof(short));
      } else {
        CU(cudaMemsetAsync(bt.d_audio, 0, (size_t)B * bt.frame_samples * sizeof(float), c.stream));
        CU(cudaMemcpyAsync(bt.h_audio, bt.d_audio, (size_t)B * bt.frame_samples * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
      }
    }
  } else {
    const int parity = (int)(bt.frame_idx & 1);
    const int idx = parity * 2 + (host_io ? 1 : 0);
    if (!bt.pipe_graph[idx]) {
      const long long before = g_launches;
      cudaGraph_t graph;
      CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
      pipelined_frame(bt, parity, host_io);
      CU(cudaStreamEndCapture(c.stream, &graph));
      bt.pipe_graph_launches[idx] = g_launches - before;
      g_launches = before;
      // (per-node launch priorities only count when the graph is instantiated with this flag)
      CU(cudaGraphInstantiate(&bt.pipe_graph[idx], graph, g_launch_prio_on ? cudaGraphInstantiateFlagUseNodePriority : 0));
      CU(cudaGraphDestroy(graph));
    }
    CU(cudaGraphLaunch(bt.pipe_graph[idx], c.stream));
    g_launches += bt.pipe_graph_launches[idx];
  }
  if (!bt.pending_mimi_reset.empty()) {
    // slots re-initialised since the previous frame: this frame's Mimi branch has just consumed the old utterance's last
    // latent with the old state; the NEXT frame decodes the new utterance's first latent and needs the fresh state
    RET(restore_mimi_slots(bt, bt.pending_mimi_reset));
    bt.pending_mimi_reset.clear();
  }
  bt.frame_idx += 1;
  for (int b = 0; b < bt.B; ++b) bt.h_len[b] += bt.h_active[b];
  return 0;
}

int ensure_step_graph(Batch& bt, bool host_noise, bool copy_out) {
  const int idx = (host_noise ? 2 : 0) + (copy_out ? 1 : 0);
  if (bt.step_graph[idx]) return 0;
  Ctx& c = *bt.ctx;
  const long long before = g_launches;
  cudaGraph_t graph;
  CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
  full_step(bt, host_noise, copy_out);
  CU(cudaStreamEndCapture(c.stream, &graph));
  bt.step_graph_launches[idx] = g_launches - before;
  g_launches = before;
  CU(cudaGraphInstantiate(&bt.step_graph[idx], graph, 0));
  CU(cudaGraphDestroy(graph));
  return 0;
}

int run_step(Batch& bt, bool host_noise, bool copy_out) {
  RET(ensure_step_graph(bt, host_noise, copy_out));
  const int idx = (host_noise ? 2 : 0) + (copy_out ? 1 : 0);
  CU(cudaGraphLaunch(bt.step_graph[idx], bt.ctx->stream));
  g_launches += bt.step_graph_launches[idx];
  for (int b = 0; b < bt.B; ++b) bt.h_len[b] += bt.h_active[b];
  return 0;
}

// sequential (non-pipelined) frame with host I/O through the second staging set
int run_step_alt(Batch& bt) {
  Ctx& c = *bt.ctx;
  if (!bt.step_graph_alt) {
    const long long before = g_launches;
    cudaGraph_t graph;
    CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
    full_step(bt, true, true, true);
    CU(cudaStreamEndCapture(c.stream, &graph));
    bt.step_graph_alt_launches = g_launches - before;
    g_launches = before;
    CU(cudaGraphInstantiate(&bt.step_graph_alt, graph, 0));
    CU(cudaGraphDestroy(graph));
  }
  CU(cudaGraphLaunch(bt.step_graph_alt, c.stream));
  g_launches += bt.step_graph_alt_launches;
  for (int b = 0; b < bt.B; ++b) bt.h_len[b] += bt.h_active[b];
  return 0;
}

// Per-sequence Mimi streaming state as (base, per-slot stride, bytes) pieces: the KV ring slices, the upsample conv's
// previous column, the ring offset, the carried rows of every streaming conv and the output-conv boundary partials.
struct StatePiece { char* base; size_t stride, bytes; };
std::vector<StatePiece> mimi_state_pieces(Batch& t) {
  Ctx& c = *t.ctx;
  const ptts_config& g = c.cfg;
  std::vector<StatePiece> v;
  const size_t esz = c.bf16 ? 2 : 4;
  const size_t slice = (size_t)g.mimi_heads * g.mimi_context * kHeadDim * esz;
  for (int l = 0; l < g.mimi_layers; ++l)
    for (int kv = 0; kv < 2; ++kv)
      v.push_back({(char*)t.ring + ((size_t)l * t.ring_layer_stride + (size_t)kv * t.ring_kv_stride) * esz, slice, slice});
  v.push_back({(char*)t.d_zprev, (size_t)g.seanet_dim * 4, (size_t)g.seanet_dim * 4});
  v.push_back({(char*)t.d_mimi_off, 4, 4});
  for (const ShiftEntry& e : t.h_shift)
    v.push_back({(char*)e.buf, (size_t)e.bs * e.esz, (size_t)e.rows * e.C * e.esz});
  if (t.d_bnd) v.push_back({(char*)t.d_bnd, (size_t)(t.frame_samples / 128 + 1) * 16, 16});
  return v;
}

// Mimi streaming state of the given slots back to the post-warm-up template (or zero), stream-ordered, ONE launch for
// all pieces and slots (an admission round of the continuous scheduler re-initialises ~16 slots x ~25 buffers)
int restore_mimi_slots(Batch& t, const std::vector<int>& slots) {
  Ctx& c = *t.ctx;
  if (slots.empty()) return 0;
  if (!t.d_state_pieces) {
    auto pieces = mimi_state_pieces(t);
    std::vector<StatePieceDev> dev;
    size_t off = 0;
    for (auto& p : pieces) {
      if ((p.bytes & 3) || (p.stride & 3)) return fail(PTTS_ERR_STATE, "state piece of %zu bytes is not a multiple of 4", p.bytes);
      dev.push_back({p.base, (unsigned long long)p.stride, (unsigned long long)p.bytes, (unsigned long long)off});
      off += (p.bytes + 15) & ~(size_t)15;
    }
    t.n_state_pieces = (int)dev.size();
    RET(t.dalloc((void**)&t.d_state_pieces, dev.size() * sizeof(StatePieceDev)));
    RET(t.dalloc((void**)&t.d_restore_slots, (size_t)t.B * sizeof(int)));
    CU(cudaMemcpyAsync(t.d_state_pieces, dev.data(), dev.size() * sizeof(StatePieceDev), cudaMemcpyHostToDevice, c.stream));
    CU(cudaStreamSynchronize(c.stream));     // dev is a local
  }
  // (a pageable host -> device copy of a few bytes has been staged by the time cudaMemcpyAsync returns)
  CU(cudaMemcpyAsync(t.d_restore_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  launch_restore_state(t.d_state_pieces, t.n_state_pieces, t.d_restore_slots, (int)slots.size(),
                       t.has_tpl ? (const char*)t.mimi_tpl : nullptr, c.stream);
  return 0;
}
int restore_mimi_slot(Batch& t, int slot) { return restore_mimi_slots(t, std::vector<int>{slot}); }

// captured frame graphs bake pointers and modes in (staging sets, cascade prefix length, PCM output): drop them all
void drop_graphs(Batch& t) {
  for (auto& g : t.step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  if (t.step_graph_alt) { cudaGraphExecDestroy(t.step_graph_alt); t.step_graph_alt = nullptr; }
  for (auto& g : t.pipe_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  for (auto& kv : t.mimi_graphs) cudaGraphExecDestroy(kv.second.first);
  t.mimi_graphs.clear();
}

int check_step_ready(Batch& bt) {
  if (!bt.prefilled) return fail(PTTS_ERR_STATE, "ptts_batch_step called before ptts_batch_prefill_text");
  for (int b = 0; b < bt.B; ++b)
    if (bt.h_active[b] && bt.h_len[b] + 1 > bt.max_len[b])
      return fail(PTTS_ERR_STATE, "sequence %d would exceed its max_len %d", b, bt.max_len[b]);
  return 0;
}

}  // namespace

// ============================================================ extern "C" ====================================
extern "C" {

static void batch_free(ptts_batch* bt);

int32_t ptts_abi_version(void) { return PTTS_ABI_VERSION; }
const char* ptts_last_error(void) { return g_err.c_str(); }

int32_t ptts_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int32_t ptts_ctx_create(int32_t device, const ptts_config* cfg, ptts_ctx** out) {
  if (!cfg || !out) return fail(PTTS_ERR_INVALID, "null argument");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(PTTS_ERR_CUDA, "no CUDA device is available; libptts_b200 has no CPU fallback");
  if (device < 0 || device >= n) return fail(PTTS_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  if (cfg->d_model != cfg->n_heads * kHeadDim || cfg->mimi_d != cfg->mimi_heads * kHeadDim)
    return fail(PTTS_ERR_INVALID, "kernels are specialised for 64-wide attention heads");
  if (cfg->d_model > 1024 || cfg->mimi_d > 1024 || cfg->flow_dim > 1024)
    return fail(PTTS_ERR_INVALID, "row-norm kernels support widths up to 1024");
  if (cfg->n_ratios < 1 || cfg->n_ratios > 8 || cfg->lsd_decode_steps < 1 || cfg->latent_dim % 8 ||
      cfg->kv_pool_tokens < kPageTokens || cfg->upsample_stride > 16)
    return fail(PTTS_ERR_INVALID, "unsupported configuration");
  if (cfg->n_heads * 32 > 1024) return fail(PTTS_ERR_INVALID, "too many heads");
  CU(cudaSetDevice(device));
  auto c = std::make_unique<ptts_ctx>();
  c->device = device;
  c->cfg = *cfg;
  c->bf16 = cfg->precision == PTTS_BF16;
  {
    const char* fs = getenv("PTTS_FORCE_SIMT");
    c->force_simt = fs && fs[0] == '1';
    // programmatic dependent launch measured slower on this chain (1.586 vs 1.547 ms/frame at batch 256):
    // graph launch gaps are already ~1 us, so it stays opt-in
    const char* np = getenv("PTTS_PDL");
    g_pdl_on = np && np[0] == '1';
  }
  {
    // PTTS_PRIO=1: the main stream (FlowLM branch of the pipelined graph: a chain of short latency-bound kernels) gets the
    // highest priority and the Mimi branch the lowest, so that the block scheduler dispatches a waiting FlowLM grid before
    // the remaining CTAs of a large Mimi grid (captured kernel nodes inherit the priority of their stream)
    const char* pv = getenv("PTTS_PRIO");
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    c->prio_hi = hi; c->prio_lo = lo;
    if (pv && pv[0] == '2') g_launch_prio_on = true;     // per-launch priorities, set around the two branches in pipelined_frame
    if (pv && pv[0] == '1') {
      CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
      CU(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, lo));
    } else {
      CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    }
  }
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (int i = 0; i < 16; ++i) {
    CU(cudaEventCreateWithFlags(&c->ev_attn[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_slice[i], cudaEventDisableTiming));
  }
  gemm_tc_init();
  *out = c.release();
  return 0;
}

void ptts_ctx_destroy(ptts_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (ptts_batch* p : c->parked) batch_free(p);
  c->parked.clear();
  free_flow_work(c->prefill_work);
  for (void* p : c->allocs) cudaFree(p);
  if (c->l2_scratch) cudaFree(c->l2_scratch);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  for (int i = 0; i < 16; ++i) { cudaEventDestroy(c->ev_attn[i]); cudaEventDestroy(c->ev_slice[i]); }
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
}

int32_t ptts_load_weight(ptts_ctx* c, const char* name, int32_t dtype, int32_t ndim, const int64_t* shape,
                         const void* data) {
  if (!c || !name || !shape || !data) return fail(PTTS_ERR_INVALID, "null argument");
  if (c->finalized) return fail(PTTS_ERR_STATE, "weights already finalized");
  const std::string nm(name);
  if (nm.rfind("flow_lm.", 0) != 0 && nm.rfind("mimi.", 0) != 0) return 1;
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  t.data.resize((size_t)n);
  if (dtype == PTTS_DT_F32) memcpy(t.data.data(), data, (size_t)n * 4);
  else if (dtype == PTTS_DT_BF16) for (int64_t i = 0; i < n; ++i) t.data[i] = bf2f(((const uint16_t*)data)[i]);
  else if (dtype == PTTS_DT_F16) for (int64_t i = 0; i < n; ++i) t.data[i] = h2f(((const uint16_t*)data)[i]);
  else return fail(PTTS_ERR_INVALID, "unsupported dtype %d for '%s'", dtype, name);
  c->host[nm] = std::move(t);
  return 0;
}

int32_t ptts_finalize_weights(ptts_ctx* c) {
  if (!c) return fail(PTTS_ERR_INVALID, "null ctx");
  if (c->finalized) return fail(PTTS_ERR_STATE, "weights already finalized");
  CU(cudaSetDevice(c->device));
  return finalize(*c);
}

const char* ptts_unused_weights(ptts_ctx* c) { return (c && c->finalized) ? c->unused_report.c_str() : ""; }

int32_t ptts_has_voice_cloning(ptts_ctx* c) { return (c && c->finalized && c->has_encoder) ? 1 : 0; }

// Voice cloning: waveform -> FlowLM conditioning (MimiModel.encode_to_latent + the speaker projection,
// models/mimi.py:77-85, models/tts_model.py:271-276).  One-off per voice, fp32 throughout, CUDA-core kernels:
// every conv is the multi-tap linear operator (a stride-s conv with kernel 2s is a 2-tap GEMM over rows of s*C
// channels: the "space to depth" view of the contiguous [T][C] buffer), the encoder transformer is the
// non-streaming windowed form of the Mimi transformer.
int32_t ptts_encode_audio(ptts_ctx* cp, const float* audio, int64_t n_samples, float* out_cond, int32_t max_frames,
                          int32_t* n_frames_out) {
  if (!cp || !audio || !out_cond || !n_frames_out) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *cp;
  if (!c.finalized) return fail(PTTS_ERR_STATE, "weights are not finalized");
  if (!c.has_encoder) return fail(PTTS_ERR_STATE, "the checkpoint was loaded without the Mimi encoder (no voice cloning)");
  if (n_samples <= 0) return fail(PTTS_ERR_INVALID, "empty audio");
  CU(cudaSetDevice(c.device));
  const ptts_config& g = c.cfg;
  const int S = g.upsample_stride, SD = g.seanet_dim, MD = g.mimi_d, D = g.d_model;
  long long hop = 1;
  for (int r = 0; r < g.n_ratios; ++r) hop *= g.ratios[r];
  const long long frame = hop * S;
  const long long n_frames = (n_samples + frame - 1) / frame;        // zero-padded at the end to whole frames
  if (n_frames > max_frames) return fail(PTTS_ERR_INVALID, "audio gives %lld frames, the output buffer holds %d", n_frames, max_frames);
  const long long T = n_frames * frame;
  const int k0 = g.kernel_size, rk = g.res_kernel_size, lk = g.last_kernel_size;
  std::vector<void*> tmp;
  auto alloc = [&](float** p, size_t n, bool zero) -> int {
    CU(cudaMalloc((void**)p, n * sizeof(float)));
    tmp.push_back(*p);
    if (zero) CU(cudaMemsetAsync(*p, 0, n * sizeof(float), c.stream));
    return 0;
  };
  auto cleanup = [&](int rc) {
    cudaStreamSynchronize(c.stream);
    for (void* p : tmp) cudaFree(p);
    return rc;
  };
#define ENC(x) do { int rc_ = (x); if (rc_ != 0) return cleanup(rc_); } while (0)
  float* xpad;
  ENC(alloc(&xpad, (size_t)(T + k0 - 1), true));
  if (cudaMemcpyAsync(xpad + (k0 - 1), audio, (size_t)n_samples * sizeof(float), cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
    return cleanup(fail(PTTS_ERR_CUDA, "audio upload failed"));
  // level 0: conv0 output behind the rk-1 zero rows of the first resblock conv
  long long Tr = T;
  int C = g.n_filters;
  float* h;
  ENC(alloc(&h, (size_t)(Tr + rk - 1) * C, true));
  launch_enc_conv0(xpad, c.enc0_w, c.enc0_b, h + (size_t)(rk - 1) * C, Tr, C, k0, c.stream);
  for (int r = 0; r < g.n_ratios; ++r) {
    auto& st = c.enc_stages[r];
    const int s = st.stride, Ch = C / g.compress;
    float *u, *h2, *hn;
    ENC(alloc(&u, (size_t)Tr * Ch, false));
    ENC(alloc(&h2, (size_t)(Tr + s) * C, true));
    {   // u = conv_k3(ELU(h))
      LinearParams p{};
      p.tag = "enc.r3"; p.A = h; p.a_bs = 0; p.a_rs = C; p.nb = 1; p.T = (int)Tr; p.taps = rk; p.C = C;
      p.a_pro = ACT_ELU; p.Y = u; p.y_rs = Ch; p.out_scale = 1.f;
      run_linear(c, st.r3, p);
    }
    {   // h2 = h + conv_k1(ELU(u)), written behind the s zero rows of the strided conv
      LinearParams p = rows_linear(u, (int)Tr, Ch, h2 + (size_t)s * C, C, "enc.r1");
      p.a_pro = ACT_ELU;
      p.res = h + (size_t)(rk - 1) * C; p.res_rs = C;
      run_linear(c, st.r1, p);
    }
    const long long Tn = Tr / s;
    const int Cn = 2 * C;
    const int pad_next = ((r + 1 < g.n_ratios) ? rk : lk) - 1;
    ENC(alloc(&hn, (size_t)(Tn + pad_next) * Cn, true));
    {   // stride-s conv with kernel 2s = 2-tap GEMM over rows of s*C channels
      LinearParams p{};
      p.tag = "enc.down"; p.A = h2; p.a_bs = 0; p.a_rs = (long long)s * C; p.nb = 1; p.T = (int)Tn; p.taps = 2; p.C = s * C;
      p.a_pro = ACT_ELU; p.Y = hn + (size_t)pad_next * Cn; p.y_rs = Cn; p.out_scale = 1.f;
      run_linear(c, st.down, p);
    }
    h = hn; Tr = Tn; C = Cn;
  }
  const int T3 = (int)Tr;                     // encoder steps (S per frame)
  float *xd, *hb, *qkv, *att, *ff, *lat, *cond;
  ENC(alloc(&xd, (size_t)(T3 + S) * SD, false));
  float* x = xd + (size_t)S * SD;
  {   // last conv: ELU, kernel lk
    LinearParams p{};
    p.tag = "enc.last"; p.A = h; p.a_bs = 0; p.a_rs = C; p.nb = 1; p.T = T3; p.taps = lk; p.C = C;
    p.a_pro = ACT_ELU; p.Y = x; p.y_rs = SD; p.out_scale = 1.f;
    run_linear(c, c.enc_last, p);
  }
  ENC(alloc(&hb, (size_t)T3 * MD, false));
  ENC(alloc(&qkv, (size_t)T3 * 3 * MD, false));
  ENC(alloc(&att, (size_t)T3 * MD, false));
  ENC(alloc(&ff, (size_t)T3 * g.mimi_ffn, false));
  for (int i = 0; i < g.mimi_layers; ++i) {
    auto& l = c.el[i];
    rows_norm(c, x, T3, MD, l.ln1w, l.ln1b, 1e-5f, hb, nullptr, nullptr, 0, nullptr, nullptr, 0);
    run_linear(c, l.qkv, rows_linear(hb, T3, MD, qkv, 3 * MD, "enc.qkv"));
    launch_enc_attention(qkv, att, c.freqs_mimi, T3, g.mimi_heads, g.mimi_context, c.stream);
    LinearParams o = rows_linear(att, T3, MD, x, MD, "enc.out");
    o.col_scale = l.ls1; o.res = x; o.res_rs = MD;
    run_linear(c, l.out, o);
    rows_norm(c, x, T3, MD, l.ln2w, l.ln2b, 1e-5f, hb, nullptr, nullptr, 0, nullptr, nullptr, 0);
    LinearParams f1 = rows_linear(hb, T3, MD, ff, g.mimi_ffn, "enc.ff1");
    f1.act = ACT_GELU;
    run_linear(c, l.ff1, f1);
    LinearParams f2 = rows_linear(ff, T3, g.mimi_ffn, x, MD, "enc.ff2");
    f2.col_scale = l.ls2; f2.res = x; f2.res_rs = MD;
    run_linear(c, l.ff2, f2);
  }
  // downsample: stride S, kernel 2S, replicate padding (S copies of the first step), no bias
  launch_replicate_row(xd, x, S, SD, c.stream);
  ENC(alloc(&lat, (size_t)n_frames * SD, false));
  {
    LinearParams p{};
    p.tag = "enc.downsample"; p.A = xd; p.a_bs = 0; p.a_rs = (long long)S * SD; p.nb = 1; p.T = (int)n_frames; p.taps = 2;
    p.C = S * SD; p.Y = lat; p.y_rs = SD; p.out_scale = 1.f;
    run_linear(c, c.enc_down, p);
  }
  ENC(alloc(&cond, (size_t)n_frames * D, false));
  run_linear(c, c.speaker_proj, rows_linear(lat, (int)n_frames, SD, cond, D, "enc.speaker_proj"));
  if (cudaMemcpyAsync(out_cond, cond, (size_t)n_frames * D * sizeof(float), cudaMemcpyDeviceToHost, c.stream) != cudaSuccess)
    return cleanup(fail(PTTS_ERR_CUDA, "conditioning download failed"));
  if (cudaStreamSynchronize(c.stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
    return cleanup(fail(PTTS_ERR_CUDA, "encoder kernels failed: %s", cudaGetErrorString(cudaGetLastError())));
#undef ENC
  *n_frames_out = (int32_t)n_frames;
  return cleanup(0);
}

int32_t ptts_voice_create(ptts_ctx* c, const float* cond, int32_t n_frames) {
  if (!c || !cond || n_frames <= 0) return fail(PTTS_ERR_INVALID, "bad voice prompt");
  if (!c->finalized) return fail(PTTS_ERR_STATE, "finalize weights first");
  CU(cudaSetDevice(c->device));
  const int D = c->cfg.d_model;
  Voice v;
  v.len = n_frames;
  const int np = (n_frames + kPageTokens - 1) / kPageTokens;
  RET(take_pages(*c, np, &v.pages));
  RET(alloc_flow_work(*c, c->prefill_work, n_frames));
  if (c->prefill_work.tc) RET(build_flow_plans(*c, c->prefill_work, n_frames));
  std::vector<int> seq(n_frames, 0), pos(n_frames);
  for (int i = 0; i < n_frames; ++i) pos[i] = i;
  int *d_seq, *d_pos, *d_pt;
  CU(cudaMalloc(&d_seq, n_frames * 4));
  CU(cudaMalloc(&d_pos, n_frames * 4));
  CU(cudaMalloc(&d_pt, np * 4));
  CU(cudaMemcpyAsync(d_seq, seq.data(), n_frames * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_pos, pos.data(), n_frames * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_pt, v.pages.data(), np * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->prefill_work.x, cond, (size_t)n_frames * D * 4, cudaMemcpyHostToDevice, c->stream));
  int* d_rows;                                         // {row0 = 0, row1 = n_frames, pos0 = 0}
  const int rows3[3] = {0, n_frames, 0};
  CU(cudaMalloc(&d_rows, 3 * 4));
  CU(cudaMemcpyAsync(d_rows, rows3, 3 * 4, cudaMemcpyHostToDevice, c->stream));
  c->prefill_work.seq_row0 = d_rows; c->prefill_work.seq_pos0 = d_rows + 2;
  c->prefill_work.n_seq = 1; c->prefill_work.max_rows_per_seq = n_frames;
  flow_layers(*c, c->prefill_work, n_frames, d_seq, d_pos, d_pt, np, (long long)n_frames * (n_frames + 1) / 2);
  c->prefill_work.seq_row0 = c->prefill_work.seq_pos0 = nullptr; c->prefill_work.n_seq = 0;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  cudaFree(d_seq); cudaFree(d_pos); cudaFree(d_pt); cudaFree(d_rows);
  v.alive = true;
  for (size_t i = 0; i < c->voices.size(); ++i)
    if (!c->voices[i].alive && c->voices[i].pages.empty()) {
      c->voices[i] = std::move(v);
      return (int32_t)i;
    }
  c->voices.push_back(std::move(v));
  return (int32_t)c->voices.size() - 1;
}

int32_t ptts_voice_destroy(ptts_ctx* c, int32_t id) {
  if (!c || id < 0 || id >= (int)c->voices.size() || !c->voices[id].alive)
    return fail(PTTS_ERR_INVALID, "unknown voice id %d", id);
  Voice& v = c->voices[id];
  if (v.refs > 0) {        // live slots still read its prefix pages: freed by the last voice_unref
    v.alive = false;
    v.doomed = true;
    return 0;
  }
  for (int p : v.pages) c->free_pages.push_back(p);
  v = Voice{};
  return 0;
}

int32_t ptts_voice_length(ptts_ctx* c, int32_t id) {
  if (!c || id < 0 || id >= (int)c->voices.size() || !c->voices[id].alive)
    return fail(PTTS_ERR_INVALID, "unknown voice id %d", id);
  return c->voices[id].len;
}

static int batch_create_impl(ptts_ctx* c, int32_t B, const int32_t* voice_ids, const int32_t* max_len,
                             std::unique_ptr<ptts_batch>& bt);

int32_t ptts_batch_create(ptts_ctx* c, int32_t B, const int32_t* voice_ids, const int32_t* max_len,
                          ptts_batch** out) {
  if (!c || !voice_ids || !max_len || !out || B <= 0) return fail(PTTS_ERR_INVALID, "bad batch arguments");
  *out = nullptr;
  std::unique_ptr<ptts_batch> bt;
  const int r = batch_create_impl(c, B, voice_ids, max_len, bt);
  if (r < 0) {
    const std::string keep = g_err;
    if (bt) ptts_batch_destroy(bt.release());
    g_err = keep;
    return r;
  }
  *out = bt.release();
  return 0;
}

// (re)initialise the per-utterance state of an allocated batch: page tables (full voice-prefix pages shared,
// the partial tail page copied), lengths, BOS flags, zeroed Mimi streaming state
static int batch_init_state(ptts_batch& t, const int32_t* voice_ids, const int32_t* max_len) {
  ptts_ctx* c = t.ctx;
  const int B = t.B, maxp = t.max_pages;
  t.voice_ids.assign(voice_ids, voice_ids + B);
  t.max_len.assign(max_len, max_len + B);
  t.h_len.resize(B);
  t.prefilled = false;
  t.frame_idx = 0;
  t.pipelined = false;
  t.async_staging = false;
  t.pcm16 = false;
  t.pending_mimi_reset.clear();
  if (t.graphs_pcm || t.graphs_async) {
    // a recycled arena whose graphs were captured with the PCM output on, or with odd frames going through the second
    // staging set: both are baked into the copy nodes and this batch starts with neither
    drop_graphs(t);
    t.graphs_pcm = false;
    t.graphs_async = false;
  }
  std::vector<int> pt((size_t)B * maxp, 0), src, dst;
  t.slot_pages.assign(B, {});
  t.slot_voice.assign(B, -1);
  t.h_active.assign(B, 1);
  t.has_tpl = false;
  for (int b = 0; b < B; ++b) {
    const Voice& v = c->voices[voice_ids[b]];
    t.h_len[b] = v.len;
    const int full = v.len / kPageTokens;
    const int need_pages = (max_len[b] + kPageTokens - 1) / kPageTokens;
    for (int i = 0; i < full; ++i) pt[(size_t)b * maxp + i] = v.pages[i];
    std::vector<int> mine;
    RET(take_pages(*c, need_pages - full, &mine));
    for (int i = full; i < need_pages; ++i) pt[(size_t)b * maxp + i] = mine[i - full];
    t.slot_pages[b] = mine;
    t.slot_voice[b] = voice_ids[b];
    c->voices[voice_ids[b]].refs += 1;
    if (v.len % kPageTokens) {
      src.push_back(v.pages[full]);
      dst.push_back(mine[0]);
    }
  }
  t.h_page_table = pt;
  CU(cudaMemcpyAsync(t.d_page_table, pt.data(), pt.size() * 4, cudaMemcpyHostToDevice, c->stream));
  launch_fill_u32((unsigned*)t.d_active, 1u, B, c->stream);
  if (!src.empty()) {
    CU(cudaMemcpyAsync(t.d_cp_src, src.data(), src.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(t.d_cp_dst, dst.data(), dst.size() * 4, cudaMemcpyHostToDevice, c->stream));
    launch_copy_pages(c->pool, c->bf16, c->layer_stride, c->page_stride, c->cfg.n_layers, t.d_cp_src, t.d_cp_dst,
                      (int)src.size(), c->stream);
  }
  // cascade attention: every sequence shares the same voice and the prefix fits the tensor-core prefix kernel
  {
    bool same = c->bf16 && t.fw.tc && B >= 32 && getenv("PTTS_NO_CASCADE") == nullptr;
    for (int b = 1; b < B && same; ++b) same = voice_ids[b] == voice_ids[0];
    const Voice& v0 = c->voices[voice_ids[0]];
    t.fw.prefix_len = 0;
    if (same && v0.len <= 128 && v0.len >= 16) {
      if (!t.fw.d_prefix_pages) {
        CU(cudaMalloc((void**)&t.fw.d_prefix_pages, 8 * sizeof(int)));
        CU(cudaMalloc((void**)&t.fw.prefix_part, (size_t)B * c->cfg.n_heads * 66 * sizeof(float)));
        t.fw.pflags_stride = 1 + ((B + 15) / 16) * c->cfg.n_heads;
        CU(cudaMalloc((void**)&t.fw.pflags, (size_t)c->cfg.n_layers * t.fw.pflags_stride * sizeof(int)));
      }
      CU(cudaMemsetAsync(t.fw.pflags, 0, (size_t)c->cfg.n_layers * t.fw.pflags_stride * sizeof(int), c->stream));
      CU(cudaMemcpyAsync(t.fw.d_prefix_pages, v0.pages.data(), v0.pages.size() * sizeof(int), cudaMemcpyHostToDevice,
                         c->stream));
      t.fw.prefix_len = v0.len;
    }
    if (t.cascade_len != t.fw.prefix_len) {      // the prefix length is baked into the captured graphs
      for (auto& g : t.step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
      if (t.step_graph_alt) { cudaGraphExecDestroy(t.step_graph_alt); t.step_graph_alt = nullptr; }
      for (auto& g : t.pipe_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
      t.cascade_len = t.fw.prefix_len;
    }
  }
  for (auto& z : t.zero_list) CU(cudaMemsetAsync(z.first, 0, z.second, c->stream));
  CU(cudaMemcpyAsync(t.d_len, t.h_len.data(), B * 4, cudaMemcpyHostToDevice, c->stream));
  launch_fill_u32((unsigned*)t.d_bos, 1u, B, c->stream);
  const unsigned long long cs[2] = {0ull, t.seed};
  CU(cudaMemcpyAsync(t.d_counter, cs, 16, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));    // pt/src/dst/cs are stack or local vectors
  CU(cudaGetLastError());
  return 0;
}

static int batch_create_impl(ptts_ctx* c, int32_t B, const int32_t* voice_ids, const int32_t* max_len,
                             std::unique_ptr<ptts_batch>& bt) {
  if (!c->finalized) return fail(PTTS_ERR_STATE, "finalize weights first");
  if (c->cfg.max_batch > 0 && B > c->cfg.max_batch) return fail(PTTS_ERR_INVALID, "batch %d exceeds max_batch", B);
  CU(cudaSetDevice(c->device));
  const ptts_config& g = c->cfg;
  int maxp = 1;
  for (int b = 0; b < B; ++b) {
    const int v = voice_ids[b];
    if (v < 0 || v >= (int)c->voices.size() || !c->voices[v].alive) return fail(PTTS_ERR_INVALID, "unknown voice id %d", v);
    if (max_len[b] < c->voices[v].len) return fail(PTTS_ERR_INVALID, "max_len[%d] shorter than the voice prefix", b);
    maxp = std::max(maxp, (max_len[b] + kPageTokens - 1) / kPageTokens);
  }
  // recycle the arena (device buffers, tensor maps, captured graphs) of a destroyed batch of the same shape
  for (size_t i = 0; i < c->parked.size(); ++i) {
    ptts_batch* p = c->parked[i];
    if (p->B == B && p->max_pages >= maxp && p->no_tail_env == sn_tail_disabled()) {
      c->parked.erase(c->parked.begin() + i);
      bt.reset(p);
      return batch_init_state(*bt, voice_ids, max_len);
    }
  }
  bt = std::make_unique<ptts_batch>();
  bt->ctx = c;
  bt->B = B;
  bt->max_pages = (maxp + 7) / 8 * 8;
  Batch& t = *bt;
  RET(t.dalloc((void**)&t.d_page_table, (size_t)B * t.max_pages * 4));
  RET(t.dalloc((void**)&t.d_cp_src, B * 4));
  RET(t.dalloc((void**)&t.d_cp_dst, B * 4));
  RET(t.dalloc((void**)&t.d_len, B * 4));
  RET(t.dalloc((void**)&t.d_bos, B * 4));
  RET(t.dalloc((void**)&t.d_active, B * 4));
  RET(t.dalloc((void**)&t.d_mimi_off, B * 4));
  RET(t.dalloc((void**)&t.d_frame_idx, 4));
  RET(t.dalloc((void**)&t.d_counter, 16));
  t.zero_list.push_back({t.d_mimi_off, (size_t)B * 4});
  t.zero_list.push_back({t.d_frame_idx, 4});
  RET(alloc_flow_work(*c, t.fw, B));
  if (t.fw.tc) RET(build_flow_plans(*c, t.fw, B));
  t.tc_head = t.fw.tc;
  t.tc_mimi = want_tc(*c, B * g.upsample_stride);
  const int D = g.d_model, L = g.latent_dim, fd = g.flow_dim;
  auto fz = [&](float** p, size_t n) -> int {
    RET(t.dalloc((void**)p, n * 4));
    t.zero_list.push_back({*p, n * 4});
    return 0;
  };
  RET(fz(&t.d_noise, (size_t)B * L));
  RET(fz(&t.d_x, (size_t)B * L));
  RET(fz(&t.d_latent, (size_t)B * L));
  RET(fz(&t.d_latent_b, (size_t)B * L));
  RET(fz(&t.d_zero_lat, (size_t)B * L));
  RET(fz(&t.d_c, (size_t)B * D));
  RET(fz(&t.d_logit, B));
  RET(fz(&t.d_ada, (size_t)B * c->n_ada));
  RET(fz(&t.d_x1, (size_t)B * fd));
  if (!t.tc_head) {
    RET(fz(&t.d_sy, (size_t)B * fd));
    RET(fz(&t.d_hh, (size_t)B * fd));
    RET(fz(&t.d_u, (size_t)B * fd));
  }
  RET(fz(&t.d_v, (size_t)B * L));
  // Mimi state + scratch.  Every conv input buffer is [B][taps-1 state rows + T rows][C], zero-initialised
  // (= the reference's zero `previous` / `partial`, modules/conv.py:113-119,176-180).
  const int T = g.upsample_stride, MD = g.mimi_d, SD = g.seanet_dim;
  t.T0 = T;
  RET(fz(&t.d_zprev, (size_t)B * SD));
  RET(fz(&t.d_mqkv, (size_t)B * T * 3 * MD));
  RET(fz(&t.d_mqrot, (size_t)B * T * MD));
  std::vector<ShiftEntry> sh;
  int Tin = T;
  if (t.tc_mimi) {
    for (int r = 0; r < g.n_ratios; ++r) Tin *= c->stages[r].stride;
  } else {
  RET(fz(&t.d_c0, (size_t)B * (T + g.kernel_size - 1) * SD));
  RET(fz(&t.d_mh, (size_t)B * T * MD));
  RET(fz(&t.d_matt, (size_t)B * T * MD));
  RET(fz(&t.d_mff, (size_t)B * T * g.mimi_ffn));
  sh.push_back({t.d_c0, (long long)(T + g.kernel_size - 1) * SD, T, g.kernel_size - 1, SD, 4});
  t.sb.resize(g.n_ratios);
  for (int r = 0; r < g.n_ratios; ++r) {
    auto& st = c->stages[r];
    auto& b = t.sb[r];
    b.T_in = Tin;
    b.T_out = Tin * st.stride;
    const int rk = g.res_kernel_size;
    RET(fz(&b.ct_in, (size_t)B * (Tin + 1) * st.c_in));
    RET(fz(&b.r_in, (size_t)B * (b.T_out + rk - 1) * st.c_out));
    RET(fz(&b.hid, (size_t)B * b.T_out * st.hidden));
    sh.push_back({b.ct_in, (long long)(Tin + 1) * st.c_in, Tin, 1, st.c_in, 4});
    if (rk > 1) sh.push_back({b.r_in, (long long)(b.T_out + rk - 1) * st.c_out, b.T_out, rk - 1, st.c_out, 4});
    Tin = b.T_out;
  }
  RET(fz(&t.d_fin, (size_t)B * (Tin + c->fin_taps - 1) * c->fin_c));
  if (c->fin_taps > 1) sh.push_back({t.d_fin, (long long)(Tin + c->fin_taps - 1) * c->fin_c, Tin, c->fin_taps - 1, c->fin_c, 4});
  t.n_shift = (int)sh.size();
  t.h_shift = sh;
  RET(t.dalloc((void**)&t.d_shift, sh.size() * sizeof(ShiftEntry)));
  CU(cudaMemcpyAsync(t.d_shift, sh.data(), sh.size() * sizeof(ShiftEntry), cudaMemcpyHostToDevice, c->stream));
  }
  t.frame_samples = Tin;
  RET(fz(&t.d_audio, (size_t)B * Tin));
  RET(build_batch_tc(t));
  t.ring_kv_stride = (long long)B * g.mimi_heads * g.mimi_context * kHeadDim;
  t.ring_layer_stride = 2 * t.ring_kv_stride;
  const size_t ring_bytes = (size_t)g.mimi_layers * t.ring_layer_stride * (c->bf16 ? 2 : 4);
  RET(t.dalloc(&t.ring, ring_bytes));
  t.zero_list.push_back({t.ring, ring_bytes});
  CU(cudaMallocHost((void**)&t.h_noise, (size_t)B * L * 4));
  CU(cudaMallocHost((void**)&t.h_latent, (size_t)B * L * 4));
  CU(cudaMallocHost((void**)&t.h_logit, (size_t)B * 4));
  CU(cudaMallocHost((void**)&t.h_audio, (size_t)B * Tin * 4));
  return batch_init_state(t, voice_ids, max_len);
}

// give the private KV pages of every slot back to the pool and drop the slots' references on their voices
static void release_slots(ptts_batch* bt) {
  Ctx* c = bt->ctx;
  for (auto& v : bt->slot_pages) for (int p : v) c->free_pages.push_back(p);
  bt->slot_pages.clear();
  for (int v : bt->slot_voice) voice_unref(*c, v);
  bt->slot_voice.clear();
}

static void batch_free(ptts_batch* bt) {
  Ctx* c = bt->ctx;
  for (auto& g : bt->step_graph) if (g) cudaGraphExecDestroy(g);
  if (bt->step_graph_alt) cudaGraphExecDestroy(bt->step_graph_alt);
  for (auto& g : bt->pipe_graph) if (g) cudaGraphExecDestroy(g);
  for (auto& kv : bt->mimi_graphs) cudaGraphExecDestroy(kv.second.first);
  free_flow_work(bt->fw);
  for (void* p : bt->allocs) cudaFree(p);
  if (bt->d_lat_all) cudaFree(bt->d_lat_all);
  if (bt->d_audio_all) cudaFree(bt->d_audio_all);
  cudaFreeHost(bt->h_noise); cudaFreeHost(bt->h_latent); cudaFreeHost(bt->h_logit); cudaFreeHost(bt->h_audio);
  if (bt->h2_noise) { cudaFreeHost(bt->h2_noise); cudaFreeHost(bt->h2_latent); cudaFreeHost(bt->h2_logit); cudaFreeHost(bt->h2_audio); }
  if (bt->h_chunk[0]) { cudaFreeHost(bt->h_chunk[0]); cudaFreeHost(bt->h_chunk[1]); }
  if (bt->h_pcm) cudaFreeHost(bt->h_pcm);
  if (bt->h2_pcm) cudaFreeHost(bt->h2_pcm);
  for (auto& e : bt->ev_set) if (e) cudaEventDestroy(e);
  release_slots(bt);
  if (bt->mimi_tpl) cudaFree(bt->mimi_tpl);
  delete bt;
}

void ptts_batch_destroy(ptts_batch* bt) {
  if (!bt) return;
  Ctx* c = bt->ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  release_slots(bt);
  bt->prefilled = false;
  if (!bt->h_audio || !bt->d_counter) {   // partially constructed: cannot be recycled
    batch_free(bt);
    return;
  }
  c->parked.push_back(bt);
  while (c->parked.size() > 3) {
    batch_free(c->parked.front());
    c->parked.erase(c->parked.begin());
  }
}

int32_t ptts_batch_prefill_text(ptts_batch* bt, const int32_t* ids, const int32_t* offsets) {
  if (!bt || !ids || !offsets) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  const int B = bt->B, M = offsets[B] - offsets[0];
  if (M < 0) return fail(PTTS_ERR_INVALID, "offsets must be non-decreasing");
  std::vector<int> seq(M), pos(M);
  for (int b = 0; b < B; ++b) {
    const int n = offsets[b + 1] - offsets[b];
    if (n < 0) return fail(PTTS_ERR_INVALID, "offsets must be non-decreasing");
    if (bt->h_len[b] + n > bt->max_len[b]) return fail(PTTS_ERR_STATE, "text of sequence %d exceeds max_len", b);
    for (int i = 0; i < n; ++i) {
      const int id = ids[offsets[b] + i];
      if (id < 0 || id > c.cfg.n_bins) return fail(PTTS_ERR_INVALID, "token id %d out of range", id);
      seq[offsets[b] - offsets[0] + i] = b;
      pos[offsets[b] - offsets[0] + i] = bt->h_len[b] + i;
    }
  }
  if (M > 0) {
    RET(alloc_flow_work(c, c.prefill_work, M));
    if (c.prefill_work.tc) RET(build_flow_plans(c, c.prefill_work, M));
    int *d_seq, *d_pos, *d_ids;
    CU(cudaMalloc(&d_seq, M * 4));
    CU(cudaMalloc(&d_pos, M * 4));
    CU(cudaMalloc(&d_ids, M * 4));
    CU(cudaMemcpyAsync(d_seq, seq.data(), M * 4, cudaMemcpyHostToDevice, c.stream));
    CU(cudaMemcpyAsync(d_pos, pos.data(), M * 4, cudaMemcpyHostToDevice, c.stream));
    CU(cudaMemcpyAsync(d_ids, ids + offsets[0], M * 4, cudaMemcpyHostToDevice, c.stream));
    launch_embed_rows(c.embed, c.bf16, d_ids, c.prefill_work.x, M, c.cfg.d_model, c.stream);
    long long total_keys = 0;
    for (int i = 0; i < M; ++i) total_keys += pos[i] + 1;
    // per-sequence row ranges and start positions for the tensor-core prefill attention
    std::vector<int> rows(2 * B + 1);
    int max_rows = 0;
    for (int b = 0; b <= B; ++b) rows[b] = offsets[b] - offsets[0];
    for (int b = 0; b < B; ++b) {
      rows[B + 1 + b] = bt->h_len[b];
      max_rows = std::max(max_rows, offsets[b + 1] - offsets[b]);
    }
    int* d_rows;
    CU(cudaMalloc(&d_rows, rows.size() * 4));
    CU(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, c.stream));
    c.prefill_work.seq_row0 = d_rows; c.prefill_work.seq_pos0 = d_rows + B + 1;
    c.prefill_work.n_seq = B; c.prefill_work.max_rows_per_seq = max_rows;
    flow_layers(c, c.prefill_work, M, d_seq, d_pos, bt->d_page_table, bt->max_pages, total_keys);
    c.prefill_work.seq_row0 = c.prefill_work.seq_pos0 = nullptr; c.prefill_work.n_seq = 0;
    for (int b = 0; b < B; ++b) bt->h_len[b] += offsets[b + 1] - offsets[b];
    CU(cudaMemcpyAsync(bt->d_len, bt->h_len.data(), B * 4, cudaMemcpyHostToDevice, c.stream));
    CU(cudaStreamSynchronize(c.stream));
    CU(cudaGetLastError());
    cudaFree(d_seq); cudaFree(d_pos); cudaFree(d_ids); cudaFree(d_rows);
  }
  bt->prefilled = true;
  return 0;
}

int32_t ptts_batch_warmup_mimi(ptts_batch* bt, int32_t n_frames) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  for (int i = 0; i < n_frames; ++i) {
    mimi_frame(*bt, bt->d_zero_lat, true);
  }
  // every slot is in the same state now: keep slot 0's as the template that ptts_batch_reset_seq copies into a
  // slot re-used for a new utterance (the warm-up input is the constant emb_mean, so it is sequence-independent)
  if (bt->frame_idx == 0) {
    auto pieces = mimi_state_pieces(*bt);
    size_t total = 0;
    for (auto& p : pieces) total += (p.bytes + 15) & ~(size_t)15;
    if (!bt->mimi_tpl) CU(cudaMalloc(&bt->mimi_tpl, total));
    size_t off = 0;
    for (auto& p : pieces) {
      CU(cudaMemcpyAsync((char*)bt->mimi_tpl + off, p.base, p.bytes, cudaMemcpyDeviceToDevice, c.stream));
      off += (p.bytes + 15) & ~(size_t)15;
    }
    bt->has_tpl = true;
  }
  CU(cudaStreamSynchronize(c.stream));
  CU(cudaGetLastError());
  return 0;
}

static int reset_seq_impl(ptts_batch* bt, int32_t slot, int32_t voice_id, int32_t max_len, std::vector<int>* restore_now = nullptr) {
  Ctx& c = *bt->ctx;
  Batch& t = *bt;
  if (slot < 0 || slot >= t.B) return fail(PTTS_ERR_INVALID, "slot %d out of range", slot);
  if (voice_id < 0 || voice_id >= (int)c.voices.size() || !c.voices[voice_id].alive)
    return fail(PTTS_ERR_INVALID, "unknown voice id %d", voice_id);
  const Voice& v = c.voices[voice_id];
  if (max_len < v.len) return fail(PTTS_ERR_INVALID, "max_len shorter than the voice prefix");
  const int need_pages = (max_len + kPageTokens - 1) / kPageTokens;
  if (need_pages > t.max_pages)
    return fail(PTTS_ERR_INVALID, "max_len %d needs %d KV pages, the batch was created for %d", max_len, need_pages, t.max_pages);
  if (t.fw.prefix_len > 0 && voice_id != t.voice_ids[0]) {
    // The batch attends its shared voice prefix once for all sequences (cascade); a slot that switches voice ends
    // that: fall back to the plain per-sequence attention.  The page tables already cover every sequence's whole
    // prefix, so only the prefix length baked into the captured graphs has to go.
    t.fw.prefix_len = 0;
    t.cascade_len = 0;
    for (auto& g : t.step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    if (t.step_graph_alt) { cudaGraphExecDestroy(t.step_graph_alt); t.step_graph_alt = nullptr; }
    for (auto& g : t.pipe_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
  }
  // KV: give the old private pages back, share the voice's full pages, copy its partial tail page
  const int full = v.len / kPageTokens;
  if ((int)(c.free_pages.size() + t.slot_pages[slot].size()) < need_pages - full)   // checked first: the slot keeps its pages on failure
    return fail(PTTS_ERR_NOMEM, "KV page pool exhausted: slot %d needs %d pages, %zu free", slot, need_pages - full,
                c.free_pages.size() + t.slot_pages[slot].size());
  for (int p : t.slot_pages[slot]) c.free_pages.push_back(p);
  t.slot_pages[slot].clear();
  std::vector<int> mine;
  RET(take_pages(c, need_pages - full, &mine));
  t.slot_pages[slot] = mine;
  int* row = t.h_page_table.data() + (size_t)slot * t.max_pages;
  std::fill(row, row + t.max_pages, 0);
  for (int i = 0; i < full; ++i) row[i] = v.pages[i];
  for (int i = full; i < need_pages; ++i) row[i] = mine[i - full];
  CU(cudaMemcpyAsync(t.d_page_table + (size_t)slot * t.max_pages, row, (size_t)t.max_pages * 4, cudaMemcpyHostToDevice, c.stream));
  if (v.len % kPageTokens) {
    // pageable host -> device copies of 4 bytes complete before cudaMemcpyAsync returns, so the locals may die
    const int sp = v.pages[full], dp = mine[0];
    CU(cudaMemcpyAsync(t.d_cp_src + slot, &sp, 4, cudaMemcpyHostToDevice, c.stream));
    CU(cudaMemcpyAsync(t.d_cp_dst + slot, &dp, 4, cudaMemcpyHostToDevice, c.stream));
    launch_copy_pages(c.pool, c.bf16, c.layer_stride, c.page_stride, c.cfg.n_layers, t.d_cp_src + slot, t.d_cp_dst + slot, 1,
                      c.stream);
  }
  if (t.slot_voice[slot] != voice_id) {
    c.voices[voice_id].refs += 1;          // take the new reference first: v must stay valid below
    voice_unref(c, t.slot_voice[slot]);
    t.slot_voice[slot] = voice_id;
  }
  t.voice_ids[slot] = voice_id;
  t.max_len[slot] = max_len;
  t.h_len[slot] = v.len;
  t.h_active[slot] = 1;
  const int one = 1;
  CU(cudaMemcpyAsync(t.d_len + slot, &t.h_len[slot], 4, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(t.d_bos + slot, &one, 4, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemcpyAsync(t.d_active + slot, &one, 4, cudaMemcpyHostToDevice, c.stream));
  // Mimi: back to the post-warm-up state (or to the zero state when the batch was never warmed up).  In pipelined
  // mode the next frame graph still decodes the previous utterance's last latent for this slot: restore after it.
  if (t.pipelined && t.frame_idx > 0) t.pending_mimi_reset.push_back(slot);
  else if (restore_now) restore_now->push_back(slot);
  else RET(restore_mimi_slot(t, slot));
  return 0;
}

int32_t ptts_batch_reset_seqs(ptts_batch* bt, int32_t n, const int32_t* slots, const int32_t* voice_ids,
                              const int32_t* max_lens) {
  if (!bt || (n > 0 && (!slots || !voice_ids || !max_lens))) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  CU(cudaStreamSynchronize(c.stream));
  std::vector<int> restore_now;          // the slots' Mimi state goes back to the template in one launch
  for (int i = 0; i < n; ++i) RET(reset_seq_impl(bt, slots[i], voice_ids[i], max_lens[i], &restore_now));
  RET(restore_mimi_slots(*bt, restore_now));
  CU(cudaStreamSynchronize(c.stream));
  CU(cudaGetLastError());
  return 0;
}

int32_t ptts_batch_reset_seq(ptts_batch* bt, int32_t slot, int32_t voice_id, int32_t max_len) {
  return ptts_batch_reset_seqs(bt, 1, &slot, &voice_id, &max_len);
}

int32_t ptts_batch_set_active(ptts_batch* bt, int32_t slot, int32_t active) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  if (slot < 0 || slot >= bt->B) return fail(PTTS_ERR_INVALID, "slot %d out of range", slot);
  // stream-ordered and synchronisation-free (4-byte copies from pageable memory are staged before the call returns):
  // frames already enqueued still see the old value, the next one the new value
  const int on = active ? 1 : 0;
  bt->h_active[slot] = on;
  if (!on && bt->h_len[slot] >= bt->max_len[slot]) {
    // a parked slot keeps being stepped with the rest of the batch: it must keep writing inside its own pages
    bt->h_len[slot] = bt->max_len[slot] - 1;
    CU(cudaMemcpyAsync(bt->d_len + slot, &bt->h_len[slot], 4, cudaMemcpyHostToDevice, c.stream));
  }
  CU(cudaMemcpyAsync(bt->d_active + slot, &on, 4, cudaMemcpyHostToDevice, c.stream));
  return 0;
}

int32_t ptts_batch_step(ptts_batch* bt, const float* noise, float* out_latent, float* out_eos_logit,
                        float* out_audio) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  RET(check_step_ready(*bt));
  if (bt->async_staging)
    return fail(PTTS_ERR_STATE, "async staging is on: frames alternate between two staging sets, use ptts_batch_step_staged_async");
  if (bt->pcm16 && out_audio)
    return fail(PTTS_ERR_INVALID, "16-bit PCM output is on: pass out_audio = NULL and read ptts_batch_host_pcm");
  const int B = bt->B, L = c.cfg.latent_dim;
  const bool copy_out = out_latent || out_eos_logit || out_audio || bt->pcm16;
  if (noise) memcpy(bt->h_noise, noise, (size_t)B * L * 4);
  if (bt->pipelined) {
    if (!noise) return fail(PTTS_ERR_INVALID, "pipelined host steps take host noise (use ptts_batch_step_device otherwise)");
    RET(run_pipelined_step(*bt, true));
  } else {
    RET(run_step(*bt, noise != nullptr, copy_out));
    bt->frame_idx += 1;
  }
  CU(cudaStreamSynchronize(c.stream));
  if (out_latent) memcpy(out_latent, bt->h_latent, (size_t)B * L * 4);
  if (out_eos_logit) memcpy(out_eos_logit, bt->h_logit, (size_t)B * 4);
  if (out_audio) memcpy(out_audio, bt->h_audio, (size_t)B * bt->frame_samples * 4);
  return 0;
}

int32_t ptts_batch_host_buffers(ptts_batch* bt, float** noise, float** latent, float** eos_logit, float** audio) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  if (noise) *noise = bt->h_noise;
  if (latent) *latent = bt->h_latent;
  if (eos_logit) *eos_logit = bt->h_logit;
  if (audio) *audio = bt->h_audio;
  return 0;
}

int32_t ptts_batch_step_staged(ptts_batch* bt) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  RET(check_step_ready(*bt));
  if (bt->async_staging)
    return fail(PTTS_ERR_STATE, "async staging is on: frames alternate between two staging sets, use ptts_batch_step_staged_async");
  if (bt->pipelined) {
    RET(run_pipelined_step(*bt, true));
  } else {
    RET(run_step(*bt, true, true));
    bt->frame_idx += 1;
  }
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

int32_t ptts_batch_set_async_staging(ptts_batch* bt, int32_t on) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  if (bt->frame_idx != 0) return fail(PTTS_ERR_STATE, "async staging can only be switched before the first frame");
  if (on && !bt->h2_noise) {
    const int B = bt->B, L = c.cfg.latent_dim;
    CU(cudaMallocHost((void**)&bt->h2_noise, (size_t)B * L * 4));
    CU(cudaMallocHost((void**)&bt->h2_latent, (size_t)B * L * 4));
    CU(cudaMallocHost((void**)&bt->h2_logit, (size_t)B * 4));
    CU(cudaMallocHost((void**)&bt->h2_audio, (size_t)B * bt->frame_samples * 4));
    for (auto& e : bt->ev_set) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (on && bt->pcm16 && !bt->h2_pcm) CU(cudaMallocHost((void**)&bt->h2_pcm, (size_t)bt->B * bt->frame_samples * sizeof(short)));
  bt->async_staging = on != 0;
  if (bt->graphs_async != bt->async_staging) {        // the host pointers are baked into the captured copy nodes
    for (auto& g : bt->pipe_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    bt->graphs_async = bt->async_staging;
  }
  return 0;
}

int32_t ptts_batch_host_buffers_set(ptts_batch* bt, int32_t set, float** noise, float** latent, float** eos_logit,
                                    float** audio) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  if (set == 0) return ptts_batch_host_buffers(bt, noise, latent, eos_logit, audio);
  if (set != 1 || !bt->h2_noise) return fail(PTTS_ERR_STATE, "buffer set %d is not allocated (ptts_batch_set_async_staging)", set);
  if (noise) *noise = bt->h2_noise;
  if (latent) *latent = bt->h2_latent;
  if (eos_logit) *eos_logit = bt->h2_logit;
  if (audio) *audio = bt->h2_audio;
  return 0;
}

int32_t ptts_batch_step_staged_async(ptts_batch* bt, int32_t* set_out) {
  if (!bt || !set_out) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  if (!bt->async_staging) return fail(PTTS_ERR_STATE, "enable async staging first (ptts_batch_set_async_staging)");
  RET(check_step_ready(*bt));
  const int set = (int)(bt->frame_idx & 1);
  if (bt->pipelined) {
    RET(run_pipelined_step(*bt, true));
  } else {
    if (set) RET(run_step_alt(*bt));
    else RET(run_step(*bt, true, true));
    bt->frame_idx += 1;
  }
  CU(cudaEventRecord(bt->ev_set[set], c.stream));
  *set_out = set;
  return 0;
}

int32_t ptts_batch_staged_wait(ptts_batch* bt, int32_t set) {
  if (!bt || set < 0 || set > 1 || !bt->ev_set[set]) return fail(PTTS_ERR_INVALID, "bad buffer set");
  CU(cudaSetDevice(bt->ctx->device));
  CU(cudaEventSynchronize(bt->ev_set[set]));
  return 0;
}

int32_t ptts_batch_step_device(ptts_batch* bt) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  CU(cudaSetDevice(bt->ctx->device));
  RET(check_step_ready(*bt));
  if (bt->pipelined) return run_pipelined_step(*bt, false);
  bt->frame_idx += 1;
  return run_step(*bt, false, false);
}

int32_t ptts_batch_set_pipelined(ptts_batch* bt, int32_t on) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  if (bt->frame_idx != 0) return fail(PTTS_ERR_STATE, "pipelining can only be switched before the first frame");
  bt->pipelined = on != 0;
  return 0;
}

int32_t ptts_batch_set_pcm16(ptts_batch* bt, int32_t on) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  if (bt->frame_idx != 0) return fail(PTTS_ERR_STATE, "the PCM output can only be switched before the first frame");
  const size_t n = (size_t)bt->B * bt->frame_samples;
  if (on && !bt->d_pcm) {
    RET(bt->dalloc((void**)&bt->d_pcm, n * sizeof(short)));
    CU(cudaMallocHost((void**)&bt->h_pcm, n * sizeof(short)));
  }
  if (on && bt->h2_noise && !bt->h2_pcm) CU(cudaMallocHost((void**)&bt->h2_pcm, n * sizeof(short)));
  bt->pcm16 = on != 0;
  if (bt->graphs_pcm != bt->pcm16) {
    drop_graphs(*bt);
    bt->graphs_pcm = bt->pcm16;
  }
  return 0;
}

int32_t ptts_batch_host_pcm(ptts_batch* bt, int32_t set, int16_t** pcm) {
  if (!bt || !pcm) return fail(PTTS_ERR_INVALID, "null argument");
  short* p = set == 0 ? bt->h_pcm : (set == 1 ? bt->h2_pcm : nullptr);
  if (!p) return fail(PTTS_ERR_STATE, "PCM buffer set %d is not allocated (ptts_batch_set_pcm16 / ptts_batch_set_async_staging)", set);
  *pcm = reinterpret_cast<int16_t*>(p);
  return 0;
}

int32_t ptts_batch_flush(ptts_batch* bt, float* out_audio) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  if (!bt->pipelined || bt->frame_idx == 0) return fail(PTTS_ERR_STATE, "nothing to flush");
  // decode the latent of the last stepped frame (it sits in the buffer of parity (frame_idx-1)&1)
  const float* lat = ((bt->frame_idx - 1) & 1) ? bt->d_latent_b : bt->d_latent;
  if (bt->pcm16 && out_audio)
    return fail(PTTS_ERR_INVALID, "16-bit PCM output is on: pass out_audio = NULL and read ptts_batch_host_pcm (set 0)");
  mimi_frame(*bt, lat, true);
  if (out_audio)
    CU(cudaMemcpyAsync(bt->h_audio, bt->d_audio, (size_t)bt->B * bt->frame_samples * 4, cudaMemcpyDeviceToHost, c.stream));
  if (bt->pcm16)
    CU(cudaMemcpyAsync(bt->h_pcm, bt->d_pcm, (size_t)bt->B * bt->frame_samples * sizeof(short), cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  CU(cudaGetLastError());
  if (out_audio) memcpy(out_audio, bt->h_audio, (size_t)bt->B * bt->frame_samples * 4);
  return 0;
}

int32_t ptts_batch_set_prev_latent(ptts_batch* bt, const float* latent) {
  if (!bt || !latent) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  // the next FlowLM step reads the latent of frame frame_idx-1: in pipelined mode that is the ping-pong buffer
  float* dst = (bt->pipelined && ((bt->frame_idx - 1) & 1)) ? bt->d_latent_b : bt->d_latent;
  CU(cudaMemcpyAsync(dst, latent, (size_t)bt->B * c.cfg.latent_dim * 4, cudaMemcpyHostToDevice, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

int32_t ptts_batch_seed(ptts_batch* bt, uint64_t seed) {
  if (!bt) return fail(PTTS_ERR_INVALID, "null batch");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  bt->seed = seed;   // read from device memory by the noise kernel, so the captured graphs stay valid
  const unsigned long long v = seed;
  CU(cudaMemcpyAsync(bt->d_counter + 1, &v, 8, cudaMemcpyHostToDevice, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

int32_t ptts_batch_lengths(ptts_batch* bt, int32_t* out_len) {
  if (!bt || !out_len) return fail(PTTS_ERR_INVALID, "null argument");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  CU(cudaMemcpyAsync(out_len, bt->d_len, bt->B * 4, cudaMemcpyDeviceToHost, c.stream));
  CU(cudaStreamSynchronize(c.stream));
  return 0;
}

int32_t ptts_batch_mimi_decode(ptts_batch* bt, const float* latents, int32_t F, float* audio) {
  if (!bt || !latents || F <= 0) return fail(PTTS_ERR_INVALID, "bad arguments");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  const int B = bt->B, L = c.cfg.latent_dim, n = bt->frame_samples;
  if ((long long)F > bt->lat_all_cap) {
    if (bt->d_lat_all) cudaFree(bt->d_lat_all);
    if (bt->d_audio_all) cudaFree(bt->d_audio_all);
    CU(cudaMalloc((void**)&bt->d_lat_all, (size_t)B * F * L * 4));
    CU(cudaMalloc((void**)&bt->d_audio_all, (size_t)B * F * n * 4));
    bt->lat_all_cap = F;
    for (auto& kv : bt->mimi_graphs) cudaGraphExecDestroy(kv.second.first);
    bt->mimi_graphs.clear();
  }
  CU(cudaMemcpyAsync(bt->d_lat_all, latents, (size_t)B * F * L * 4, cudaMemcpyHostToDevice, c.stream));
  CU(cudaMemsetAsync(bt->d_frame_idx, 0, 4, c.stream));
  auto it = bt->mimi_graphs.find(F);
  if (it == bt->mimi_graphs.end()) {
    const long long before = g_launches;
    cudaGraph_t graph;
    CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
    launch_gather_frame(bt->d_lat_all, bt->d_x, B, F, L, bt->d_frame_idx, c.stream);
    mimi_frame(*bt, bt->d_x, true);
    launch_scatter_audio(bt->d_audio, bt->d_audio_all, B, F, n, bt->d_frame_idx, c.stream);
    launch_inc(bt->d_frame_idx, 1, c.stream);
    CU(cudaStreamEndCapture(c.stream, &graph));
    const long long cnt = g_launches - before;
    g_launches = before;
    cudaGraphExec_t exec;
    CU(cudaGraphInstantiate(&exec, graph, 0));
    CU(cudaGraphDestroy(graph));
    it = bt->mimi_graphs.emplace(F, std::make_pair(exec, cnt)).first;
  }
  if (!audio) {
    for (int f = 0; f < F; ++f) {
      CU(cudaGraphLaunch(it->second.first, c.stream));
      g_launches += it->second.second;
    }
    CU(cudaStreamSynchronize(c.stream));
    CU(cudaGetLastError());
    return 0;
  }
  // Waveforms to the host while the decoder keeps running: frames go out in chunks of kChunk through two pinned staging
  // buffers (strided device -> host copy on the second stream), and this thread moves a finished chunk into the
  // caller's (pageable) array while the GPU decodes the next one.
  constexpr int kChunk = 16;
  const size_t row = (size_t)n * 4;
  if (!bt->h_chunk[0]) {
    CU(cudaHostAlloc((void**)&bt->h_chunk[0], (size_t)B * kChunk * row, cudaHostAllocDefault));
    CU(cudaHostAlloc((void**)&bt->h_chunk[1], (size_t)B * kChunk * row, cudaHostAllocDefault));
  }
  auto drain = [&](int k) -> int {                       // chunk k: staging buffer -> audio[b][f0 .. f0 + nf)
    const int f0 = k * kChunk, nf = std::min(kChunk, F - f0);
    CU(cudaEventSynchronize(c.ev_slice[k & 1]));
    const char* src = reinterpret_cast<const char*>(bt->h_chunk[k & 1]);
    for (int b = 0; b < B; ++b)
      memcpy(reinterpret_cast<char*>(audio) + ((size_t)b * F + f0) * row, src + (size_t)b * nf * row, (size_t)nf * row);
    return 0;
  };
  const int n_chunks = (F + kChunk - 1) / kChunk;
  for (int k = 0; k < n_chunks; ++k) {
    const int f0 = k * kChunk, nf = std::min(kChunk, F - f0);
    for (int f = 0; f < nf; ++f) {
      CU(cudaGraphLaunch(it->second.first, c.stream));
      g_launches += it->second.second;
    }
    CU(cudaEventRecord(c.ev_attn[k & 1], c.stream));
    CU(cudaStreamWaitEvent(c.stream2, c.ev_attn[k & 1], 0));
    CU(cudaMemcpy2DAsync(bt->h_chunk[k & 1], (size_t)nf * row, bt->d_audio_all + (size_t)f0 * n, (size_t)F * row, (size_t)nf * row, B,
                         cudaMemcpyDeviceToHost, c.stream2));
    CU(cudaEventRecord(c.ev_slice[k & 1], c.stream2));
    if (k >= 1) RET(drain(k - 1));
  }
  RET(drain(n_chunks - 1));
  CU(cudaStreamSynchronize(c.stream));
  CU(cudaGetLastError());
  return 0;
}

int32_t ptts_sync(ptts_ctx* c) {
  if (!c) return fail(PTTS_ERR_INVALID, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return 0;
}

int32_t ptts_timer_begin(ptts_ctx* c) {
  if (!c) return fail(PTTS_ERR_INVALID, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev0, c->stream));
  return 0;
}

int32_t ptts_timer_end(ptts_ctx* c, float* ms) {
  if (!c || !ms) return fail(PTTS_ERR_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return 0;
}

int64_t ptts_launch_count(ptts_ctx*, int32_t reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int32_t ptts_batch_profile_step(ptts_batch* bt, const char** report) {
  if (!bt || !report) return fail(PTTS_ERR_INVALID, "bad arguments");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  RET(check_step_ready(*bt));
  CU(cudaStreamSynchronize(c.stream));
  prof_start();
  full_step(*bt, false, false);
  CU(cudaStreamSynchronize(c.stream));
  *report = prof_report();
  CU(cudaGetLastError());
  for (auto& l : bt->h_len) l += 1;
  return 0;
}

int32_t ptts_batch_profile_sections(ptts_batch* bt, float* ms, int32_t cap) {
  if (!bt || !ms || cap < 7) return fail(PTTS_ERR_INVALID, "bad arguments");
  Ctx& c = *bt->ctx;
  CU(cudaSetDevice(c.device));
  RET(check_step_ready(*bt));
  if (!bt->tc_mimi) return fail(PTTS_ERR_STATE, "section profile needs the tensor-core pipeline (batch >= 2, bf16)");
  const long long before = g_launches;
  const int mimi_cap = [] { const char* v = getenv("PTTS_MIMI_GRID"); return v ? atoi(v) : 74; }();
  for (int sec = 0; sec < 7; ++sec) {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CU(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
    switch (sec) {
      case 0: flow_step(*bt, false, 1); break;                 // FlowLM backbone (input, 6 layers)
      case 1: flow_step(*bt, false, 2); break;                 // out-norm + EOS + flow head
      case 2: mimi_frame_tc(*bt, bt->d_latent, 1); break;      // quantizer/upsample + Mimi transformer
      case 3: mimi_frame_tc(*bt, bt->d_latent, 2); break;      // SEANet decoder
      case 4: full_step(*bt, false, false); break;             // whole frame (advances the batch)
      // the Mimi sections as they run in the pipelined frame graph: persistent kernels capped at PTTS_MIMI_GRID SMs
      case 5: gemm_tc_set_grid_cap(mimi_cap); mimi_frame_tc(*bt, bt->d_latent, 1); gemm_tc_set_grid_cap(0); break;
      default: gemm_tc_set_grid_cap(mimi_cap); mimi_frame_tc(*bt, bt->d_latent, 2); gemm_tc_set_grid_cap(0); break;
    }
    CU(cudaStreamEndCapture(c.stream, &graph));
    CU(cudaGraphInstantiate(&exec, graph, 0));
    CU(cudaGraphDestroy(graph));
    const int reps = (sec == 4) ? 4 : 10;
    for (int i = 0; i < 2; ++i) CU(cudaGraphLaunch(exec, c.stream));
    CU(cudaEventRecord(c.ev0, c.stream));
    for (int i = 0; i < reps; ++i) CU(cudaGraphLaunch(exec, c.stream));
    CU(cudaEventRecord(c.ev1, c.stream));
    CU(cudaEventSynchronize(c.ev1));
    float t = 0;
    CU(cudaEventElapsedTime(&t, c.ev0, c.ev1));
    ms[sec] = t / reps;
    if (sec == 4) for (auto& l : bt->h_len) l += reps + 2;
    cudaGraphExecDestroy(exec);
  }
  g_launches = before;
  CU(cudaGetLastError());
  return 7;
}

int32_t ptts_flush_l2(ptts_ctx* c) {
  if (!c) return fail(PTTS_ERR_INVALID, "null ctx");
  CU(cudaSetDevice(c->device));
  if (!c->l2_scratch) {
    c->l2_bytes = 256ull << 20;
    CU(cudaMalloc(&c->l2_scratch, c->l2_bytes));
  }
  launch_fill_u32((unsigned*)c->l2_scratch, 0u, (long long)(c->l2_bytes / 4), c->stream);
  return 0;
}

int32_t ptts_debug_gemm_bench(ptts_ctx* c, int32_t nb, int32_t T, int32_t taps, int32_t C, int32_t N, int32_t epi,
                              const int32_t* force, int32_t reps, float* us, int32_t* chosen) {
  if (!c || !us || !chosen) return fail(PTTS_ERR_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  const int r = gemm_tc_bench(nb, T, taps, C, N, epi, force, reps, us, chosen, c->stream);
  if (r == -1) return fail(PTTS_ERR_INVALID, "configuration not supported by the tcgen05 GEMM");
  if (r < 0) return fail(PTTS_ERR_CUDA, "gemm bench failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Stand-alone run of the cluster chain kernel on a miniature flow head (kernel-level parity test):
//   sy  = silu(a0 W0^T + b0)                       CH_STORE16      [K0 -> D]
//   ada = sy Wa^T + ba  = shift | scale | gate     CH_STORE32      [D -> 3D]
//   x1  = a1 Wi^T + bi ; h = LN(x1; g, b)(1 + scale) + shift       CH_RES_LN (x_init)   [64 -> D]
//   u   = silu(h W1^T + b1)                        CH_STORE16      [D -> D]
//   x1 += gate * (u W2^T + b2) ; h = LN(x1)        CH_RES_LN       [D -> D]
//   lat = lat_in + s (h Wf^T + bf)                 CH_FIN          [D -> 32]
// All matrices are given in fp32 and rounded to bf16 on upload (operands) ; biases / LN / lat_in stay fp32.
int32_t ptts_debug_chain(ptts_ctx* c, int32_t M, int32_t D, int32_t K0, const float* a0, const float* a1, const float* w0,
                         const float* b0, const float* wa, const float* ba, const float* wi, const float* bi,
                         const float* lnw, const float* lnb, const float* w1, const float* b1, const float* w2,
                         const float* b2, const float* wf, const float* bf, const float* lat_in, float out_scale,
                         float* out_x1, float* out_h, float* out_lat, float* out_ada) {
  if (!c) return fail(PTTS_ERR_INVALID, "null ctx");
  CU(cudaSetDevice(c->device));
  const int nc = chain_cluster_size();
  if (nc == 0) return fail(PTTS_ERR_STATE, "the cluster chain kernel is not available on this device");
  std::vector<void*> tmp;
  auto up16 = [&](const float* src, size_t n, __nv_bfloat16** dst) -> int {
    std::vector<uint16_t> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = f2bf(src[i]);
    CU(cudaMalloc((void**)dst, n * 2));
    tmp.push_back(*dst);
    CU(cudaMemcpy(*dst, h.data(), n * 2, cudaMemcpyHostToDevice));
    return 0;
  };
  auto up32 = [&](const float* src, size_t n, float** dst) -> int {
    CU(cudaMalloc((void**)dst, n * 4));
    tmp.push_back(*dst);
    if (src) CU(cudaMemcpy(*dst, src, n * 4, cudaMemcpyHostToDevice));
    else CU(cudaMemset(*dst, 0, n * 4));
    return 0;
  };
  auto done = [&](int rc) { for (void* p : tmp) cudaFree(p); return rc; };
#define DC(x) do { int rc_ = (x); if (rc_ != 0) return done(rc_); } while (0)
  __nv_bfloat16 *d_a0, *d_a1, *d_w0, *d_wa, *d_wi, *d_w1, *d_w2, *d_wf, *d_sy, *d_h, *d_u;
  float *d_b0, *d_ba, *d_bi, *d_lnw, *d_lnb, *d_b1, *d_b2, *d_bf, *d_lat_in, *d_x1, *d_lat, *d_ada;
  DC(up16(a0, (size_t)M * K0, &d_a0)); DC(up16(a1, (size_t)M * 64, &d_a1));
  DC(up16(w0, (size_t)D * K0, &d_w0)); DC(up16(wa, (size_t)3 * D * D, &d_wa)); DC(up16(wi, (size_t)D * 64, &d_wi));
  DC(up16(w1, (size_t)D * D, &d_w1)); DC(up16(w2, (size_t)D * D, &d_w2)); DC(up16(wf, (size_t)32 * D, &d_wf));
  DC(up32(b0, D, &d_b0)); DC(up32(ba, 3 * D, &d_ba)); DC(up32(bi, D, &d_bi)); DC(up32(lnw, D, &d_lnw)); DC(up32(lnb, D, &d_lnb));
  DC(up32(b1, D, &d_b1)); DC(up32(b2, D, &d_b2)); DC(up32(bf, 32, &d_bf)); DC(up32(lat_in, (size_t)M * 32, &d_lat_in));
  DC(up32(nullptr, (size_t)M * D, &d_x1)); DC(up32(nullptr, (size_t)M * 32, &d_lat)); DC(up32(nullptr, (size_t)M * 3 * D, &d_ada));
  {
    std::vector<float> z((size_t)M * D, 0.f);
    DC(up16(z.data(), z.size(), &d_sy)); DC(up16(z.data(), z.size(), &d_h)); DC(up16(z.data(), z.size(), &d_u));
  }
  std::vector<ChainOp> ops(6);
  memset(ops.data(), 0, ops.size() * sizeof(ChainOp));
  bool ok = true;
  auto gemm = [&](ChainOp& o, const __nv_bfloat16* a, int K, const __nv_bfloat16* w, int N, const float* bias, int kind) {
    o.K = K; o.N = N; o.bn = chain_pick_bn(N, nc); o.kind = kind; o.bias = bias;
    ok = ok && o.bn > 0 && chain_encode_a(&o.tm_a, a, M, K, K) && chain_encode_w(&o.tm_w, w, N, K, o.bn);
  };
  gemm(ops[0], d_a0, K0, d_w0, D, d_b0, CH_STORE16); ops[0].act = ACT_SILU; ops[0].y16 = d_sy; ops[0].y_rs = D;
  gemm(ops[1], d_sy, D, d_wa, 3 * D, d_ba, CH_STORE32); ops[1].y32 = d_ada; ops[1].y_rs = 3 * D;
  gemm(ops[2], d_a1, 64, d_wi, D, d_bi, CH_RES_LN);
  ops[2].x = d_x1; ops[2].x_rs = D; ops[2].x_init = 1; ops[2].ln_on = 1; ops[2].ln_w = d_lnw; ops[2].ln_b = d_lnb; ops[2].ln_eps = 1e-6f;
  ops[2].mod_shift = d_ada; ops[2].mod_scale = d_ada + D; ops[2].mod_rs = 3 * D; ops[2].h16 = d_h; ops[2].h_rs = D;
  gemm(ops[3], d_h, D, d_w1, D, d_b1, CH_STORE16); ops[3].act = ACT_SILU; ops[3].y16 = d_u; ops[3].y_rs = D;
  gemm(ops[4], d_u, D, d_w2, D, d_b2, CH_RES_LN);
  ops[4].x = d_x1; ops[4].x_rs = D; ops[4].gate = d_ada + 2 * D; ops[4].gate_rs = 3 * D; ops[4].ln_on = 1; ops[4].ln_eps = 1e-6f;
  ops[4].h16 = d_h; ops[4].h_rs = D;
  gemm(ops[5], d_h, D, d_wf, 32, d_bf, CH_FIN);
  ops[5].out_scale = out_scale; ops[5].lat_in = d_lat_in; ops[5].lat_out = d_lat; ops[5].lat_rs = 32;
  if (!ok) return done(fail(PTTS_ERR_INVALID, "chain plan failed (M=%d D=%d K0=%d nc=%d)", M, D, K0, nc));
  ChainOp* d_ops;
  CU(cudaMalloc((void**)&d_ops, ops.size() * sizeof(ChainOp)));
  tmp.push_back(d_ops);
  CU(cudaMemcpy(d_ops, ops.data(), ops.size() * sizeof(ChainOp), cudaMemcpyHostToDevice));
  chain_launch(d_ops, (int)ops.size(), M, nc, "debug", 0, 0, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return done(fail(PTTS_ERR_CUDA, "debug_chain: %s", cudaGetErrorString(e)));
  std::vector<uint16_t> hh((size_t)M * D);
  CU(cudaMemcpy(hh.data(), d_h, hh.size() * 2, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < hh.size(); ++i) out_h[i] = bf2f(hh[i]);
  CU(cudaMemcpy(out_x1, d_x1, (size_t)M * D * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out_lat, d_lat, (size_t)M * 32 * 4, cudaMemcpyDeviceToHost));
  if (out_ada) CU(cudaMemcpy(out_ada, d_ada, (size_t)M * 3 * D * 4, cudaMemcpyDeviceToHost));
#undef DC
  return done(nc);
}

// PTTS_ATTN_DBG=1: per FlowLM layer {earliest CTA start, latest CTA end} (globaltimer ns) of the decode attention
// launches since the last call; resets the stamps.  Returns the number of layers written (0 when the switch is off).
int32_t ptts_debug_attention_stamps(ptts_ctx* c, uint64_t* out, int32_t max_layers) {
  if (!c || !out) return fail(PTTS_ERR_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  unsigned long long* buf = flow_attention_dbg_buffer();
  if (!buf) return 0;
  const int n = std::min(max_layers, std::min(32, c->cfg.n_layers));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(out, buf, (size_t)n * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  std::vector<unsigned long long> init(64 * 2);
  for (int i = 0; i < 64; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
  CU(cudaMemcpy(buf, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  return n;
}

int32_t ptts_debug_linear(ptts_ctx* c, int32_t path, int32_t nb, int32_t T, int32_t taps, int32_t C, int32_t N,
                          const float* a, const float* w, const float* bias, float* y) {
  if (!c || !a || !w || !y) return fail(PTTS_ERR_INVALID, "null argument");
  CU(cudaSetDevice(c->device));
  const size_t na = (size_t)nb * (T + taps - 1) * C, nw = (size_t)N * taps * C, ny = (size_t)nb * T * N;
  float *d_a, *d_y, *d_b = nullptr;
  void* d_w;
  CU(cudaMalloc((void**)&d_a, na * 4));
  CU(cudaMalloc((void**)&d_y, ny * 4));
  CU(cudaMemcpy(d_a, a, na * 4, cudaMemcpyHostToDevice));
  if (c->bf16) {
    std::vector<uint16_t> h(nw);
    for (size_t i = 0; i < nw; ++i) h[i] = f2bf(w[i]);
    CU(cudaMalloc(&d_w, nw * 2));
    CU(cudaMemcpy(d_w, h.data(), nw * 2, cudaMemcpyHostToDevice));
  } else {
    CU(cudaMalloc(&d_w, nw * 4));
    CU(cudaMemcpy(d_w, w, nw * 4, cudaMemcpyHostToDevice));
  }
  if (bias) {
    CU(cudaMalloc((void**)&d_b, (size_t)N * 4));
    CU(cudaMemcpy(d_b, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  }
  LinearParams p{};
  p.A = d_a; p.a_bs = (long long)(T + taps - 1) * C; p.a_rs = C; p.nb = nb; p.T = T; p.taps = taps; p.C = C;
  p.W = d_w; p.w_bf16 = c->bf16; p.N = N; p.bias = d_b; p.out_scale = 1.f;
  p.Y = d_y; p.y_bs = (long long)T * N; p.y_rs = N;
  int rc = 0;
  if (path == 1) launch_linear_tile(p, c->stream);
  else if (path == 2) {
    if (!linear_gemv_supported(p)) rc = fail(PTTS_ERR_INVALID, "shape not supported by the GEMV path");
    else launch_linear_gemv(p, c->stream);
  } else if (path == 3) {
    rc = gemm_tc_debug(p, c->bf16, c->stream);
    if (rc < 0) rc = fail(PTTS_ERR_INVALID, "shape not supported by the tcgen05 path");
  } else {
    if (linear_gemv_supported(p)) launch_linear_gemv(p, c->stream);
    else launch_linear_tile(p, c->stream);
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess && rc >= 0) e = cudaMemcpy(y, d_y, ny * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_a); cudaFree(d_y); cudaFree(d_w);
  if (d_b) cudaFree(d_b);
  if (rc < 0) return rc;
  if (e != cudaSuccess) return fail(PTTS_ERR_CUDA, "debug_linear: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
