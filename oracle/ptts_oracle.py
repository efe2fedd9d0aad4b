"""CPU oracle for the pocket-tts streaming-generation hot path.  TEST INFRASTRUCTURE ONLY.

This is an independent NumPy restatement of the arithmetic the reference performs on its hot
path.  It is the checker for the CUDA implementation: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline legs may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors and its tensor runtime (`mlx>=0.20.0`,
`pyproject.toml:27`) cannot be installed here, so the oracle is pinned against the reference's
*own Python* executed over `oracle/mlx_shim` (a NumPy restatement of the MLX primitives it
calls).  `oracle/gen_golden.py` produced `tests/golden/*.npz` that way and
`tests/test_oracle_golden.py` holds the oracle to those vectors.  What remains unpinned is MLX's
own kernel rounding (fp32 summation order), which no CPU restatement can reproduce bit-for-bit.

Layout conventions: weights are consumed in the checkpoint's PyTorch layout (Linear [out,in],
Conv1d [out,in,k], ConvTranspose1d [in,out,k]); activations are time-major [T, C]; one object
holds ONE sequence (the reference is batch-1 only, `models/tts_model.py:232-236`).

Each function cites the reference lines it restates (paths relative to
/root/reference/pocket_tts_mlx/).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
from scipy.special import erf as _erf

FRAME_SAMPLES = 1920


# ----------------------------------------------------------------------------- primitives

def layer_norm(x, w, b, eps):
    """Biased-variance LayerNorm over the last axis (nn.LayerNorm; modules/mlp.py:35-50)."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    y = (x - mu) / np.sqrt(var + x.dtype.type(eps))
    if w is not None:
        y = y * w + b
    return y


def rms_star(z, alpha, eps=1e-5):
    """The reference's variance-based "RMSNorm": z*alpha/sqrt(eps+var_unbiased(z)), z NOT centred
    (modules/mlp.py:16-21)."""
    n = z.shape[-1]
    mu = z.mean(axis=-1, keepdims=True)
    var = ((z - mu) ** 2).sum(axis=-1, keepdims=True) / z.dtype.type(n - 1)
    return z * (alpha / np.sqrt(z.dtype.type(eps) + var))


def gelu_erf(x):
    """Exact GELU (nn.gelu; modules/mimi_transformer.py:56)."""
    return x * (1.0 + _erf(x / math.sqrt(2.0)).astype(x.dtype)) * x.dtype.type(0.5)


def silu(x):
    return x / (1.0 + np.exp(-x))


def elu(x):
    """ELU alpha=1 (modules/seanet.py:27,140,158)."""
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1).astype(x.dtype)


def rope_rotate(x, positions, max_period=10000.0):
    """Interleaved-pair RoPE on x [T,H,D] at integer `positions` [T] (modules/rope.py:9-42)."""
    t, h, d = x.shape
    half = d // 2
    dt = x.dtype
    freqs = np.exp(np.arange(half, dtype=np.float32) * np.float32(-math.log(max_period) * 2.0 / d))
    ang = (np.asarray(positions, dtype=np.float32)[:, None] * freqs[None, :]).astype(np.float32)
    if dt == np.float64:  # fp64 mode: exact angles
        freqs64 = np.exp(np.arange(half, dtype=np.float64) * (-math.log(max_period) * 2.0 / d))
        ang = np.asarray(positions, dtype=np.float64)[:, None] * freqs64[None, :]
    c = np.cos(ang).astype(dt)[:, None, :]
    s = np.sin(ang).astype(dt)[:, None, :]
    xr = x[..., 0::2]
    xi = x[..., 1::2]
    out = np.empty_like(x)
    out[..., 0::2] = xr * c - xi * s
    out[..., 1::2] = xr * s + xi * c
    return out


def softmax_rows(s):
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    return e / e.sum(axis=-1, keepdims=True)


# ----------------------------------------------------------------------------- state

@dataclass
class FlowState:
    """Per-sequence FlowLM KV (modules/attention.py:124-137): per layer K,V [L,H,D] (K post-RoPE)."""
    k: List[np.ndarray]
    v: List[np.ndarray]

    @property
    def length(self) -> int:
        return int(self.k[0].shape[0])

    def clone(self) -> "FlowState":
        return FlowState([a.copy() for a in self.k], [a.copy() for a in self.v])


@dataclass
class MimiState:
    """Per-sequence Mimi decode state."""
    up_partial: np.ndarray                   # [16, 512]  upsample overlap (conv.py:176-180)
    ring_k: List[np.ndarray]                 # per layer [250, H, D]
    ring_v: List[np.ndarray]
    offset: int = 0                          # attention.py:200,206-208
    end_offset: int = 0                      # attention.py:202
    conv_prev: Dict[str, np.ndarray] = field(default_factory=dict)     # [k-1, C_in]
    convtr_partial: Dict[str, np.ndarray] = field(default_factory=dict)  # [k-s, C_out]


# ----------------------------------------------------------------------------- the oracle

class Oracle:
    def __init__(self, weights: Dict[str, np.ndarray], cfg, dtype=np.float32, temp: float = 0.7,
                 lsd_decode_steps: int = 1, noise_clamp: Optional[float] = None,
                 eos_threshold: float = -4.0):
        self.cfg = cfg
        self.dt = np.dtype(dtype)
        self.w = {k: np.asarray(v, dtype=self.dt) for k, v in weights.items()}
        self.temp = temp
        self.lsd_decode_steps = lsd_decode_steps
        self.noise_clamp = noise_clamp
        self.eos_threshold = eos_threshold
        t = cfg.flow_lm.transformer
        self.d = t.d_model
        self.h = t.num_heads
        self.n_layers = t.num_layers
        self.max_period = float(t.max_period)
        self.ldim = cfg.mimi.quantizer.dimension
        m = cfg.mimi.transformer
        self.md = m.d_model
        self.mh = m.num_heads
        self.m_layers = m.num_layers
        self.context = m.context
        self.m_max_period = float(m.max_period)
        sn = cfg.mimi.seanet
        self.ratios = list(sn.ratios)
        hop = int(np.prod(self.ratios))
        self.up = int(round(cfg.mimi.sample_rate / hop / cfg.mimi.frame_rate))   # 16
        self.frame_samples = self.up * hop                                       # 1920
        self.frame_rate = cfg.mimi.frame_rate

    # ------------------------------------------------------------------ FlowLM backbone
    def new_flow_state(self) -> FlowState:
        dh = self.d // self.h
        z = lambda: np.zeros((0, self.h, dh), dtype=self.dt)
        return FlowState([z() for _ in range(self.n_layers)], [z() for _ in range(self.n_layers)])

    def _attn_causal(self, q, k, v, first_q_pos):
        """softmax(q k^T / sqrt(D) + M) v with M = -1e9 for key_pos > query_pos
        (modules/attention.py:29-39,167-178).  q [T,H,D]; k,v [L,H,D]; query t sits at
        absolute position first_q_pos + t, key j at position j."""
        t, h, dh = q.shape
        length = k.shape[0]
        s = np.einsum("thd,lhd->htl", q, k) * self.dt.type(1.0 / math.sqrt(dh))
        qpos = first_q_pos + np.arange(t)[:, None]
        kpos = np.arange(length)[None, :]
        s = s + np.where(kpos <= qpos, 0.0, -1e9).astype(self.dt)[None]
        p = softmax_rows(s)
        return np.einsum("htl,lhd->thd", p, v).reshape(t, h * dh)

    def _flow_layer(self, i, x, st: FlowState):
        """One pre-LN layer: x += Attn(LN1 x); x += W2 gelu(W1 LN2 x)
        (modules/mimi_transformer.py:52-69; modules/attention.py:150-182)."""
        w = self.w
        p = f"flow_lm.transformer.layers.{i}"
        t = x.shape[0]
        dh = self.d // self.h
        hcur = layer_norm(x, w[p + ".norm1.weight"], w[p + ".norm1.bias"], 1e-5)
        qkv = (hcur @ w[p + ".self_attn.in_proj.weight"].T).reshape(t, 3, self.h, dh)
        pos0 = st.k[i].shape[0]
        pos = pos0 + np.arange(t)
        q = rope_rotate(qkv[:, 0], pos, self.max_period)
        k = rope_rotate(qkv[:, 1], pos, self.max_period)
        st.k[i] = np.concatenate([st.k[i], k], axis=0)
        st.v[i] = np.concatenate([st.v[i], qkv[:, 2]], axis=0)
        a = self._attn_causal(q, st.k[i], st.v[i], pos0)
        x = x + a @ w[p + ".self_attn.out_proj.weight"].T
        hcur = layer_norm(x, w[p + ".norm2.weight"], w[p + ".norm2.bias"], 1e-5)
        x = x + gelu_erf(hcur @ w[p + ".linear1.weight"].T) @ w[p + ".linear2.weight"].T
        return x

    def backbone(self, rows, st: FlowState):
        """rows [T,d] appended after the cached prefix -> out_norm'd hidden [T,d]
        (models/flow_lm.py:116-122)."""
        x = np.asarray(rows, dtype=self.dt)
        for i in range(self.n_layers):
            x = self._flow_layer(i, x, st)
        return layer_norm(x, self.w["flow_lm.out_norm.weight"], self.w["flow_lm.out_norm.bias"], 1e-5)

    # ------------------------------------------------------------------ flow head
    def time_embedding(self, tau, j):
        """RMS*(W2 silu(W1 [cos(tau f) | sin(tau f)] + b1) + b2; alpha) (modules/mlp.py:53-74)."""
        w = self.w
        p = f"flow_lm.flow_net.time_embed.{j}.mlp"
        half = 128
        f = np.exp(-math.log(10000.0) * np.arange(half, dtype=np.float32) / half).astype(self.dt)
        arg = self.dt.type(tau) * f
        e = np.concatenate([np.cos(arg), np.sin(arg)]).astype(self.dt)
        z = silu(w[p + ".0.weight"] @ e + w[p + ".0.bias"])
        z = w[p + ".2.weight"] @ z + w[p + ".2.bias"]
        return rms_star(z, w[p + ".3.alpha"], 1e-5)

    def flow_velocity(self, c, s, t, x):
        """v(c, s, t, x) of SimpleMLPAdaLN (modules/mlp.py:158-168, 77-119); c [d], x [ldim]."""
        w = self.w
        p = "flow_lm.flow_net"
        x1 = w[p + ".input_proj.weight"] @ x + w[p + ".input_proj.bias"]
        y = (self.time_embedding(s, 0) + self.time_embedding(t, 1)) / self.dt.type(2)
        y = y + (w[p + ".cond_embed.weight"] @ c + w[p + ".cond_embed.bias"])
        sy = silu(y)
        n = x1.shape[0]
        for i in range(self.cfg.flow_lm.flow.depth):
            q = f"{p}.res_blocks.{i}"
            ada = w[q + ".adaLN_modulation.1.weight"] @ sy + w[q + ".adaLN_modulation.1.bias"]
            shift, scale, gate = ada[:n], ada[n:2 * n], ada[2 * n:]
            hcur = layer_norm(x1, w[q + ".in_ln.weight"], w[q + ".in_ln.bias"], 1e-6) * (1 + scale) + shift
            hcur = silu(w[q + ".mlp.0.weight"] @ hcur + w[q + ".mlp.0.bias"])
            hcur = w[q + ".mlp.2.weight"] @ hcur + w[q + ".mlp.2.bias"]
            x1 = x1 + gate * hcur
        ada = w[p + ".final_layer.adaLN_modulation.1.weight"] @ sy + w[p + ".final_layer.adaLN_modulation.1.bias"]
        shift, scale = ada[:n], ada[n:]
        hcur = layer_norm(x1, None, None, 1e-6) * (1 + scale) + shift
        return w[p + ".final_layer.linear.weight"] @ hcur + w[p + ".final_layer.linear.bias"]

    def scaled_noise(self, z):
        """sqrt(temp)*z, optionally clipped (models/flow_lm.py:103-109); z is a raw N(0,1) draw."""
        n = np.asarray(z, dtype=self.dt) * self.dt.type(self.temp ** 0.5)
        if self.noise_clamp is not None:
            n = np.clip(n, -self.noise_clamp, self.noise_clamp)
        return n

    def sample_latent(self, c, z):
        """Euler/LSD integration x <- x + v(s,t,x)/n from the scaled noise (models/flow_lm.py:18-28)."""
        n = self.lsd_decode_steps
        x = self.scaled_noise(z)
        for i in range(n):
            x = x + self.flow_velocity(c, i / n, (i + 1) / n, x) / self.dt.type(n)
        return x

    def eos_logit(self, c):
        return float((self.w["flow_lm.out_eos.weight"] @ c + self.w["flow_lm.out_eos.bias"])[0])

    # ------------------------------------------------------------------ FlowLM calls (tts_model.py:223-269)
    def _flow_call(self, st, text_ids=None, audio_cond=None, latents=None, z=None):
        """[embed(text) | audio_cond | input_linear(latents, NaN->bos)] through the backbone, then
        EOS + one flow sample from the LAST row (models/flow_lm.py:82-114).  Returns
        (latent, eos_logit, hidden_last)."""
        w = self.w
        rows = []
        if text_ids is not None and len(text_ids):
            rows.append(w["flow_lm.conditioner.embed.weight"][np.asarray(text_ids, dtype=np.int64)])
        if audio_cond is not None and len(audio_cond):
            rows.append(np.asarray(audio_cond, dtype=self.dt))
        if latents is not None and len(latents):
            lat = np.asarray(latents, dtype=self.dt)
            lat = np.where(np.isnan(lat), w["flow_lm.bos_emb"][None, :], lat)
            rows.append(lat @ w["flow_lm.input_linear.weight"].T)
        hidden = self.backbone(np.concatenate(rows, axis=0), st)
        c = hidden[-1]
        latent = self.sample_latent(c, z) if z is not None else None
        return latent, self.eos_logit(c), c

    def prefill_audio(self, st, cond, z=None):
        """Voice-prompt prefill (models/tts_model.py:510-512)."""
        return self._flow_call(st, audio_cond=np.asarray(cond).reshape(-1, self.d), z=z)

    def prefill_text(self, st, ids, z=None):
        """Text prefill (models/tts_model.py:388-391)."""
        return self._flow_call(st, text_ids=ids, z=z)

    def step(self, st, prev_latent, z):
        """One autoregressive frame; prev_latent None => BOS (NaN row) (tts_model.py:393-406)."""
        lat = np.full((1, self.ldim), np.nan, dtype=self.dt) if prev_latent is None \
            else np.asarray(prev_latent, dtype=self.dt).reshape(1, self.ldim)
        return self._flow_call(st, latents=lat, z=z)

    # ------------------------------------------------------------------ Mimi decode
    def new_mimi_state(self) -> MimiState:
        dh = self.md // self.mh
        z = lambda *s: np.zeros(s, dtype=self.dt)
        st = MimiState(
            up_partial=z(self.up, self.cfg.mimi.seanet.dimension),
            ring_k=[z(self.context, self.mh, dh) for _ in range(self.m_layers)],
            ring_v=[z(self.context, self.mh, dh) for _ in range(self.m_layers)],
        )
        return st

    def _conv(self, name, x, st: MimiState):
        """Streaming causal Conv1d, stride 1, dilation 1: y[t] = b + sum_j W[:,:,j] x~[t+j],
        x~ = [previous | x], previous = last k-1 rows (modules/conv.py:121-150)."""
        wt = self.w[name + ".weight"]          # [out, in, k]
        b = self.w[name + ".bias"]
        k = wt.shape[2]
        t = x.shape[0]
        if k > 1:
            prev = st.conv_prev.get(name)
            if prev is None:
                prev = np.zeros((k - 1, wt.shape[1]), dtype=self.dt)
            xx = np.concatenate([prev, x], axis=0)
            st.conv_prev[name] = xx[-(k - 1):].copy()
        else:
            xx = x
        y = np.broadcast_to(b, (t, wt.shape[0])).astype(self.dt).copy()
        for j in range(k):
            y += xx[j:j + t] @ wt[:, :, j].T
        return y

    def _convtr(self, name, x, st: MimiState, stride, bias=True, depthwise=False):
        """Streaming ConvTranspose1d with overlap-add (modules/conv.py:182-200):
        full[t*s+j] += x[t] W[:,:,j] (+b everywhere); head += partial; emit T*s rows;
        partial <- tail - b."""
        wt = self.w[name + ".weight"]          # [in, out, k]  (depthwise: [C,1,k])
        k = wt.shape[2]
        t = x.shape[0]
        c_out = x.shape[1] if depthwise else wt.shape[1]
        full = np.zeros(((t - 1) * stride + k, c_out), dtype=self.dt)
        for j in range(k):
            contrib = x * wt[:, 0, j][None, :] if depthwise else x @ wt[:, :, j]
            full[j: j + (t - 1) * stride + 1: stride] += contrib
        b = self.w[name + ".bias"] if bias else None
        if b is not None:
            full = full + b
        pt = k - stride
        key = name
        part = st.convtr_partial.get(key)
        if part is None:
            part = np.zeros((pt, c_out), dtype=self.dt)
        full[:pt] += part
        tail = full[-pt:]
        st.convtr_partial[key] = (tail - b if b is not None else tail).copy()
        return full[: t * stride]

    def _mimi_attention(self, i, x, st: MimiState):
        """Ring-buffer windowed attention incl. the write-before-attend visibility quirk
        (modules/attention.py:67-105, 220-264)."""
        w = self.w
        p = f"mimi.decoder_transformer.transformer.layers.{i}.self_attn"
        t = x.shape[0]
        dh = self.md // self.mh
        cap = self.context
        qkv = (x @ w[p + ".in_proj.weight"].T).reshape(t, 3, self.mh, dh)
        pos_q = st.offset + np.arange(t)
        q = rope_rotate(qkv[:, 0], pos_q, self.m_max_period)
        k = rope_rotate(qkv[:, 1], pos_q, self.m_max_period)
        e = st.end_offset
        for tt in range(t):
            slot = (e + tt) % cap
            st.ring_k[i][slot] = k[tt]
            st.ring_v[i][slot] = qkv[tt, 2]
        last = e + t - 1
        end_index = last % cap
        slots = np.arange(cap)
        delta = slots - end_index
        pos_k = np.where(delta <= 0, last + delta, last + delta - cap)
        pos_k = np.where(slots >= e + t, -1, pos_k)
        dq = pos_q[:, None] - pos_k[None, :]
        visible = (pos_k[None, :] >= 0) & (dq >= 0) & (dq < cap)
        s = np.einsum("thd,lhd->htl", q, st.ring_k[i]) * self.dt.type(1.0 / math.sqrt(dh))
        s = s + np.where(visible, 0.0, -1e9).astype(self.dt)[None]
        pr = softmax_rows(s)
        a = np.einsum("htl,lhd->thd", pr, st.ring_v[i]).reshape(t, self.md)
        return a @ w[p + ".out_proj.weight"].T

    def mimi_decode_frame(self, st: MimiState, latent):
        """latent [ldim] (un-normalised) -> 1920 samples (tts_model.py:415-419; models/mimi.py:70-75)."""
        w = self.w
        lat = np.asarray(latent, dtype=self.dt).reshape(self.ldim)
        z = w["mimi.quantizer.output_proj.weight"][:, :, 0] @ (lat * w["flow_lm.emb_std"] + w["flow_lm.emb_mean"])
        # depthwise upsample k=2*up, stride up, no bias (modules/resample.py:27-42)
        wu = w["mimi.upsample.convtr.convtr.weight"][:, 0, :]      # [C, 2*up]
        full = (z[:, None] * wu).T                                  # [2*up, C]
        full[: self.up] += st.up_partial
        st.up_partial = full[self.up:].copy()
        x = full[: self.up]                                          # [16, 512]
        # decoder transformer with LayerScale (modules/mimi_transformer.py:52-69,160-171)
        t = x.shape[0]
        for i in range(self.m_layers):
            p = f"mimi.decoder_transformer.transformer.layers.{i}"
            hcur = layer_norm(x, w[p + ".norm1.weight"], w[p + ".norm1.bias"], 1e-5)
            x = x + w[p + ".layer_scale_1.scale"] * self._mimi_attention(i, hcur, st)
            hcur = layer_norm(x, w[p + ".norm2.weight"], w[p + ".norm2.bias"], 1e-5)
            x = x + w[p + ".layer_scale_2.scale"] * (gelu_erf(hcur @ w[p + ".linear1.weight"].T) @ w[p + ".linear2.weight"].T)
        st.offset += t
        st.end_offset += t
        # SEANet decoder (modules/seanet.py:136-170)
        idx = 0
        x = self._conv(f"mimi.decoder.model.{idx}.conv", x, st)
        idx += 1
        for r in self.ratios:
            idx += 1
            x = self._convtr(f"mimi.decoder.model.{idx}.convtr", elu(x), st, r)
            idx += 1
            hcur = self._conv(f"mimi.decoder.model.{idx}.block.1.conv", elu(x), st)
            hcur = self._conv(f"mimi.decoder.model.{idx}.block.3.conv", elu(hcur), st)
            x = x + hcur
            idx += 1
        idx += 1
        y = self._conv(f"mimi.decoder.model.{idx}.conv", elu(x), st)
        return y[:, 0]

    def warmup_mimi(self, st: MimiState, warmup_frames: int):
        """Decode `warmup_frames` zero latents and discard (tts_model.py:464-476)."""
        for _ in range(max(0, warmup_frames)):
            self.mimi_decode_frame(st, np.zeros(self.ldim, dtype=self.dt))

    # ------------------------------------------------------------------ generation loop
    # ------------------------------------------------------------------ voice cloning (Mimi encode side)
    def _conv_ns(self, name, x, stride=1, replicate=False, bias=True):
        """Non-streaming call of StreamingConv1d (model_state=None => fresh state, modules/conv.py:121-150):
        left context = kernel - stride columns, zeros or (pad_mode="replicate") copies of the first column;
        y[t] = b + sum_j W[:,:,j] x~[t*stride + j].  x [T, C_in] -> [T/stride, C_out]."""
        wt = self.w[name + ".weight"]          # [out, in, k]
        k = wt.shape[2]
        t = x.shape[0]
        assert t % stride == 0
        tp = k - stride
        if tp > 0:
            head = np.repeat(x[:1], tp, axis=0) if replicate else np.zeros((tp, x.shape[1]), dtype=self.dt)
            xx = np.concatenate([head, x], axis=0)
        else:
            xx = x
        t_out = t // stride
        y = np.zeros((t_out, wt.shape[0]), dtype=self.dt)
        if bias:
            y += self.w[name + ".bias"]
        for j in range(k):
            y += xx[j: j + (t_out - 1) * stride + 1: stride] @ wt[:, :, j].T
        return y

    def _enc_attention(self, i, x):
        """Non-streaming windowed causal attention of the encoder transformer (modules/attention.py:210-264 with
        model_state=None): positions 0..T-1, key visible iff 0 <= q - k < context."""
        w = self.w
        p = f"mimi.encoder_transformer.transformer.layers.{i}.self_attn"
        t = x.shape[0]
        dh = self.md // self.mh
        qkv = (x @ w[p + ".in_proj.weight"].T).reshape(t, 3, self.mh, dh)
        pos = np.arange(t)
        q = rope_rotate(qkv[:, 0], pos, self.m_max_period)
        k = rope_rotate(qkv[:, 1], pos, self.m_max_period)
        dq = pos[:, None] - pos[None, :]
        visible = (dq >= 0) & (dq < self.context)
        sc = np.einsum("thd,lhd->htl", q, k) * self.dt.type(1.0 / math.sqrt(dh))
        sc = sc + np.where(visible, 0.0, -1e9).astype(self.dt)[None]
        pr = softmax_rows(sc)
        a = np.einsum("htl,lhd->thd", pr, qkv[:, 2]).reshape(t, self.md)
        return a @ w[p + ".out_proj.weight"].T

    def encode_audio(self, audio):
        """Waveform [T] (24 kHz mono) -> FlowLM conditioning [T_v, d_model]: zero-pad the end to whole frames
        (conv.py:12-26), SEANet encoder (seanet.py:45-108), encoder transformer, stride-16 replicate-padded
        downsample (resample.py:8-24), speaker projection (tts_model.py:271-276)."""
        w = self.w
        sn = self.cfg.mimi.seanet
        x = np.asarray(audio, dtype=self.dt).reshape(-1)
        frame = int(self.cfg.mimi.sample_rate / self.frame_rate)
        n_frames = math.ceil((x.shape[0] - frame) / frame + 1)
        ideal = (n_frames - 1) * frame + frame
        if ideal > x.shape[0]:
            x = np.concatenate([x, np.zeros(ideal - x.shape[0], dtype=self.dt)])
        h = self._conv_ns("mimi.encoder.model.0.conv", x[:, None])
        idx = 1
        for r in reversed(sn.ratios):
            pre = f"mimi.encoder.model.{idx}.block"
            y = self._conv_ns(pre + ".1.conv", elu(h))
            h = h + self._conv_ns(pre + ".3.conv", elu(y))
            idx += 2
            h = self._conv_ns(f"mimi.encoder.model.{idx}.conv", elu(h), stride=r)
            idx += 1
        idx += 1
        h = self._conv_ns(f"mimi.encoder.model.{idx}.conv", elu(h))
        for i in range(self.m_layers):
            p = f"mimi.encoder_transformer.transformer.layers.{i}"
            a = self._enc_attention(i, layer_norm(h, w[p + ".norm1.weight"], w[p + ".norm1.bias"], 1e-5))
            h = h + a * w[p + ".layer_scale_1.scale"]
            f = gelu_erf(layer_norm(h, w[p + ".norm2.weight"], w[p + ".norm2.bias"], 1e-5) @ w[p + ".linear1.weight"].T)
            h = h + (f @ w[p + ".linear2.weight"].T) * w[p + ".layer_scale_2.scale"]
        lat = self._conv_ns("mimi.downsample.conv.conv", h, stride=self.up, replicate=True, bias=False)
        return (lat @ w["flow_lm.speaker_proj_weight"].T).astype(self.dt)

    def max_gen_len(self, n_tok: int) -> int:
        """ceil((n_tok/3 + 2) * frame_rate) (tts_model.py:440-444)."""
        return math.ceil((n_tok / 3.0 + 2.0) * self.frame_rate)

    def generate(self, voice_state: FlowState, ids: Sequence[int], noise, frames_after_eos: int,
                 warmup_frames: int = 1, max_frames: Optional[int] = None, decode_audio: bool = True,
                 teacher_latents=None):
        """The per-chunk loop of tts_model.py:363-428.

        `noise` [1 + n_frames, ldim]: row 0 is consumed by the text prefill (the reference draws
        there too and discards the sample), row 1+g by frame g.  Returns dict with latents
        [F,ldim], eos_logits [G], audio [F*1920], n_frames F, hidden [G,d].
        `teacher_latents` (optional [G,ldim]) feeds those back instead of the oracle's own output."""
        st = voice_state.clone()
        mst = self.new_mimi_state()
        self.warmup_mimi(mst, warmup_frames)
        noise = np.asarray(noise)
        self.prefill_text(st, ids, z=noise[0])
        limit = self.max_gen_len(len(ids))
        if max_frames is not None:
            limit = min(limit, max_frames)
        prev = None
        eos_step = None
        latents, logits, audio, hidden = [], [], [], []
        for g in range(limit):
            lat, logit, c = self.step(st, prev, noise[1 + g])
            logits.append(logit)
            hidden.append(c)
            if logit > self.eos_threshold and eos_step is None:
                eos_step = g
            if eos_step is not None and g >= eos_step + frames_after_eos:
                break
            latents.append(lat)
            if decode_audio:
                audio.append(self.mimi_decode_frame(mst, lat))
            prev = lat if teacher_latents is None else teacher_latents[g]
        return {
            "latents": np.array(latents, dtype=self.dt).reshape(-1, self.ldim),
            "eos_logits": np.array(logits, dtype=np.float64),
            "audio": np.concatenate(audio) if audio else np.zeros(0, dtype=self.dt),
            "n_frames": len(latents),
            "hidden": np.array(hidden, dtype=self.dt),
        }


def postprocess_audio_start(audio, sample_rate, trim_start_ms=0, fade_in_ms=0):
    """Trim then linear fade-in, endpoints inclusive (tts_model.py:446-462)."""
    audio = np.asarray(audio)
    if trim_start_ms > 0:
        n = int(sample_rate * trim_start_ms / 1000)
        if 0 < n < audio.shape[0]:
            audio = audio[n:]
    if fade_in_ms > 0 and audio.shape[0] > 1:
        n = min(max(0, int(sample_rate * fade_in_ms / 1000)), audio.shape[0])
        if n > 1:
            ramp = np.linspace(0.0, 1.0, n).astype(audio.dtype)
            audio = np.concatenate([audio[:n] * ramp, audio[n:]])
    return audio
