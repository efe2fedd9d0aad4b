"""ctypes binding of libptts_b200.so (include/ptts.h).  No torch, no CPU fallback: if the shared
library is missing or no CUDA device is usable the calls raise."""

from __future__ import annotations

import ctypes as C
import math
import weakref
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libptts_b200.so"

PTTS_BF16, PTTS_FP32 = 0, 1
DT_F32, DT_BF16, DT_F16 = 0, 1, 2

EXPORTED = [
    "ptts_abi_version", "ptts_last_error", "ptts_device_count", "ptts_ctx_create", "ptts_ctx_destroy",
    "ptts_load_weight", "ptts_finalize_weights", "ptts_voice_create", "ptts_voice_destroy",
    "ptts_voice_length", "ptts_batch_create", "ptts_batch_destroy", "ptts_batch_prefill_text",
    "ptts_batch_warmup_mimi", "ptts_batch_step", "ptts_batch_set_prev_latent", "ptts_batch_step_device",
    "ptts_batch_seed", "ptts_batch_lengths", "ptts_batch_mimi_decode", "ptts_sync", "ptts_timer_begin",
    "ptts_timer_end", "ptts_launch_count", "ptts_batch_profile_step", "ptts_flush_l2", "ptts_debug_linear",
    "ptts_debug_gemm_bench", "ptts_batch_profile_sections", "ptts_batch_set_pipelined", "ptts_batch_flush",
    "ptts_batch_reset_seq", "ptts_batch_reset_seqs", "ptts_batch_set_active",
    "ptts_has_voice_cloning", "ptts_encode_audio",
    "ptts_batch_set_async_staging", "ptts_batch_host_buffers_set", "ptts_batch_step_staged_async", "ptts_batch_staged_wait",
    "ptts_batch_host_buffers", "ptts_batch_step_staged",
    "ptts_batch_set_pcm16", "ptts_batch_host_pcm", "ptts_unused_weights", "ptts_debug_chain", "ptts_debug_attention_stamps",
]


class PttsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libptts_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("n_heads", C.c_int32), ("n_layers", C.c_int32), ("ffn_dim", C.c_int32),
        ("n_bins", C.c_int32), ("latent_dim", C.c_int32), ("max_period", C.c_float),
        ("flow_dim", C.c_int32), ("flow_depth", C.c_int32),
        ("mimi_d", C.c_int32), ("mimi_heads", C.c_int32), ("mimi_layers", C.c_int32), ("mimi_ffn", C.c_int32),
        ("mimi_context", C.c_int32), ("mimi_max_period", C.c_float),
        ("seanet_dim", C.c_int32), ("n_filters", C.c_int32), ("n_ratios", C.c_int32), ("ratios", C.c_int32 * 8),
        ("kernel_size", C.c_int32), ("res_kernel_size", C.c_int32), ("last_kernel_size", C.c_int32),
        ("compress", C.c_int32), ("upsample_stride", C.c_int32),
        ("temp", C.c_float), ("lsd_decode_steps", C.c_int32), ("noise_clamp", C.c_float),
        ("eos_threshold", C.c_float),
        ("precision", C.c_int32), ("kv_pool_tokens", C.c_int64), ("max_batch", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pocket_tts_mlx_b200.build_native` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, f32p, i32p = C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sig = {
        "ptts_abi_version": (i32, []),
        "ptts_last_error": (C.c_char_p, []),
        "ptts_device_count": (i32, []),
        "ptts_ctx_create": (i32, [i32, C.POINTER(Config), C.POINTER(vp)]),
        "ptts_ctx_destroy": (None, [vp]),
        "ptts_load_weight": (i32, [vp, C.c_char_p, i32, i32, C.POINTER(C.c_int64), vp]),
        "ptts_finalize_weights": (i32, [vp]),
        "ptts_unused_weights": (C.c_char_p, [vp]),
        "ptts_voice_create": (i32, [vp, f32p, i32]),
        "ptts_voice_destroy": (i32, [vp, i32]),
        "ptts_voice_length": (i32, [vp, i32]),
        "ptts_batch_create": (i32, [vp, i32, i32p, i32p, C.POINTER(vp)]),
        "ptts_batch_destroy": (None, [vp]),
        "ptts_batch_prefill_text": (i32, [vp, i32p, i32p]),
        "ptts_batch_warmup_mimi": (i32, [vp, i32]),
        "ptts_batch_step": (i32, [vp, f32p, f32p, f32p, f32p]),
        "ptts_batch_set_prev_latent": (i32, [vp, f32p]),
        "ptts_batch_step_device": (i32, [vp]),
        "ptts_batch_seed": (i32, [vp, C.c_uint64]),
        "ptts_batch_lengths": (i32, [vp, i32p]),
        "ptts_batch_mimi_decode": (i32, [vp, f32p, i32, f32p]),
        "ptts_sync": (i32, [vp]),
        "ptts_timer_begin": (i32, [vp]),
        "ptts_timer_end": (i32, [vp, f32p]),
        "ptts_launch_count": (C.c_int64, [vp, i32]),
        "ptts_batch_profile_step": (i32, [vp, C.POINTER(C.c_char_p)]),
        "ptts_flush_l2": (i32, [vp]),
        "ptts_debug_linear": (i32, [vp, i32, i32, i32, i32, i32, i32, f32p, f32p, f32p, f32p]),
        "ptts_batch_profile_sections": (i32, [vp, f32p, i32]),
        "ptts_batch_host_buffers": (i32, [vp, C.POINTER(f32p), C.POINTER(f32p), C.POINTER(f32p), C.POINTER(f32p)]),
        "ptts_batch_step_staged": (i32, [vp]),
        "ptts_batch_set_pipelined": (i32, [vp, i32]),
        "ptts_batch_set_async_staging": (i32, [vp, i32]),
        "ptts_batch_host_buffers_set": (i32, [vp, i32, C.POINTER(f32p), C.POINTER(f32p), C.POINTER(f32p), C.POINTER(f32p)]),
        "ptts_batch_step_staged_async": (i32, [vp, i32p]),
        "ptts_batch_staged_wait": (i32, [vp, i32]),
        "ptts_has_voice_cloning": (i32, [vp]),
        "ptts_encode_audio": (i32, [vp, f32p, C.c_int64, f32p, i32, i32p]),
        "ptts_batch_reset_seq": (i32, [vp, i32, i32, i32]),
        "ptts_batch_reset_seqs": (i32, [vp, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "ptts_batch_set_active": (i32, [vp, i32, i32]),
        "ptts_batch_flush": (i32, [vp, f32p]),
        "ptts_batch_set_pcm16": (i32, [vp, i32]),
        "ptts_batch_host_pcm": (i32, [vp, i32, C.POINTER(C.POINTER(C.c_int16))]),
        "ptts_debug_gemm_bench": (i32, [vp, i32, i32, i32, i32, i32, i32, i32p, i32, f32p, i32p]),
        "ptts_debug_attention_stamps": (i32, [vp, C.POINTER(C.c_uint64), i32]),
        "ptts_debug_chain": (i32, [vp, i32, i32, i32] + [f32p] * 17 + [C.c_float] + [f32p] * 4),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise PttsError(rc, lib().ptts_last_error().decode("utf-8", "replace"))
    return rc


def device_count() -> int:
    return int(lib().ptts_device_count())


def _big_empty(shape, dtype) -> np.ndarray:
    """Large result arrays are touched for the first time while the GPU is producing them: with 4 KB pages the page
    faults of a 256-utterance job (~0.5 GB) cost about as much host time as the job itself.  Ask for transparent huge
    pages (2 MB) when the platform offers them; otherwise this is np.empty."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= (64 << 20):
        try:
            import mmap
            if hasattr(mmap, "MADV_HUGEPAGE"):
                buf = mmap.mmap(-1, nbytes + (2 << 20))
                buf.madvise(mmap.MADV_HUGEPAGE)
                base = np.frombuffer(buf, dtype=np.uint8)
                off = (-base.ctypes.data) % (2 << 20)
                return base[off:off + nbytes].view(dtype).reshape(shape)
        except (OSError, ValueError, AttributeError):
            pass
    return np.empty(shape, dtype=dtype)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def make_config(cfg, temp: float, lsd_decode_steps: int, noise_clamp: Optional[float], eos_threshold: float,
                precision: str = "bf16", kv_pool_tokens: int = 262144, max_batch: int = 0) -> Config:
    """pydantic Config (config.py) -> C struct."""
    t, m, sn = cfg.flow_lm.transformer, cfg.mimi.transformer, cfg.mimi.seanet
    c = Config()
    c.d_model, c.n_heads, c.n_layers = t.d_model, t.num_heads, t.num_layers
    c.ffn_dim = int(t.d_model * t.hidden_scale)
    c.n_bins, c.latent_dim = cfg.flow_lm.lookup_table.n_bins, cfg.mimi.quantizer.dimension
    c.max_period = float(t.max_period)
    c.flow_dim, c.flow_depth = cfg.flow_lm.flow.dim, cfg.flow_lm.flow.depth
    c.mimi_d, c.mimi_heads, c.mimi_layers = m.d_model, m.num_heads, m.num_layers
    c.mimi_ffn, c.mimi_context, c.mimi_max_period = m.dim_feedforward, m.context, float(m.max_period)
    c.seanet_dim, c.n_filters, c.n_ratios = sn.dimension, sn.n_filters, len(sn.ratios)
    for i, r in enumerate(sn.ratios):
        c.ratios[i] = int(r)
    c.kernel_size, c.res_kernel_size, c.last_kernel_size = sn.kernel_size, sn.residual_kernel_size, sn.last_kernel_size
    c.compress = sn.compress
    hop = int(np.prod(sn.ratios))
    c.upsample_stride = int(round(cfg.mimi.sample_rate / hop / cfg.mimi.frame_rate))
    c.temp = float(temp)
    c.lsd_decode_steps = int(lsd_decode_steps)
    c.noise_clamp = -1.0 if noise_clamp is None else float(noise_clamp)
    c.eos_threshold = float(min(max(eos_threshold, -3.0e38), 3.0e38))
    if precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    c.precision = PTTS_BF16 if precision == "bf16" else PTTS_FP32
    c.kv_pool_tokens = int(kv_pool_tokens)
    c.max_batch = int(max_batch)
    if m.d_model != m.input_dimension or tuple(m.output_dimensions) != (m.d_model,):
        raise ValueError("Mimi transformer input/output projections are not supported (identity at 512/512)")
    if sn.n_residual_layers != 1 or sn.pad_mode != "constant" or sn.channels != 1:
        raise ValueError("only the b6369a24 SEANet decoder topology is supported")
    return c


class Context:
    """One GPU: weights, KV page pool, voices."""

    def __init__(self, config: Config, device: int = 0):
        self._h = C.c_void_p()
        self.config = config
        self._batches = weakref.WeakSet()       # live batches: closed before the context goes away
        check(lib().ptts_ctx_create(device, C.byref(config), C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            for b in list(self._batches):
                b.close()
            lib().ptts_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weight(self, name: str, arr: np.ndarray) -> bool:
        a = _f32(arr)
        shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
        rc = check(lib().ptts_load_weight(self._h, name.encode(), DT_F32, a.ndim, shape, a.ctypes.data_as(C.c_void_p)))
        return rc == 0

    def finalize(self):
        check(lib().ptts_finalize_weights(self._h))

    def unused_weights(self):
        """Checkpoint keys that were offered to load_weight but that nothing consumed."""
        return [k for k in lib().ptts_unused_weights(self._h).decode().splitlines() if k]

    def voice_create(self, cond: np.ndarray) -> int:
        a = _f32(cond).reshape(-1, self.config.d_model)
        return check(lib().ptts_voice_create(self._h, _fp(a), a.shape[0]))

    @property
    def has_voice_cloning(self) -> bool:
        return bool(lib().ptts_has_voice_cloning(self._h))

    def encode_audio(self, audio: np.ndarray, frame_samples: int = 1920) -> np.ndarray:
        """Mono waveform at the model sample rate -> conditioning [ceil(T / frame), d_model] (Mimi encoder +
        speaker projection)."""
        a = _f32(audio).reshape(-1)
        max_frames = (a.shape[0] + frame_samples - 1) // frame_samples + 1
        out = np.empty((max_frames, self.config.d_model), dtype=np.float32)
        n = C.c_int32(0)
        check(lib().ptts_encode_audio(self._h, _fp(a), a.shape[0], _fp(out), max_frames, C.byref(n)))
        return out[: n.value].copy()

    def voice_destroy(self, vid: int):
        check(lib().ptts_voice_destroy(self._h, vid))

    def voice_length(self, vid: int) -> int:
        return check(lib().ptts_voice_length(self._h, vid))

    def sync(self):
        check(lib().ptts_sync(self._h))

    def timer_begin(self):
        check(lib().ptts_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        check(lib().ptts_timer_end(self._h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self, reset: bool = False) -> int:
        return int(lib().ptts_launch_count(self._h, 1 if reset else 0))

    def flush_l2(self):
        check(lib().ptts_flush_l2(self._h))

    def gemm_bench(self, nb, t, taps, c_in, n_out, epi=0, force=None, reps=5):
        """Median microseconds of the tcgen05 GEMM on synthetic operands -> (us, (bn, stages, splits, persist))."""
        us = C.c_float()
        chosen = (C.c_int32 * 4)()
        f = (C.c_int32 * 4)(*force) if force is not None else None
        check(lib().ptts_debug_gemm_bench(self._h, nb, t, taps, c_in, n_out, epi, f, reps, C.byref(us), chosen))
        return float(us.value), tuple(int(x) for x in chosen)

    def attention_stamps(self, max_layers: int = 16) -> np.ndarray:
        """PTTS_ATTN_DBG=1: [layers, 2] {earliest CTA start, latest CTA end} in ns of the last decode-attention launches."""
        out = np.zeros((max_layers, 2), dtype=np.uint64)
        n = check(lib().ptts_debug_attention_stamps(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), max_layers))
        return out[:n]

    def debug_chain(self, t: dict, out_scale: float = 1.0):
        """Miniature flow head through the cluster chain kernel (see ptts_debug_chain); t holds the fp32 inputs
        a0, a1, w0, b0, wa, ba, wi, bi, lnw, lnb, w1, b1, w2, b2, wf, bf, lat_in.  -> dict(x1, h, lat, ada, nc)."""
        a0 = _f32(t["a0"])
        m, k0 = a0.shape
        d = t["w0"].shape[0]
        names = ["a0", "a1", "w0", "b0", "wa", "ba", "wi", "bi", "lnw", "lnb", "w1", "b1", "w2", "b2", "wf", "bf", "lat_in"]
        arrs = [_f32(t[k]) for k in names]
        x1 = np.empty((m, d), np.float32)
        h = np.empty((m, d), np.float32)
        lat = np.empty((m, 32), np.float32)
        ada = np.empty((m, 3 * d), np.float32)
        nc = check(lib().ptts_debug_chain(self._h, m, d, k0, *[_fp(a) for a in arrs], C.c_float(out_scale),
                                          _fp(x1), _fp(h), _fp(lat), _fp(ada)))
        return {"x1": x1, "h": h, "lat": lat, "ada": ada, "nc": nc}

    def debug_linear(self, a, w, bias=None, taps=1, path=0):
        """a [nb, T+taps-1, C], w [N, taps*C] -> y [nb, T, N] through the chosen kernel path."""
        a = _f32(a)
        w = _f32(w)
        nb, tp, c_in = a.shape
        t = tp - taps + 1
        n = w.shape[0]
        assert w.shape[1] == taps * c_in
        y = np.empty((nb, t, n), dtype=np.float32)
        b = _f32(bias) if bias is not None else None
        check(lib().ptts_debug_linear(self._h, path, nb, t, taps, c_in, n, _fp(a), _fp(w), _fp(b), _fp(y)))
        return y


class Batch:
    """n_seq sequences generated in lock-step on one Context."""

    def __init__(self, ctx: Context, voice_ids: Sequence[int], max_len: Sequence[int]):
        self.ctx = ctx
        self.n = len(voice_ids)
        self._h = C.c_void_p()
        v = np.ascontiguousarray(voice_ids, dtype=np.int32)
        m = np.ascontiguousarray(max_len, dtype=np.int32)
        check(lib().ptts_batch_create(ctx._h, self.n, _ip(v), _ip(m), C.byref(self._h)))
        ctx._batches.add(self)
        self.latent_dim = ctx.config.latent_dim
        hop = 1
        for i in range(ctx.config.n_ratios):
            hop *= ctx.config.ratios[i]
        self.frame_samples = hop * ctx.config.upsample_stride

    def close(self):
        if self._h:
            self._staging = None
            self._staging_sets = None
            self._pcm_views = {}
            lib().ptts_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prefill_text(self, ids_per_seq: Sequence[Sequence[int]]):
        assert len(ids_per_seq) == self.n
        offs = np.zeros(self.n + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(x) for x in ids_per_seq])
        flat = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in ids_per_seq])
                                    if offs[-1] else np.zeros(1, dtype=np.int32), dtype=np.int32)
        check(lib().ptts_batch_prefill_text(self._h, _ip(flat), _ip(offs)))

    def warmup_mimi(self, n_frames: int):
        check(lib().ptts_batch_warmup_mimi(self._h, int(n_frames)))

    def step(self, noise: Optional[np.ndarray] = None, want_audio: bool = True):
        """-> (latent [n, L], eos_logit [n], audio [n, 1920] or None)."""
        z = _f32(noise).reshape(self.n, self.latent_dim) if noise is not None else None
        lat = np.empty((self.n, self.latent_dim), dtype=np.float32)
        logit = np.empty((self.n,), dtype=np.float32)
        if getattr(self, "_pcm16", False):       # int16 samples come back through the pinned PCM buffer
            check(lib().ptts_batch_step(self._h, _fp(z), _fp(lat), _fp(logit), None))
            return lat, logit, (self.pcm(0).copy() if want_audio else None)
        audio = np.empty((self.n, self.frame_samples), dtype=np.float32) if want_audio else None
        check(lib().ptts_batch_step(self._h, _fp(z), _fp(lat), _fp(logit), _fp(audio)))
        return lat, logit, audio

    def set_pcm16(self, on: bool = True):
        """16-bit PCM output: the final-sample kernels also store int16 = trunc(clip(v, -1, 1) * 32767) and host steps
        copy those out instead of the fp32 samples; read them with pcm(set).  Before the first frame only."""
        check(lib().ptts_batch_set_pcm16(self._h, 1 if on else 0))
        self._pcm16 = bool(on)
        self._pcm_views = {}

    def pcm(self, k: int = 0) -> np.ndarray:
        """int16 [n, frame_samples] view of the pinned PCM buffer of staging set k."""
        views = getattr(self, "_pcm_views", None)
        if views is None:
            views = self._pcm_views = {}
        if k not in views:
            ptr = C.POINTER(C.c_int16)()
            check(lib().ptts_batch_host_pcm(self._h, int(k), C.byref(ptr)))
            views[k] = np.ctypeslib.as_array(ptr, shape=(self.n, self.frame_samples))
        return views[k]

    def set_pipelined(self, on: bool = True):
        """Throughput mode: step() then returns the audio of the PREVIOUS frame; flush() decodes the last one."""
        check(lib().ptts_batch_set_pipelined(self._h, 1 if on else 0))

    def reset_seq(self, slot: int, voice_id: int, max_len: int):
        """Continuous batching: re-initialise one slot for a new utterance (KV pages, length, BOS, warm Mimi state);
        follow with prefill_text where the other sequences get empty token lists."""
        check(lib().ptts_batch_reset_seq(self._h, int(slot), int(voice_id), int(max_len)))

    def reset_seqs(self, slots: Sequence[int], voice_ids: Sequence[int], max_lens: Sequence[int]):
        """reset_seq for several slots with one synchronisation."""
        a, v, m = (np.ascontiguousarray(x, dtype=np.int32) for x in (slots, voice_ids, max_lens))
        check(lib().ptts_batch_reset_seqs(self._h, len(a), _ip(a), _ip(v), _ip(m)))

    def set_active(self, slot: int, active: bool):
        """Park (False) or resume a slot: a parked slot is still computed but stops growing its KV cache."""
        check(lib().ptts_batch_set_active(self._h, int(slot), 1 if active else 0))

    def flush(self, want_audio: bool = True):
        if getattr(self, "_pcm16", False):
            check(lib().ptts_batch_flush(self._h, None))
            return self.pcm(0).copy() if want_audio else None
        audio = np.empty((self.n, self.frame_samples), dtype=np.float32) if want_audio else None
        check(lib().ptts_batch_flush(self._h, _fp(audio)))
        return audio

    def staging(self):
        """NumPy views of the library's pinned staging buffers: (noise, latent, eos_logit, audio)."""
        if getattr(self, "_staging", None) is None:
            f32p = C.POINTER(C.c_float)
            ptrs = [f32p() for _ in range(4)]
            check(lib().ptts_batch_host_buffers(self._h, *[C.byref(p) for p in ptrs]))
            shapes = [(self.n, self.latent_dim), (self.n, self.latent_dim), (self.n,), (self.n, self.frame_samples)]
            self._staging = tuple(np.ctypeslib.as_array(p, shape=sh) for p, sh in zip(ptrs, shapes))
        return self._staging

    def set_async_staging(self, on: bool = True):
        """Pipelined mode only, before the first frame: frames alternate between two pinned buffer sets so that
        step_staged_async() can be called for frame t+1 while frame t is still running."""
        check(lib().ptts_batch_set_async_staging(self._h, 1 if on else 0))
        self._staging_sets = None

    def staging_sets(self):
        """[(noise, latent, eos_logit, audio)] * 2: frame t reads / writes set t & 1."""
        if getattr(self, "_staging_sets", None) is None:
            f32p = C.POINTER(C.c_float)
            shapes = [(self.n, self.latent_dim), (self.n, self.latent_dim), (self.n,), (self.n, self.frame_samples)]
            sets = []
            for k in range(2):
                ptrs = [f32p() for _ in range(4)]
                check(lib().ptts_batch_host_buffers_set(self._h, k, *[C.byref(p) for p in ptrs]))
                sets.append(tuple(np.ctypeslib.as_array(p, shape=sh) for p, sh in zip(ptrs, shapes)))
            self._staging_sets = sets
        return self._staging_sets

    def step_staged_async(self) -> int:
        """Enqueue one frame (noise taken from the current set); returns the set index to staged_wait() on."""
        k = C.c_int32(0)
        check(lib().ptts_batch_step_staged_async(self._h, C.byref(k)))
        return k.value

    def staged_wait(self, k: int):
        check(lib().ptts_batch_staged_wait(self._h, int(k)))

    def step_staged(self):
        """One frame with zero host copies: fill staging()[0] with N(0,1) noise, call, read staging()[1:]."""
        check(lib().ptts_batch_step_staged(self._h))

    def step_device(self):
        check(lib().ptts_batch_step_device(self._h))

    def set_prev_latent(self, latent: np.ndarray):
        a = _f32(latent).reshape(self.n, self.latent_dim)
        check(lib().ptts_batch_set_prev_latent(self._h, _fp(a)))

    def seed(self, seed: int):
        check(lib().ptts_batch_seed(self._h, C.c_uint64(seed)))

    def lengths(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int32)
        check(lib().ptts_batch_lengths(self._h, _ip(out)))
        return out

    def mimi_decode(self, latents: np.ndarray, want_audio: bool = True, out: Optional[np.ndarray] = None):
        """Latents [n, F, latent_dim] -> waveforms [n, F * frame_samples] (models/mimi.py:70-75 over F frames); `out`
        re-uses a result array of that shape (float32, C-contiguous) instead of allocating one."""
        a = _f32(latents).reshape(self.n, -1, self.latent_dim)
        f = a.shape[1]
        if want_audio and out is not None:
            if out.shape != (self.n, f * self.frame_samples) or out.dtype != np.float32 or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float32 array of shape (n, F * frame_samples)")
        elif want_audio:
            out = _big_empty((self.n, f * self.frame_samples), np.float32)
        else:
            out = None
        check(lib().ptts_batch_mimi_decode(self._h, _fp(a), f, _fp(out)))
        return out

    def profile_sections(self):
        """In-graph milliseconds per frame section (perturbs the streaming state; profiling only)."""
        ms = (C.c_float * 8)()
        n = check(lib().ptts_batch_profile_sections(self._h, ms, 8))
        names = ["flow_backbone", "eos_flow_head", "mimi_transformer", "seanet", "whole_frame",
                 "mimi_transformer_capped", "seanet_capped"]
        return dict(zip(names, [float(ms[i]) for i in range(n)]))

    def profile_step(self):
        """One eager frame with per-kernel CUDA-event timing -> list of dicts sorted by time."""
        rep = C.c_char_p()
        check(lib().ptts_batch_profile_step(self._h, C.byref(rep)))
        rows = []
        for line in rep.value.decode().strip().splitlines():
            name, n, ms, fl, by = line.rsplit(",", 4)
            rows.append({"kernel": name, "launches": int(n), "ms": float(ms), "flops": float(fl), "bytes": float(by)})
        return rows


def max_gen_len(n_tok: int, frame_rate: float = 12.5) -> int:
    """ceil((n_tok/3 + 2) * frame_rate), the reference's `_estimate_max_gen_len` (tts_model.py:440-444)."""
    return math.ceil((n_tok / 3.0 + 2.0) * frame_rate)
