"""`TTSModel`: the reference's public API for streaming generation, backed by libptts_b200 (CUDA, sm_100a).

Drop-in surface (reference `pocket_tts_mlx/models/tts_model.py`): `load_model` (:202-221),
`get_state_for_audio_prompt` (:484-518, predefined-voice branch), `generate_audio` (:308-334),
`generate_audio_stream` (:336-361), properties `device` / `sample_rate` (:79-85) and the attributes
`config, temp, lsd_decode_steps, noise_clamp, eos_threshold, has_voice_cloning` (:70-77).  Host logic kept
in Python (sentence chunking, EOS bookkeeping, max-length estimate, trim/fade) follows :336-462; all
arithmetic happens on the GPU through the C ABI in include/ptts.h.  There is no CPU fallback.

Extensions (keyword-only, all optional): `precision`, `device_id`, `kv_pool_tokens` on `load_model`;
`noise` / `seed` on the generate calls (injectable flow noise for parity tests);
`generate_audio_batch` for many utterances in lock-step on one GPU.
"""

from __future__ import annotations

import logging
import os
import time
import weakref
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, Generator, List, Optional, Sequence, Union

import numpy as np

from . import _native
from .assets import resolve_asset
from .config import Config, load_config
from .default_parameters import (
    DEFAULT_EOS_THRESHOLD,
    DEFAULT_LSD_DECODE_STEPS,
    DEFAULT_NOISE_CLAMP,
    DEFAULT_TEMPERATURE,
    DEFAULT_VARIANT,
    MAX_TOKEN_PER_CHUNK,
)
from .safetensors_io import read_safetensors
from .text import SentencePieceTokenizer, prepare_text_prompt, split_into_best_sentences

logger = logging.getLogger(__name__)

_VOICE_NAMES = ["alba", "marius", "javert", "jean", "fantine", "cosette", "eponine", "azelma"]
PREDEFINED_VOICES = {
    v: f"hf://kyutai/pocket-tts-without-voice-cloning/embeddings/{v}.safetensors@d4fdd22ae8c8e1cb3634e150ebeff1dab2d16df3"
    for v in _VOICE_NAMES
}

VOICE_CLONING_UNSUPPORTED = (
    "Voice cloning from an audio file needs the Mimi encoder, which this build does not include "
    f"(SURVEY.md 8f-2). Use one of the predefined voices {list(PREDEFINED_VOICES)} or pass a precomputed "
    "conditioning array to get_state_for_conditioning()."
)


class VoiceHandle:
    """Owner of one voice's KV prefix pages on the GPU.  The reference's model state is a dict that the garbage
    collector frees; here the state dict carries this handle, and when the last state (or copy of it) that refers
    to the handle dies, the pages go back to the pool (`ptts_voice_destroy`; slots of live batches that still attend
    the prefix keep it alive on the C side).  `copy.deepcopy(state)` shares the handle: the prefix is immutable."""

    def __init__(self, ctx: "_native.Context", voice_id: int):
        self.voice_id = int(voice_id)
        self._fin = weakref.finalize(self, VoiceHandle._release, weakref.ref(ctx), self.voice_id)

    @staticmethod
    def _release(ctx_ref, voice_id):
        ctx = ctx_ref()
        if ctx is not None and ctx._h:
            try:
                ctx.voice_destroy(voice_id)
            except Exception:      # the context may already be closing
                pass

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self


class TTSModel:
    _TOKENS_PER_SECOND_ESTIMATE = 3.0
    _GEN_SECONDS_PADDING = 2.0
    _MIMI_WARMUP_FRAMES = 1

    def __init__(self, config: Config, ctx: "_native.Context", tokenizer: SentencePieceTokenizer, temp: float,
                 lsd_decode_steps: int, noise_clamp: Optional[float], eos_threshold: float, precision: str,
                 weights_file: Path):
        self.config = config
        self.temp = temp
        self.lsd_decode_steps = lsd_decode_steps
        self.noise_clamp = noise_clamp
        self.eos_threshold = eos_threshold
        self.has_voice_cloning = bool(ctx.has_voice_cloning)     # the checkpoint carried the Mimi encoder
        self.precision = precision
        self._ctx = ctx
        self._tokenizer = tokenizer
        self._weights_file = Path(weights_file)
        self._state_cache: "OrderedDict" = OrderedDict()     # _cached_get_state_for_audio_prompt (LRU, 2 entries)

    # ------------------------------------------------------------------ properties
    @property
    def device(self) -> str:
        return f"cuda:{self._ctx.device}"

    @property
    def sample_rate(self) -> int:
        return self.config.mimi.sample_rate

    @property
    def frame_samples(self) -> int:
        return int(self.config.mimi.sample_rate / self.config.mimi.frame_rate)

    # ------------------------------------------------------------------ loading
    @classmethod
    def load_model(cls, config: Union[str, Path] = DEFAULT_VARIANT, temp: float = DEFAULT_TEMPERATURE,
                   lsd_decode_steps: int = DEFAULT_LSD_DECODE_STEPS,
                   noise_clamp: Optional[float] = DEFAULT_NOISE_CLAMP,
                   eos_threshold: float = DEFAULT_EOS_THRESHOLD, *, precision: str = "bf16",
                   device_id: int = 0, kv_pool_tokens: int = 262144, max_batch: int = 0) -> "TTSModel":
        if str(config).endswith(".yaml"):
            cfg = load_config(Path(config))
            logger.info("Loading model from config at %s...", config)
        else:
            cfg = load_config(Path(__file__).parent / "config" / f"{config}.yaml")
        if cfg.flow_lm.weights_path is not None and cfg.mimi.weights_path is None:
            raise ValueError("If you specify flow_lm.weights_path you should specify mimi.weights_path")
        if cfg.mimi.weights_path is not None and cfg.flow_lm.weights_path is None:
            raise ValueError("If you specify mimi.weights_path you should specify flow_lm.weights_path")
        if cfg.flow_lm.weights_path is not None:
            raise ValueError("split flow_lm/mimi checkpoints are not supported by this build; use weights_path")
        if cfg.weights_path is None:
            raise ValueError("config has no weights_path: an uninitialised model cannot be run")
        try:
            weights_file = resolve_asset(cfg.weights_path)
            if not weights_file.exists():
                raise FileNotFoundError(str(weights_file))
        except FileNotFoundError:
            if cfg.weights_path_without_voice_cloning is None:
                raise
            weights_file = resolve_asset(cfg.weights_path_without_voice_cloning)
        tokenizer = SentencePieceTokenizer(cfg.flow_lm.lookup_table.n_bins,
                                           resolve_asset(cfg.flow_lm.lookup_table.tokenizer_path))
        ccfg = _native.make_config(cfg, temp, lsd_decode_steps, noise_clamp, eos_threshold, precision,
                                   kv_pool_tokens, max_batch)
        ctx = _native.Context(ccfg, device_id)
        loaded = skipped = 0
        for key, arr in read_safetensors(weights_file).items():
            if arr.dtype.kind != "f":
                skipped += 1
                continue
            if ctx.load_weight(key, arr):
                loaded += 1
            else:
                skipped += 1
        ctx.finalize()
        unused = ctx.unused_weights()
        logger.info("Loaded %d weights, skipped %d", loaded - len(unused), skipped + len(unused))
        if unused:
            logger.debug("checkpoint keys nothing consumed: %s", unused)
        model = cls(cfg, ctx, tokenizer, temp, lsd_decode_steps, noise_clamp, eos_threshold, precision, weights_file)
        model.unused_checkpoint_keys = unused
        return model

    # ------------------------------------------------------------------ voices
    def _voice_file(self, name: str) -> Path:
        import os
        cands = []
        if os.environ.get("POCKET_TTS_VOICES_DIR"):
            cands.append(Path(os.environ["POCKET_TTS_VOICES_DIR"]) / f"{name}.safetensors")
        cands.append(self._weights_file.parent / "embeddings" / f"{name}.safetensors")
        for c in cands:
            if c.exists():
                return c
        return resolve_asset(PREDEFINED_VOICES[name])

    def get_state_for_conditioning(self, conditioning: np.ndarray) -> Dict:
        """Prefill an already-projected voice conditioning [T, d_model] (or [1, T, d_model])."""
        cond = np.asarray(conditioning, dtype=np.float32)
        d = self.config.flow_lm.transformer.d_model
        cond = cond.reshape(-1, d)
        t0 = time.monotonic()
        vid = self._ctx.voice_create(cond)
        logger.info("Prompting audio took %d ms", int((time.monotonic() - t0) * 1000))
        return {"voice_id": vid, "prompt_len": int(cond.shape[0]), "_handle": VoiceHandle(self._ctx, vid)}

    def _cached_get_state_for_audio_prompt(self, audio_conditioning, truncate: bool = False) -> Dict:
        """The reference's `lru_cache(maxsize=2)` variant (models/tts_model.py:478-482): repeated prompts (a server that
        speaks with the same few voices) reuse the prefilled state instead of prefilling -- and pinning KV pages -- again.
        Keys are hashable prompts only (voice names, paths), like the reference's."""
        key = (str(audio_conditioning) if isinstance(audio_conditioning, Path) else audio_conditioning, bool(truncate))
        hash(key)
        if key in self._state_cache:
            self._state_cache.move_to_end(key)
            return self._state_cache[key]
        state = self.get_state_for_audio_prompt(audio_conditioning, truncate)
        self._state_cache[key] = state
        while len(self._state_cache) > 2:
            self._state_cache.popitem(last=False)     # the evicted state's pages are freed once nobody holds it
        return state

    def get_state_for_audio_prompt(self, audio_conditioning, truncate: bool = False) -> Dict:
        if isinstance(audio_conditioning, str) and audio_conditioning in PREDEFINED_VOICES:
            tensors = read_safetensors(self._voice_file(audio_conditioning))
            prompt = tensors.get("audio_prompt")
            if prompt is None:
                raise KeyError("audio_prompt not found in voice embedding file")
            return self.get_state_for_conditioning(prompt)
        if isinstance(audio_conditioning, (str, Path)):
            if isinstance(audio_conditioning, str) and not Path(audio_conditioning).exists() \
                    and "/" not in audio_conditioning and "." not in audio_conditioning:
                raise ValueError(
                    f"Predefined voice '{audio_conditioning}' not found, available voices are {list(PREDEFINED_VOICES)}.")
            if not self.has_voice_cloning:
                raise ValueError(VOICE_CLONING_UNSUPPORTED)
            # voice cloning from a file (reference tts_model.py:493-508): read, optional 30 s truncation, mono 24 kHz
            from .audio import audio_read, convert_audio
            audio, rate = audio_read(Path(audio_conditioning))
            if truncate:
                max_samples = int(30 * rate)
                if audio.shape[-1] > max_samples:
                    audio = audio[..., :max_samples]
                    logger.info("Audio truncated to 30 seconds")
            audio_conditioning = convert_audio(audio, rate, self.config.mimi.sample_rate, 1)
        # an array is a waveform [1, T] / [T] at the model sample rate, like the reference's mx.array branch
        if not self.has_voice_cloning:
            raise ValueError(VOICE_CLONING_UNSUPPORTED)
        wave_ = np.asarray(audio_conditioning, dtype=np.float32)
        if wave_.ndim == 2:
            if wave_.shape[0] != 1:
                raise ValueError("voice cloning takes mono audio [1, T]")
            wave_ = wave_[0]
        if wave_.ndim != 1 or wave_.shape[0] == 0:
            raise ValueError("voice cloning takes a non-empty waveform [T] or [1, T]")
        return self.get_state_for_conditioning(self.encode_audio(wave_))

    def encode_audio(self, audio: np.ndarray) -> np.ndarray:
        """Waveform (mono, model sample rate) -> FlowLM conditioning [T_v, d_model]: the reference's
        `_encode_audio` (tts_model.py:271-276) = Mimi encoder + speaker projection, on the GPU."""
        frame = int(self.config.mimi.sample_rate / self.config.mimi.frame_rate)
        return self._ctx.encode_audio(audio, frame)

    # ------------------------------------------------------------------ generation
    def _estimate_max_gen_len(self, token_count: int) -> int:
        return _native.max_gen_len(token_count, self.config.mimi.frame_rate)

    def generate_audio(self, model_state: Dict, text_to_generate: str, max_tokens: int = MAX_TOKEN_PER_CHUNK,
                       frames_after_eos: Optional[int] = None, copy_state: bool = True, trim_start_ms: int = 0,
                       fade_in_ms: int = 0, warmup_frames: int = _MIMI_WARMUP_FRAMES, *,
                       noise: Optional[np.ndarray] = None, seed: Optional[int] = None) -> np.ndarray:
        chunks = list(self.generate_audio_stream(model_state, text_to_generate, max_tokens=max_tokens,
                                                 frames_after_eos=frames_after_eos, copy_state=copy_state,
                                                 warmup_frames=warmup_frames, noise=noise, seed=seed))
        audio = np.concatenate(chunks, axis=0) if chunks else np.zeros(0, dtype=np.float32)
        return postprocess_audio_start(audio, self.sample_rate, trim_start_ms, fade_in_ms)

    def generate_audio_stream(self, model_state: Dict, text_to_generate: str, max_tokens: int = MAX_TOKEN_PER_CHUNK,
                              frames_after_eos: Optional[int] = None, copy_state: bool = True,
                              warmup_frames: int = _MIMI_WARMUP_FRAMES, *, noise: Optional[np.ndarray] = None,
                              seed: Optional[int] = None, pcm16: bool = False) -> Generator[np.ndarray, None, None]:
        """Yields one 1920-sample frame at a time.  pcm16=True yields int16 frames converted on the GPU in the kernels
        that produce the samples (the format `StreamingWAVWriter` / `stream_audio_chunks` write, data/audio.py:64-70)."""
        if not copy_state:
            # the reference appends the utterance to the passed state (later calls then attend to it); here a voice
            # prefix is immutable shared KV pages, so every chunk starts from the prefix alone
            logger.warning("copy_state=False is not supported: the voice state is immutable and is never modified")
        chunks = split_into_best_sentences(self._tokenizer, text_to_generate, max_tokens)
        noise_rows = None if noise is None else np.asarray(noise, dtype=np.float32).reshape(-1, self._ctx.config.latent_dim)
        cursor = 0
        for ci, chunk in enumerate(chunks):
            _, guess = prepare_text_prompt(chunk)
            effective = frames_after_eos if frames_after_eos is not None else guess + 2
            sub_noise = None if noise_rows is None else noise_rows[cursor:]
            used = [0]
            yield from self._generate_chunk(model_state, chunk, effective, warmup_frames, sub_noise,
                                            None if seed is None else seed + ci, used, pcm16)
            cursor += used[0]

    def _generate_chunk(self, model_state: Dict, chunk: str, frames_after_eos: int, warmup_frames: int,
                        noise_rows: Optional[np.ndarray], seed: Optional[int], used: List[int], pcm16: bool = False):
        tokens = self._tokenizer.encode(chunk)
        n_tok = int(tokens.shape[0])
        max_gen_len = self._estimate_max_gen_len(n_tok)
        required = int(model_state["prompt_len"]) + n_tok + max_gen_len
        batch = _native.Batch(self._ctx, [int(model_state["voice_id"])], [required])
        try:
            if seed is None:
                seed = int(np.random.SeedSequence().entropy % (1 << 63))
            batch.seed(seed)
            if pcm16:
                batch.set_pcm16(True)
            batch.warmup_mimi(warmup_frames)
            t_gen = time.monotonic()
            batch.prefill_text([tokens])
            used[0] = 1                                  # the reference draws (and discards) noise here
            eos_step = None
            produced = 0
            for step in range(max_gen_len):
                z = None
                if noise_rows is not None:
                    if used[0] >= noise_rows.shape[0]:
                        raise ValueError("injected noise has fewer rows than generated frames")
                    z = noise_rows[used[0]][None, :]
                    used[0] += 1
                _, logit, audio = batch.step(z, want_audio=True)
                if bool(logit[0] > self.eos_threshold) and eos_step is None:
                    eos_step = step
                if eos_step is not None and step >= eos_step + frames_after_eos:
                    break
                produced += audio.shape[1]
                yield audio[0].copy()
            ms_audio = int(produced * 1000 / self.sample_rate)
            ms_gen = int((time.monotonic() - t_gen) * 1000)
            logger.info("Generated: %d ms of audio in %d ms so %.2fx faster than real-time", ms_audio, ms_gen,
                        ms_audio / max(1, ms_gen))
        finally:
            batch.close()

    def generate_audio_batch(self, model_states: Sequence[Dict], token_ids: Sequence[Sequence[int]],
                             frames_after_eos: Union[int, Sequence[int]] = 3,
                             warmup_frames: int = _MIMI_WARMUP_FRAMES, max_frames: Optional[int] = None,
                             noise: Optional[np.ndarray] = None, seed: int = 0,
                             return_latents: bool = False, pipelined: bool = True, pcm16: bool = False):
        """Lock-step generation of many single-chunk utterances (token ids already prepared).

        pcm16=True returns int16 waveforms (`trunc(clip(x, -1, 1) * 32767)`, the reference's streaming-WAV sample
        format, data/audio.py:70) converted on the GPU where the samples are produced; half the device->host bytes.

        noise: optional [1 + max_frames, n, latent_dim] (row 0 is the unused text-prefill draw, as in the
        reference); without it the host draws N(0,1) from `seed`.  pipelined=True overlaps the Mimi decode of
        frame t-1 with the FlowLM step of frame t on the GPU (same results, audio arrives one step later).
        Returns a list of 1-D float32 waveforms (and per-sequence latents when asked)."""
        n = len(model_states)
        if n == 0:
            return ([], []) if return_latents else []
        fae = [frames_after_eos] * n if isinstance(frames_after_eos, int) else list(frames_after_eos)
        n_tok = [len(t) for t in token_ids]
        limits = [self._estimate_max_gen_len(k) for k in n_tok]
        if max_frames is not None:
            limits = [min(l, max_frames) for l in limits]
        # pipelined mode runs one frame ahead of the host's bookkeeping (asynchronous staged steps): a sequence is
        # parked one frame after its last accepted one, hence the slack in its KV reservation
        req = [int(s["prompt_len"]) + k + l + (2 if pipelined else 0) for s, k, l in zip(model_states, n_tok, limits)]
        batch = _native.Batch(self._ctx, [int(s["voice_id"]) for s in model_states], req)
        rng = np.random.Generator(np.random.PCG64(seed))
        ldim = self._ctx.config.latent_dim
        try:
            batch.seed(seed)
            if pipelined:
                batch.set_pipelined(True)
                batch.set_async_staging(True)
            if pcm16:
                batch.set_pcm16(True)
            batch.warmup_mimi(warmup_frames)
            batch.prefill_text(token_ids)
            # Lock-step bookkeeping on whole arrays: every sequence accepts frames 0 .. n_acc-1, so the per-step
            # [n, .] blocks returned by step() are kept as they are and sliced per sequence at the end.
            lim = np.asarray(limits, dtype=np.int64)
            fae_a = np.asarray(fae, dtype=np.int64)
            eos_step = np.full(n, -1, dtype=np.int64)
            done = np.zeros(n, dtype=bool)
            n_acc = np.zeros(n, dtype=np.int64)
            n_steps = int(lim.max()) if n else 0
            # sequence-major output arrays: step blocks are written straight into them, and a sequence's waveform /
            # latents are a contiguous slice (returned as a view, no concatenation at the end)
            aud_all = _alloc_output((n, n_steps, self.frame_samples), np.int16 if pcm16 else np.float32)
            lat_all = np.empty((n, n_steps, ldim), dtype=np.float32)
            counts = {"lat": 0, "audio": 0}
            thr = self.eos_threshold
            # A frame's [n, 1920] block becomes column `k` of the sequence-major array: an n-way strided scatter into
            # memory that is touched for the first time (page faults).  For large batches that is ~1 ms of host work per
            # frame, so it is taken off the thread that feeds the GPU: the block is copied (contiguously) out of the
            # pinned staging buffer into a small ring and worker threads scatter it.
            scatter = _BlockScatter(aud_all) if n * self.frame_samples * aud_all.itemsize >= (1 << 18) else None

            def put_audio(block):
                if scatter is not None:
                    scatter.put(counts["audio"], block)
                else:
                    aud_all[:, counts["audio"], :] = block
                counts["audio"] += 1

            def account(step, lat, logit):
                """EOS / frame-budget bookkeeping of frame `step` (reference tts_model.py:404-426), all sequences."""
                lat_all[:, counts["lat"], :] = lat
                counts["lat"] += 1
                live = ~done
                first = live & (eos_step < 0) & (logit > thr)
                eos_step[first] = step
                stop = live & (eos_step >= 0) & (step >= eos_step + fae_a)   # the reference breaks before this frame
                accept = live & ~stop
                n_acc[accept] = step + 1
                newly = stop | (accept & (step + 1 >= lim))                  # ... or its frame budget is used up
                done[newly] = True
                if not done.all():
                    for b in np.nonzero(newly)[0]:
                        batch.set_active(int(b), False)     # parked: computed with the batch, KV no longer grows

            if pipelined:
                # frame s is enqueued before frame s-1 is read back: the GPU never waits for the host
                sets = batch.staging_sets()
                aud_of = (lambda k: batch.pcm(k)) if pcm16 else (lambda k: sets[k][3])
                enq = 0
                for step in range(n_steps):
                    zbuf = sets[step & 1][0]
                    if noise is None:
                        rng.standard_normal(zbuf.shape, dtype=np.float32, out=zbuf)
                    else:
                        zbuf[...] = np.asarray(noise[1 + step], dtype=np.float32)
                    batch.step_staged_async()
                    enq = step + 1
                    if step >= 1:
                        k = (step - 1) & 1
                        batch.staged_wait(k)
                        if step >= 2:
                            put_audio(aud_of(k))                             # audio of frame step-2
                        account(step - 1, sets[k][1], sets[k][2].copy())
                        if done.all():
                            break
                if enq:
                    k = (enq - 1) & 1
                    batch.staged_wait(k)
                    if enq >= 2:
                        put_audio(aud_of(k))                                 # audio of frame enq-2
                    if not done.all():
                        account(enq - 1, sets[k][1], sets[k][2].copy())
                    if counts["audio"] < counts["lat"]:
                        put_audio(batch.flush())                             # audio of the last accounted frame
            else:
                for step in range(n_steps):
                    z = rng.standard_normal((n, ldim), dtype=np.float32) if noise is None \
                        else np.asarray(noise[1 + step], dtype=np.float32)
                    lat, logit, audio = batch.step(z, want_audio=True)
                    put_audio(audio)
                    account(step, lat, logit)
                    if done.all():
                        break
            if scatter is not None:
                scatter.close()
                scatter = None
            waves = [aud_all[b, :int(n_acc[b])].reshape(-1) for b in range(n)]
            lats = [lat_all[b, :int(n_acc[b])] for b in range(n)]
            if return_latents:
                return waves, lats
            return waves
        finally:
            if scatter is not None:
                scatter.close()
            batch.close()

    def generate_audio_continuous(self, model_states: Sequence[Dict], token_ids: Sequence[Sequence[int]],
                                  slots: int = 256, frames_after_eos: Union[int, Sequence[int]] = 3,
                                  warmup_frames: int = _MIMI_WARMUP_FRAMES, max_frames: Optional[int] = None,
                                  noise: Optional[Sequence[np.ndarray]] = None, seed: int = 0,
                                  return_latents: bool = False, min_admit: Optional[int] = None, pipelined: bool = True,
                                  _scheduled: bool = False):
        """Continuous batching over `slots` lock-step sequences: when an utterance ends (EOS rule or frame limit of
        the reference, tts_model.py:404-426) its slot is parked and later re-initialised for the next queued
        utterance while the others keep decoding, so the batch stays full.  Admissions are
        grouped: freed slots wait until `min_admit` of them (default slots/16) can be prefilled together, which turns
        the text prefill into one dense GEMM pass instead of one GEMV pass per utterance.  Results per utterance are
        the same as decoding it in a batch of its own.  pipelined=True uses the two-branch frame graph (FlowLM step t next
        to the Mimi decode of frame t-1): the audio of a frame then arrives one step after its latent, also across the
        re-use of a slot.

        noise: optional per-utterance arrays [1 + frames, latent_dim] (row 0 = the unused prefill draw, like the
        reference); without it the host draws one N(0,1) block per step from `seed`.
        Utterances are started longest frame budget first (shortest tail); the waveforms (and the per-utterance
        latents when asked) come back in input order."""
        n_jobs = len(model_states)
        if n_jobs == 0:
            return ([], []) if return_latents else []
        fae = [frames_after_eos] * n_jobs if isinstance(frames_after_eos, int) else list(frames_after_eos)
        n_tok = [len(t) for t in token_ids]
        limits = [self._estimate_max_gen_len(k) for k in n_tok]
        if max_frames is not None:
            limits = [min(l, max_frames) for l in limits]
        if not _scheduled and n_jobs > int(slots) and min(limits) != max(limits):
            # longest budget first: the run then ends with the SHORT utterances, so the tail in which the queue is
            # empty and the slots drain one by one is as short as it can be; results go back in submission order
            order = sorted(range(n_jobs), key=lambda j: -limits[j])
            r = self.generate_audio_continuous(
                [model_states[j] for j in order], [token_ids[j] for j in order], slots=slots,
                frames_after_eos=[fae[j] for j in order], warmup_frames=warmup_frames, max_frames=max_frames,
                noise=None if noise is None else [noise[j] for j in order], seed=seed, return_latents=return_latents,
                min_admit=min_admit, pipelined=pipelined, _scheduled=True)
            inv = [0] * n_jobs
            for pos, j in enumerate(order):
                inv[j] = pos
            if return_latents:
                return [r[0][inv[j]] for j in range(n_jobs)], [r[1][inv[j]] for j in range(n_jobs)]
            return [r[inv[j]] for j in range(n_jobs)]
        need = [int(s["prompt_len"]) + k + l for s, k, l in zip(model_states, n_tok, limits)]
        # host memory: the frames of all utterances of one run are kept in one slot-major array (waveforms are views
        # into it); very long job lists are cut into runs of at most ~max_host_gb of output each
        max_host_gb = float(os.environ.get("PTTS_CONT_MAX_GB", "6"))
        bytes_per_frame = 4.0 * (self.frame_samples + self._ctx.config.latent_dim)
        if n_jobs > int(slots) and sum(limits) * bytes_per_frame * 1.3 > max_host_gb * 2 ** 30:
            per_run = max(int(slots), int(n_jobs * max_host_gb * 2 ** 30 / (sum(limits) * bytes_per_frame * 1.3)))
            per_run -= per_run % int(slots)              # whole waves of the slots: a run does not end with a few stragglers
            waves_all, lats_all = [], []
            for lo in range(0, n_jobs, per_run):
                hi = min(n_jobs, lo + per_run)
                r = self.generate_audio_continuous(
                    model_states[lo:hi], token_ids[lo:hi], slots=slots, frames_after_eos=fae[lo:hi],
                    warmup_frames=warmup_frames, max_frames=max_frames,
                    noise=None if noise is None else noise[lo:hi], seed=seed + lo, return_latents=True,
                    min_admit=min_admit, pipelined=pipelined, _scheduled=True)
                waves_all += r[0]
                lats_all += r[1]
            return (waves_all, lats_all) if return_latents else waves_all
        n_slots = min(int(slots), n_jobs)
        cap = max(need) + 2                               # any utterance fits any slot; +2: a slot is parked one frame late
        group = max(1, n_slots // 16) if min_admit is None else max(1, int(min_admit))
        ldim = self._ctx.config.latent_dim
        batch = _native.Batch(self._ctx, [int(model_states[j]["voice_id"]) for j in range(n_slots)], [cap] * n_slots)
        rng = np.random.Generator(np.random.PCG64(seed))
        try:
            batch.seed(seed)
            if pipelined:
                batch.set_pipelined(True)
            batch.set_async_staging(True)                 # frame t+1 is enqueued before frame t is read back
            sets = batch.staging_sets()
            lag = 1 if pipelined else 0                   # audio block that holds the frame whose latent is in block t: t + lag
            batch.warmup_mimi(warmup_frames)
            batch.prefill_text([token_ids[j] for j in range(n_slots)])
            # Whole-array bookkeeping.  Per slot: the utterance it holds (-1 = parked), the frame index its next
            # enqueued step produces, the frame of the first EOS crossing (-1 = none yet), its frame budget / EOS tail.
            job = np.arange(n_slots, dtype=np.int64)
            frame = np.zeros(n_slots, dtype=np.int64)
            eos_at = np.full(n_slots, -1, dtype=np.int64)
            lim_all = np.asarray(limits, dtype=np.int64)
            fae_all = np.asarray(fae, dtype=np.int64)
            # Per utterance: slot, index of the processed block holding its frame 0, accepted frames.  Every enqueued
            # step is read back exactly once and in order, so an utterance's frames sit in consecutive blocks.
            slot_of = np.full(n_jobs, -1, dtype=np.int64)
            t0_of = np.zeros(n_jobs, dtype=np.int64)
            n_of = np.zeros(n_jobs, dtype=np.int64)
            slot_of[:n_slots] = np.arange(n_slots)
            free: List[int] = []                          # parked slots waiting for the next admission round
            next_job = n_slots
            est_steps = int(np.ceil(lim_all.sum() / n_slots * 1.25)) + int(lim_all.max()) + 8 + lag
            # slot-major, so that an utterance's waveform is one contiguous slice (returned as a view, no final gather)
            aud = np.empty((n_slots, est_steps, self.frame_samples), dtype=np.float32)   # pages are touched on write
            lat_b = np.empty((n_slots, est_steps, ldim), dtype=np.float32)
            thr = self.eos_threshold
            state = {"live": n_slots, "enq": 0, "done_blocks": 0}

            def enqueue():
                """Fill the noise of the next frame, enqueue it, remember which (utterance, frame) each slot computes."""
                z = sets[state["enq"] & 1][0]
                if noise is None:
                    rng.standard_normal(z.shape, dtype=np.float32, out=z)
                else:
                    for s_ in np.nonzero(job >= 0)[0]:      # (the frame enqueued past an utterance's last one is discarded)
                        nz = noise[int(job[s_])]
                        z[s_] = nz[min(1 + int(frame[s_]), len(nz) - 1)]
                k = batch.step_staged_async()
                snap = (k, job.copy(), frame.copy())
                frame[job >= 0] += 1
                state["enq"] += 1
                return snap

            def process(snap):
                """Read one finished frame back and apply the reference's EOS / frame-budget rule to every slot."""
                nonlocal aud, lat_b
                k, jobs, frames = snap
                batch.staged_wait(k)
                t = state["done_blocks"]
                if t >= aud.shape[1]:                    # estimate exceeded: grow
                    aud = np.concatenate([aud, np.empty_like(aud)], axis=1)
                    lat_b = np.concatenate([lat_b, np.empty_like(lat_b)], axis=1)
                aud[:, t] = sets[k][3]
                lat_b[:, t] = sets[k][1]
                logit = sets[k][2]
                state["done_blocks"] = t + 1
                valid = (jobs >= 0) & (job == jobs)      # not parked, and not ended while this frame was in flight
                jj = np.where(valid, jobs, 0)
                first = valid & (eos_at < 0) & (logit > thr)
                eos_at[first] = frames[first]
                stop = valid & (eos_at >= 0) & (frames >= eos_at + fae_all[jj])    # the reference breaks before this frame
                accept = valid & ~stop
                n_of[jobs[accept]] += 1
                ended = stop | (accept & (frames + 1 >= lim_all[jj]))
                for s_ in np.nonzero(ended)[0]:
                    batch.set_active(int(s_), False)
                    job[s_] = -1
                    free.append(int(s_))
                    state["live"] -= 1

            inflight = None
            while state["live"] > 0 or inflight is not None:
                nxt = enqueue() if state["live"] > 0 else None
                if inflight is not None:
                    process(inflight)
                inflight = nxt
                pending = n_jobs - next_job
                if pending > 0 and free and (len(free) >= min(group, pending) or state["live"] == 0):
                    if inflight is not None:              # drain before slots are re-initialised
                        process(inflight)
                        inflight = None
                    take = free[:pending]
                    del free[:len(take)]
                    jobs_new = list(range(next_job, next_job + len(take)))
                    next_job += len(take)
                    batch.reset_seqs(take, [int(model_states[j]["voice_id"]) for j in jobs_new], [cap] * len(take))
                    toks: List[Sequence[int]] = [[] for _ in range(n_slots)]
                    for s_, j in zip(take, jobs_new):
                        job[s_], frame[s_], eos_at[s_] = j, 0, -1
                        slot_of[j], t0_of[j] = s_, state["enq"]      # its frame 0 is the next enqueued step
                        toks[s_] = token_ids[j]
                    batch.prefill_text(toks)
                    state["live"] += len(take)
            if pipelined and state["done_blocks"] > 0:
                # the frames whose latents came with the last step are still to be decoded
                t = state["done_blocks"]
                if t >= aud.shape[1]:
                    aud = np.concatenate([aud, np.empty((n_slots, 1, self.frame_samples), dtype=aud.dtype)], axis=1)
                aud[:, t] = batch.flush()
            waves, lats = [], []
            for j in range(n_jobs):
                s_, t0, k = int(slot_of[j]), int(t0_of[j]), int(n_of[j])
                waves.append(aud[s_, t0 + lag:t0 + lag + k].reshape(-1))
                lats.append(lat_b[s_, t0:t0 + k])
            if return_latents:
                return waves, lats
            return waves
        finally:
            batch.close()

    def generate_audio_sharded(self, model_states: Sequence[Dict], token_ids: Sequence[Sequence[int]], *,
                               rank: Optional[int] = None, world_size: Optional[int] = None, slots: int = 256,
                               gather: bool = False, **kwargs):
        """Multi-GPU entry point (one process per GPU, e.g. under torchrun): decode this rank's share of the utterance
        set on this model's GPU; see `sharding.generate_sharded`.  rank / world_size default to the launcher's
        environment (RANK / WORLD_SIZE)."""
        from .sharding import dist_env, generate_sharded
        w, r, _ = dist_env()
        return generate_sharded(self, model_states, token_ids, rank=r if rank is None else rank,
                                world_size=w if world_size is None else world_size, slots=slots, gather=gather, **kwargs)

    def close(self):
        self._state_cache.clear()
        self._ctx.close()


_alloc_output = _native._big_empty


class _BlockScatter:
    """Writes per-frame blocks [n, w] into column k of a sequence-major array [n, steps, w] on worker threads."""

    RING = 8

    def __init__(self, dst: np.ndarray, workers: int = 3):
        self.dst = dst
        self.ring = np.empty((self.RING,) + (dst.shape[0], dst.shape[2]), dtype=dst.dtype)
        self.futs = [None] * self.RING
        self.pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="ptts-scatter")

    def _scatter(self, k: int, slot: int):
        self.dst[:, k, :] = self.ring[slot]

    def put(self, k: int, block: np.ndarray):
        slot = k % self.RING
        if self.futs[slot] is not None:
            self.futs[slot].result()
        np.copyto(self.ring[slot], block)
        self.futs[slot] = self.pool.submit(self._scatter, k, slot)

    def close(self):
        for i, f in enumerate(self.futs):
            if f is not None:
                f.result()
                self.futs[i] = None
        self.pool.shutdown(wait=True)


def postprocess_audio_start(audio: np.ndarray, sample_rate: int, trim_start_ms: int = 0, fade_in_ms: int = 0):
    """Trim the first `trim_start_ms` (only if 0 < trim < len) and apply a linear fade-in whose ramp
    includes both endpoints (reference `_postprocess_audio_start`, tts_model.py:446-462)."""
    if trim_start_ms > 0:
        n = int(sample_rate * trim_start_ms / 1000)
        if 0 < n < audio.shape[0]:
            audio = audio[n:]
    if fade_in_ms > 0 and audio.shape[0] > 1:
        n = min(max(0, int(sample_rate * fade_in_ms / 1000)), audio.shape[0])
        if n > 1:
            ramp = np.linspace(0.0, 1.0, n).astype(audio.dtype)
            audio = np.concatenate([audio[:n] * ramp, audio[n:]], axis=0)
    return audio
