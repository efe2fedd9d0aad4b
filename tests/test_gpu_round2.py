"""GPU parity tests added in round 2: the BENCHMARKED configuration (batch 256, BASELINE configs 4 and 5 shapes) against
the oracle, non-default sampling knobs in bf16 mode, 16-bit PCM output, voice lifetime, mixed-voice continuous batching,
the multi-chunk public call against the reference's golden run, the sharded launcher.  Run with `pytest -m gpu`.

Tolerances (BASELINE.json north_star): per-frame latents rel-L2 <= 1e-2 in bf16 mode (teacher-forced) and <= 1e-4 in
fp32 mode; waveform SNR >= 30 dB; integer outputs (PCM, frame counts) bit-exact.
"""

import gc
import io

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2, snr_db

pytestmark = pytest.mark.gpu

TEXT_MULTI = "First sentence here. Second one follows! Is this the third? Yes it is."


@pytest.fixture(scope="module")
def model_b256(bundle):
    """bf16 model with a KV pool sized for 256 sequences (the bench configuration)."""
    from pocket_tts_mlx_b200 import TTSModel
    m = TTSModel.load_model(str(bundle), eos_threshold=1e30, precision="bf16", kv_pool_tokens=256 * 384 + 8192)
    yield m
    m.close()


@pytest.fixture(scope="module")
def model_fp32(bundle):
    from pocket_tts_mlx_b200 import TTSModel
    m = TTSModel.load_model(str(bundle), eos_threshold=1e30, precision="fp32", kv_pool_tokens=65536)
    yield m
    m.close()


def _oracle(weights, cfg, cond, **kw):
    from oracle.ptts_oracle import Oracle
    orc = Oracle(weights, cfg, dtype=np.float32, **kw)
    st = orc.new_flow_state()
    orc.prefill_audio(st, cond, z=None)
    return orc, st


def _teacher_forced(batch, noise, refs, frames, pipelined):
    """Step `frames` frames feeding the oracle's latents back for the sequences in `refs`; returns the per-frame
    latents / logits of those sequences and their audio (re-aligned by one frame in pipelined mode)."""
    lat_rec = {b: [] for b in refs}
    log_rec = {b: [] for b in refs}
    aud_rec = {b: [] for b in refs}
    for f in range(frames):
        lat, logit, au = batch.step(noise[1 + f])
        forced = lat.copy()
        for b in refs:
            lat_rec[b].append(lat[b].copy())
            log_rec[b].append(float(logit[b]))
            if not pipelined or f >= 1:
                aud_rec[b].append(au[b].copy())
            forced[b] = refs[b]["latents"][f]
        batch.set_prev_latent(forced)
    if pipelined:
        last = batch.flush()
        for b in refs:
            aud_rec[b].append(last[b].copy())
    return lat_rec, log_rec, aud_rec


# ------------------------------------------------------------------------------------ the benchmarked configuration
def test_config4_batch256_shared_voice_vs_oracle(model_b256, cfg, weights, voices):
    """BASELINE config 4 as bench.py runs it: 256 sequences x 60 tokens, one shared 125-frame voice (cascade attention
    at 4096 CTAs), the two-branch pipelined frame graph; 6 teacher-forced frames; 8 sampled sequences (first, last, the
    128-row tile edges) against the fp32 oracle.  This is where the planner picks the persistent two-accumulator GEMMs,
    the 16-warp epilogue, split-K with the plane sum in the LayerNorm and TMA-staged residuals."""
    from pocket_tts_mlx_b200 import _native
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    n, frames = 256, 6
    ids = list(synthetic_token_ids(2, n, 60))
    rng = np.random.Generator(np.random.PCG64(256))
    noise = rng.standard_normal((1 + frames, n, 32)).astype(np.float32)
    sample = [0, 1, 63, 127, 128, 129, 200, 255]
    orc, st = _oracle(weights, cfg, voices("alba")[0], eos_threshold=1e30)
    refs = {b: orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames) for b in sample}
    state = model_b256.get_state_for_audio_prompt("alba")
    batch = _native.Batch(model_b256._ctx, [state["voice_id"]] * n, [state["prompt_len"] + 60 + frames + 4] * n)
    batch.set_pipelined(True)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    lat, logit, aud = _teacher_forced(batch, noise, refs, frames, pipelined=True)
    batch.close()
    for b in sample:
        for f in range(frames):
            e = rel_l2(lat[b][f], refs[b]["latents"][f])
            assert e < 1e-2, (b, f, e)
            assert abs(logit[b][f] - refs[b]["eos_logits"][f]) < 5e-2, (b, f)
        s = snr_db(np.concatenate(aud[b]), refs[b]["audio"])
        assert s > 30.0, (b, s)


def test_config4_folded_prefix_attention_matches_separate_launch(model_b256):
    """PTTS_FOLD=1 (read when a batch records its step graph): the shared-voice prefix partials are computed inside the
    persistent stream-attention kernel -- (16 rows, head) tiles claimed through an atomic counter, published through
    per-tile flags -- instead of by flow_prefix_attention_kernel.  Same math on the same bf16 keys: the generated
    latents agree with the two-launch path to accumulation-order noise over 6 free-running frames."""
    import os
    from pocket_tts_mlx_b200 import _native
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    n, frames = 256, 6
    ids = list(synthetic_token_ids(5, n, 60))
    rng = np.random.Generator(np.random.PCG64(99))
    noise = rng.standard_normal((frames, n, 32)).astype(np.float32)
    state = model_b256.get_state_for_audio_prompt("alba")
    out = {}
    for fold in ("0", "1"):
        os.environ["PTTS_FOLD"] = fold
        try:
            batch = _native.Batch(model_b256._ctx, [state["voice_id"]] * n, [state["prompt_len"] + 60 + frames + 4] * n)
            batch.set_pipelined(True)
            batch.warmup_mimi(1)
            batch.prefill_text(ids)
            out[fold] = [tuple(a.copy() for a in batch.step(noise[f])) for f in range(frames)]
            batch.close()
        finally:
            os.environ.pop("PTTS_FOLD", None)
    for f in range(frames):
        e = rel_l2(out["1"][f][0], out["0"][f][0])
        assert e < 5e-3, (f, e)
        assert np.max(np.abs(out["1"][f][1] - out["0"][f][1])) < 5e-2, f


def test_config4_batch256_async_staging_is_bit_identical(model_b256):
    """The exact call pattern of the bench's e2e leg (pipelined graph + asynchronous double-buffered staged steps)
    returns bit for bit what synchronous pipelined host steps return at batch 256."""
    from pocket_tts_mlx_b200 import _native
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    n, frames = 256, 5
    ids = list(synthetic_token_ids(3, n, 60))
    rng = np.random.Generator(np.random.PCG64(77))
    noise = rng.standard_normal((frames, n, 32)).astype(np.float32)
    state = model_b256.get_state_for_audio_prompt("alba")
    out = {}
    for mode in ("sync", "async"):
        batch = _native.Batch(model_b256._ctx, [state["voice_id"]] * n, [state["prompt_len"] + 60 + frames + 4] * n)
        batch.set_pipelined(True)
        batch.warmup_mimi(1)
        batch.prefill_text(ids)
        rec = []
        if mode == "sync":
            for f in range(frames):
                rec.append(tuple(a.copy() for a in batch.step(noise[f])))
        else:
            batch.set_async_staging(True)
            sets = batch.staging_sets()
            pending = None
            for f in range(frames):
                sets[f & 1][0][...] = noise[f]
                k = batch.step_staged_async()
                if pending is not None:
                    batch.staged_wait(pending)
                    rec.append(tuple(a.copy() for a in sets[pending][1:]))
                pending = k
            batch.staged_wait(pending)
            rec.append(tuple(a.copy() for a in sets[pending][1:]))
            with pytest.raises(_native.PttsError, match="async staging"):
                batch.step(noise[0])                  # the single-set entry points are refused while it is on
            with pytest.raises(_native.PttsError, match="async staging"):
                batch.step_staged()
        rec.append(batch.flush())
        out[mode] = rec
        batch.close()
    for f in range(frames):
        for a, b in zip(out["sync"][f], out["async"][f]):
            assert np.array_equal(a, b), f
    assert np.array_equal(out["sync"][frames], out["async"][frames])


def test_recycled_arena_forgets_async_staging_graphs(model_b256):
    """A batch that ran with asynchronous staging leaves frame graphs whose odd frames copy through the second staging
    set; the next batch recycled from its arena, stepped with the synchronous call, must not replay them (found by the
    bench's own parity check in round 2)."""
    from pocket_tts_mlx_b200 import _native
    rng = np.random.Generator(np.random.PCG64(606))
    st = model_b256.get_state_for_audio_prompt("alba")
    n, frames = 40, 4
    ids = [rng.integers(0, 4000, size=7).astype(np.int32) for _ in range(n)]
    noise = rng.standard_normal((frames, n, 32)).astype(np.float32)

    def make():
        b = _native.Batch(model_b256._ctx, [st["voice_id"]] * n, [st["prompt_len"] + 7 + frames + 4] * n)
        b.set_pipelined(True)
        return b

    first = make()
    first.set_async_staging(True)
    first.warmup_mimi(1)
    first.prefill_text(ids)
    sets = first.staging_sets()
    for f in range(frames):
        sets[f & 1][0][...] = noise[f]
        first.staged_wait(first.step_staged_async())
    first.close()
    outs = []
    for _ in range(2):                       # the first of these recycles the async arena
        b = make()
        b.warmup_mimi(1)
        b.prefill_text(ids)
        outs.append([tuple(a.copy() for a in b.step(noise[f])) for f in range(frames)] + [(b.flush(),)])
        b.close()
    for x, y in zip(*outs):
        for a, c in zip(x, y):
            assert np.array_equal(a, c)
    assert np.abs(outs[0][1][0]).max() > 0 and np.isfinite(outs[0][1][0]).all()


def test_batch256_mixed_voices_plain_attention_vs_oracle(model_b256, cfg, weights, voices):
    """256 sequences over four different voices: no shared prefix, so the per-sequence attention walks every key
    (the non-cascade path of the bench shape); sequential frame graph."""
    from pocket_tts_mlx_b200 import _native
    from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
    n, frames = 256, 4
    names = ["alba", "marius", "jean", "cosette"]
    ids = list(synthetic_token_ids(5, n, 60))
    rng = np.random.Generator(np.random.PCG64(1256))
    noise = rng.standard_normal((1 + frames, n, 32)).astype(np.float32)
    sample = [0, 1, 127, 128, 254, 255]
    refs = {}
    for b in sample:
        orc, st = _oracle(weights, cfg, voices(names[b % 4])[0], eos_threshold=1e30)
        refs[b] = orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames)
    vstates = [model_b256.get_state_for_audio_prompt(v) for v in names]
    batch = _native.Batch(model_b256._ctx, [vstates[b % 4]["voice_id"] for b in range(n)],
                          [vstates[b % 4]["prompt_len"] + 60 + frames + 2 for b in range(n)])
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    lat, logit, aud = _teacher_forced(batch, noise, refs, frames, pipelined=False)
    batch.close()
    for b in sample:
        for f in range(frames):
            e = rel_l2(lat[b][f], refs[b]["latents"][f])
            assert e < 1e-2, (b, f, e)
        assert snr_db(np.concatenate(aud[b]), refs[b]["audio"]) > 30.0, b


def test_config5_batch256_long_context_vs_oracle(bundle, cfg, weights, voices):
    """BASELINE config 5 length at the bench batch size: an 800-frame voice prompt + 150..174 text tokens puts every
    decode step at KV length ~ 960..980 (31 pages deep, second page-table row of 8 entries and beyond), 256 sequences.
    Three sampled sequences, 4 teacher-forced frames, against the oracle."""
    from pocket_tts_mlx_b200 import TTSModel, _native
    n, frames, v_len = 256, 4, 800
    rng = np.random.Generator(np.random.PCG64(5256))
    base = voices("alba")[0]
    cond = (rng.standard_normal((v_len, base.shape[1])) * base.std()).astype(np.float32)
    n_tok = [int(rng.integers(150, 175)) for _ in range(n)]
    n_tok[0], n_tok[255] = 174, 150
    ids = [rng.integers(0, 4000, size=k).astype(np.int32) for k in n_tok]
    noise = rng.standard_normal((1 + frames, n, 32)).astype(np.float32)
    sample = [0, 128, 255]
    orc, st = _oracle(weights, cfg, cond, eos_threshold=1e30)
    refs = {b: orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames) for b in sample}
    m = TTSModel.load_model(str(bundle), eos_threshold=1e30, precision="bf16", kv_pool_tokens=v_len + 64 + n * 256)
    try:
        state = m.get_state_for_conditioning(cond)
        batch = _native.Batch(m._ctx, [state["voice_id"]] * n, [v_len + k + frames + 2 for k in n_tok])
        batch.set_pipelined(True)
        batch.warmup_mimi(1)
        batch.prefill_text(ids)
        assert batch.lengths().tolist() == [v_len + k for k in n_tok]
        lat, logit, aud = _teacher_forced(batch, noise, refs, frames, pipelined=True)
        batch.close()
    finally:
        m.close()
    for b in sample:
        for f in range(frames):
            e = rel_l2(lat[b][f], refs[b]["latents"][f])
            assert e < 1e-2, (b, f, e)
        assert snr_db(np.concatenate(aud[b]), refs[b]["audio"]) > 30.0, b


def test_bf16_knobs_lsd2_clamp_temp_vs_oracle(bundle, cfg, weights, voices):
    """bf16 tensor-core pipeline with non-default sampling: temp 0.9, two LSD (Euler) steps, noise clamp 1.0 -- the flow
    head then runs its GEMM chain twice per frame with the per-step time-embedding constants."""
    from pocket_tts_mlx_b200 import TTSModel, _native
    n, frames = 20, 5
    rng = np.random.Generator(np.random.PCG64(2020))
    ids = [rng.integers(0, 4000, size=int(rng.integers(10, 24))).astype(np.int32) for _ in range(n)]
    noise = (rng.standard_normal((1 + frames, n, 32)) * 1.5).astype(np.float32)      # the clamp bites
    sample = [0, 7, 19]
    orc, st = _oracle(weights, cfg, voices("jean")[0], eos_threshold=1e30, temp=0.9, lsd_decode_steps=2, noise_clamp=1.0)
    refs = {b: orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames) for b in sample}
    m = TTSModel.load_model(str(bundle), temp=0.9, lsd_decode_steps=2, noise_clamp=1.0, eos_threshold=1e30,
                            precision="bf16", kv_pool_tokens=32768)
    try:
        state = m.get_state_for_audio_prompt("jean")
        for pipelined in (False, True):
            batch = _native.Batch(m._ctx, [state["voice_id"]] * n, [state["prompt_len"] + len(t) + frames + 4 for t in ids])
            batch.set_pipelined(pipelined)
            batch.warmup_mimi(1)
            batch.prefill_text(ids)
            lat, logit, aud = _teacher_forced(batch, noise, refs, frames, pipelined)
            batch.close()
            for b in sample:
                for f in range(frames):
                    e = rel_l2(lat[b][f], refs[b]["latents"][f])
                    assert e < 1e-2, (pipelined, b, f, e)
                assert snr_db(np.concatenate(aud[b]), refs[b]["audio"]) > 30.0, (pipelined, b)
    finally:
        m.close()


# ------------------------------------------------------------------------------------ 16-bit PCM / streaming output
@pytest.mark.parametrize("n_seq", [1, 3, 64])
def test_pcm16_output_is_bit_exact(model_b256, model_fp32, n_seq):
    """with pcm16=True the kernels that produce the final samples also store trunc(clip(v) * 32767); the
    result must equal the reference's host conversion (data/audio.py:70) of the fp32 output of the same run, sample for
    sample: fused tcgen05 SEANet tail + boundary fix-up (3 sequences: fewer tiles than CTAs; 64 sequences: several tiles
    per CTA and the threaded block scatter of the facade) and the CUDA-core output conv (batch 1 in fp32 mode), pipelined
    and sequential."""
    from pocket_tts_mlx_b200.audio import to_pcm16
    model = model_fp32 if n_seq == 1 else model_b256
    rng = np.random.Generator(np.random.PCG64(900 + n_seq))
    st = model.get_state_for_audio_prompt("alba")
    ids = [rng.integers(0, 4000, size=int(rng.integers(4, 9))).astype(np.int32) for _ in range(n_seq)]
    frames = 5
    noise = (rng.standard_normal((1 + frames, n_seq, 32)) * 2.0).astype(np.float32)       # loud: some samples clip
    for pipelined in (False, True):
        f32 = model.generate_audio_batch([st] * n_seq, ids, max_frames=frames, noise=noise, pipelined=pipelined)
        i16 = model.generate_audio_batch([st] * n_seq, ids, max_frames=frames, noise=noise, pipelined=pipelined, pcm16=True)
        for a, b in zip(f32, i16):
            assert b.dtype == np.int16 and b.shape == a.shape == (frames * 1920,)
            assert np.array_equal(b, to_pcm16(a)), (n_seq, pipelined)


def test_streaming_wav_from_gpu_pcm(model_fp32):
    """generate_audio_stream(pcm16=True) -> stream_audio_chunks: the WAV stream written from the GPU's int16 frames is
    byte-identical to the one written from the fp32 frames through the reference's host conversion."""
    from pocket_tts_mlx_b200.audio import stream_audio_chunks

    class Keep(io.BytesIO):
        def close(self):
            self.final = self.getvalue()
            super().close()

    st = model_fp32.get_state_for_audio_prompt("marius")
    rng = np.random.Generator(np.random.PCG64(33))
    noise = rng.standard_normal((64, 32)).astype(np.float32)
    outs = []
    for pcm in (False, True):
        sink = Keep()
        chunks = model_fp32.generate_audio_stream(st, "Hello from MLX!", frames_after_eos=2, noise=noise, pcm16=pcm)
        stream_audio_chunks(sink, chunks, model_fp32.sample_rate)
        outs.append(sink.final)
    assert outs[0] == outs[1] and len(outs[0]) > 44 + 2 * 1920


# ------------------------------------------------------------------------------------ public call, multi-chunk
def test_multichunk_generate_audio_against_reference_golden(bundle):
    """The whole public `generate_audio` over a text that splits into 4 chunks (max_tokens=8), live EOS rule with the
    per-chunk `frames_after_eos` guess, one noise stream across chunks, trim + fade: fp32 mode vs the reference's run."""
    from pocket_tts_mlx_b200 import TTSModel
    g = np.load(GOLDEN / "ref_multichunk.npz")
    m = TTSModel.load_model(str(bundle), eos_threshold=-1e30, precision="fp32", kv_pool_tokens=16384)
    try:
        st = m.get_state_for_audio_prompt("cosette")
        audio = m.generate_audio(st, TEXT_MULTI, max_tokens=8, trim_start_ms=10, fade_in_ms=25, noise=g["noise"])
        assert audio.shape == g["audio_post"].shape                 # frame count of every chunk bit-exact
        assert snr_db(audio, g["audio_post"]) > 60.0
        frames = list(m.generate_audio_stream(st, TEXT_MULTI, max_tokens=8, noise=g["noise"]))
        assert len(frames) == int(g["n_frames"]) and all(f.shape == (1920,) for f in frames)
        assert snr_db(np.concatenate(frames), g["audio"]) > 60.0
    finally:
        m.close()


# ------------------------------------------------------------------------------------ voice lifetime
def test_voice_states_release_their_pages(bundle):
    """A server that fetches a voice per request must not run out of KV pages: the state owns its prefix pages and gives
    them back when it is dropped; a voice that live batch slots still attend is kept until they let go; the 2-entry LRU
    of `_cached_get_state_for_audio_prompt` (reference tts_model.py:478-482) reuses prefills."""
    from pocket_tts_mlx_b200 import TTSModel, _native
    m = TTSModel.load_model(str(bundle), eos_threshold=1e30, precision="bf16", kv_pool_tokens=1024)     # 32 pages
    try:
        for _ in range(40):                              # 4 pages each: 160 pages if they leaked
            st = m.get_state_for_audio_prompt("alba")
            del st
            gc.collect()
        st = m.get_state_for_audio_prompt("alba")
        rng = np.random.Generator(np.random.PCG64(3))
        ids = [rng.integers(0, 4000, size=6).astype(np.int32)]
        noise = rng.standard_normal((5, 1, 32)).astype(np.float32)
        ref = m.generate_audio_batch([st], ids, max_frames=4, noise=noise, return_latents=True, pipelined=False)
        # drop the state while a batch made from it is alive: the slot keeps the prefix pages
        batch = _native.Batch(m._ctx, [st["voice_id"]], [st["prompt_len"] + 6 + 8])
        vid = st["voice_id"]
        del st
        gc.collect()
        with pytest.raises(_native.PttsError):
            m._ctx.voice_length(vid)                     # the id is retired ...
        other = [m.get_state_for_audio_prompt("marius") for _ in range(3)]      # ... and churn cannot take its pages
        batch.warmup_mimi(1)
        batch.prefill_text(ids)
        lats = [batch.step(noise[1 + f])[0][0].copy() for f in range(4)]
        batch.close()
        assert np.array_equal(np.stack(lats), ref[1][0])
        del other
        gc.collect()
        # everything is back: a batch needing 28 of the 32 pages fits (4 stay with the cached voice below)
        a = m._cached_get_state_for_audio_prompt("alba")
        assert m._cached_get_state_for_audio_prompt("alba") is a
        big = _native.Batch(m._ctx, [a["voice_id"]], [1024 - 128 + 29])
        big.close()
        b_ = m._cached_get_state_for_audio_prompt("marius")
        del big
        c_ = m._cached_get_state_for_audio_prompt("jean")            # evicts "alba"
        assert m._cached_get_state_for_audio_prompt("marius") is b_
        assert m._cached_get_state_for_audio_prompt("alba") is not a
        assert c_["prompt_len"] == 125
    finally:
        m.close()


def test_unused_checkpoint_keys_are_reported(bundle, tmp_path):
    """§8f-1: the loader tells which flow_lm.* / mimi.* keys nothing consumed (none for a checkpoint in the expected
    layout), so the Appendix-B key map can be validated the day a real checkpoint is present."""
    from pocket_tts_mlx_b200 import TTSModel, _native
    from pocket_tts_mlx_b200.config import load_config
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors
    m = TTSModel.load_model(str(bundle), eos_threshold=1e30, precision="bf16", kv_pool_tokens=4096)
    try:
        assert m.unused_checkpoint_keys == []
    finally:
        m.close()
    cfg = load_config(bundle)
    ccfg = _native.make_config(cfg, 0.7, 1, None, 1e30, "bf16", 4096, 0)
    ctx = _native.Context(ccfg, 0)
    try:
        for k, a in read_safetensors(cfg.weights_path).items():
            ctx.load_weight(k, a)
        ctx.load_weight("flow_lm.not_a_real_module.weight", np.zeros((3, 3), np.float32))
        assert not ctx.load_weight("optimizer.step", np.zeros(1, np.float32))       # foreign prefix: refused up front
        ctx.finalize()
        assert ctx.unused_weights() == ["flow_lm.not_a_real_module.weight"]
    finally:
        ctx.close()


def test_real_checkpoint_key_map(tmp_path):
    """Opt-in (no network here): POCKET_TTS_REAL_CKPT=<dir with tts_b6369a24.safetensors, tokenizer.model,
    embeddings/> loads the real kyutai checkpoint, requires that every checkpoint key is consumed, and speaks."""
    import os
    root = os.environ.get("POCKET_TTS_REAL_CKPT")
    if not root:
        pytest.skip("set POCKET_TTS_REAL_CKPT to a directory holding the real checkpoint to run this")
    from pathlib import Path
    import yaml
    from pocket_tts_mlx_b200 import TTSModel
    from pocket_tts_mlx_b200.synthetic import default_bundle_dir, write_synthetic_bundle
    y = yaml.safe_load(Path(write_synthetic_bundle(default_bundle_dir(), seed=0)).read_text())
    y["weights_path"] = str(Path(root) / "tts_b6369a24.safetensors")
    y["flow_lm"]["lookup_table"]["tokenizer_path"] = str(Path(root) / "tokenizer.model")
    yml = tmp_path / "real.yaml"
    yml.write_text(yaml.safe_dump(y))
    os.environ["POCKET_TTS_VOICES_DIR"] = str(Path(root) / "embeddings")
    m = TTSModel.load_model(str(yml), precision="bf16")
    try:
        assert m.unused_checkpoint_keys == [], m.unused_checkpoint_keys
        audio = m.generate_audio(m.get_state_for_audio_prompt("alba"), "Hello from the real checkpoint.")
        assert audio.ndim == 1 and audio.shape[0] > 24000 // 2 and np.isfinite(audio).all()
        assert float(np.abs(audio).max()) > 1e-3
    finally:
        m.close()


def test_mimi_decode_chunked_readback_and_reused_buffer(model_fp32, cfg, weights):
    """BASELINE config 3 path (`ptts_batch_mimi_decode`, models/mimi.py:70-75): the waveforms leave the device in chunks
    of 16 frames while the decoder keeps running; 37 frames = two whole chunks and a short one, read into a caller-owned
    array that held garbage, against the oracle frame by frame."""
    from oracle.ptts_oracle import Oracle
    from pocket_tts_mlx_b200 import _native
    n, frames = 3, 37
    rng = np.random.Generator(np.random.PCG64(41))
    lat = rng.standard_normal((n, frames, 32)).astype(np.float32)
    vid = model_fp32.get_state_for_audio_prompt("alba")
    batch = _native.Batch(model_fp32._ctx, [vid["voice_id"]] * n, [vid["prompt_len"] + 8] * n)
    batch.warmup_mimi(1)
    out = np.full((n, frames * 1920), np.nan, dtype=np.float32)
    got = batch.mimi_decode(lat, out=out)
    batch.close()
    assert got is out and np.isfinite(out).all()
    with pytest.raises(ValueError):
        batch2 = _native.Batch(model_fp32._ctx, [vid["voice_id"]] * n, [vid["prompt_len"] + 8] * n)
        try:
            batch2.mimi_decode(lat, out=np.empty((n, 5), np.float32))
        finally:
            batch2.close()
    orc = Oracle(weights, cfg, dtype=np.float32)
    for b in range(n):
        ms = orc.new_mimi_state()
        orc.warmup_mimi(ms, 1)
        ref = np.concatenate([orc.mimi_decode_frame(ms, lat[b, f]) for f in range(frames)])
        assert snr_db(out[b], ref) > 60.0, b


# ------------------------------------------------------------------------------------ continuous batching / sharding
def test_continuous_batching_mixed_voices_with_cascade_start(model_b256):
    """40 jobs through 32 slots where the first 32 share one voice (so the batch starts with cascade
    attention) and the later ones use another: the batch must drop the cascade when a slot switches voice and every
    utterance must still equal its standalone decode."""
    rng = np.random.Generator(np.random.PCG64(4040))
    sa = model_b256.get_state_for_audio_prompt("alba")
    sm = model_b256.get_state_for_audio_prompt("marius")
    n_jobs, frames = 40, 5
    states = [sa] * 32 + [sm, sa, sm, sm, sa, sm, sa, sm]
    ids = [rng.integers(0, 4000, size=int(rng.integers(4, 9))).astype(np.int32) for _ in range(n_jobs)]
    noise = [rng.standard_normal((1 + frames, 32)).astype(np.float32) for _ in range(n_jobs)]
    waves, lats = model_b256.generate_audio_continuous(states, ids, slots=32, noise=noise, max_frames=frames,
                                                       return_latents=True)
    assert all(len(l) == frames for l in lats)
    for j in (0, 31, 32, 33, 35, 39):
        w1, l1 = model_b256.generate_audio_batch([states[j]], [ids[j]], noise=noise[j][:, None, :], max_frames=frames,
                                                 return_latents=True, pipelined=False)
        assert rel_l2(lats[j], l1[0]) < 1e-2, (j, rel_l2(lats[j], l1[0]))
        assert snr_db(waves[j], w1[0]) > 30.0, j


def test_sharded_launcher_equals_single_replica(model_fp32):
    """N replicas == 1 replica on the same utterance set.  Two ranks are played one after the other
    on this GPU (`generate_audio_sharded(rank=r, world_size=2)`); with per-utterance noise every utterance must come
    out as in the single-replica run, whatever rank and slot it landed in."""
    rng = np.random.Generator(np.random.PCG64(88))
    names = ["alba", "marius", "jean"]
    vs = {v: model_fp32.get_state_for_audio_prompt(v) for v in names}
    n_jobs = 9
    n_tok = [3, 8, 5, 4, 9, 6, 3, 7, 5]
    states = [vs[names[j % 3]] for j in range(n_jobs)]
    ids = [rng.integers(0, 4000, size=k).astype(np.int32) for k in n_tok]
    noise = [rng.standard_normal((1 + 70, 32)).astype(np.float32) for _ in range(n_jobs)]
    whole = model_fp32.generate_audio_sharded(states, ids, rank=0, world_size=1, slots=4, noise=noise, return_latents=True)
    assert whole[0] == list(range(n_jobs)) or sorted(whole[0]) == list(range(n_jobs))
    one = dict(zip(whole[0], zip(*whole[1])))
    seen = []
    for r in range(2):
        idx, (waves, lats) = model_fp32.generate_audio_sharded(states, ids, rank=r, world_size=2, slots=2, noise=noise,
                                                               return_latents=True)
        seen += idx
        for j, w, l in zip(idx, waves, lats):
            assert l.shape == one[j][1].shape
            assert rel_l2(l, one[j][1]) < 1e-5, (r, j)
            assert snr_db(w, one[j][0]) > 80.0, (r, j)
    assert sorted(seen) == list(range(n_jobs))


# ------------------------------------------------------------------------------------ cluster chain kernel
def _bf16(x):
    from pocket_tts_mlx_b200.safetensors_io import bf16_bits_to_f32, f32_to_bf16_bits
    x = np.asarray(x, dtype=np.float32)
    return bf16_bits_to_f32(f32_to_bf16_bits(x)).reshape(x.shape).astype(np.float64)


@pytest.mark.parametrize("m", [256, 128, 37, 300])
def test_cluster_chain_kernel_matches_numpy(model_b256, m):
    """csrc/chain_tc.cu on a miniature flow head: six dependent GEMM steps in one cluster launch (multicast A blocks,
    cluster-scope step barriers, LayerNorm with row statistics exchanged through distributed shared memory, AdaLN
    modulation, gated residual, Euler update) against NumPy with bf16 rounding at the same points; M = 256 (two
    clusters), one cluster, ragged tiles."""
    rng = np.random.Generator(np.random.PCG64(1000 + m))
    d, k0 = 512, 1024
    f = lambda *s, sc=1.0: (rng.standard_normal(s) * sc).astype(np.float32)
    t = dict(a0=f(m, k0), a1=np.concatenate([f(m, 32), np.zeros((m, 32), np.float32)], axis=1),
             w0=f(d, k0, sc=k0 ** -0.5), b0=f(d, sc=0.1), wa=f(3 * d, d, sc=d ** -0.5), ba=f(3 * d, sc=0.1),
             wi=f(d, 64, sc=32 ** -0.5), bi=f(d, sc=0.1), lnw=1 + f(d, sc=0.1), lnb=f(d, sc=0.1),
             w1=f(d, d, sc=d ** -0.5), b1=f(d, sc=0.1), w2=f(d, d, sc=d ** -0.5), b2=f(d, sc=0.1),
             wf=f(32, d, sc=d ** -0.5), bf=f(32, sc=0.1), lat_in=f(m, 32))
    out = model_b256._ctx.debug_chain(t, out_scale=0.5)
    assert out["nc"] in (8, 16)
    silu = lambda x: x / (1 + np.exp(-x))

    def ln(x, eps):
        mu = x.mean(-1, keepdims=True)
        return (x - mu) / np.sqrt(x.var(-1, keepdims=True) + eps)

    sy = _bf16(silu(_bf16(t["a0"]) @ _bf16(t["w0"]).T + t["b0"]))
    ada = sy @ _bf16(t["wa"]).T + t["ba"]
    shift, scale, gate = ada[:, :d], ada[:, d:2 * d], ada[:, 2 * d:]
    x1 = _bf16(t["a1"]) @ _bf16(t["wi"]).T + t["bi"]
    h = _bf16((ln(x1, 1e-6) * t["lnw"] + t["lnb"]) * (1 + scale) + shift)
    u = _bf16(silu(h @ _bf16(t["w1"]).T + t["b1"]))
    x1 = x1 + gate * (u @ _bf16(t["w2"]).T + t["b2"])
    h2 = _bf16(ln(x1, 1e-6))
    lat = t["lat_in"] + 0.5 * (h2 @ _bf16(t["wf"]).T + t["bf"])
    # bf16 intermediates (sy, h, u) are rounded from fp32 on the device and from fp64 here: a value that sits on a
    # rounding boundary flips by one bf16 ulp (2^-8 relative), which is what the tolerances below leave room for
    assert rel_l2(out["ada"], ada) < 1e-3, rel_l2(out["ada"], ada)
    assert rel_l2(out["x1"], x1) < 2e-3, rel_l2(out["x1"], x1)
    assert rel_l2(out["h"], h2) < 4e-3, rel_l2(out["h"], h2)
    assert rel_l2(out["lat"], lat) < 2e-3, rel_l2(out["lat"], lat)


@pytest.mark.parametrize("lsd", [1, 2])
def test_single_launch_flow_head_vs_oracle(bundle, cfg, weights, voices, monkeypatch, lsd):
    """PTTS_CHAIN=1: the whole flow head (all Euler steps: cond embedding, AdaLN table, input projection, six
    LayerNorm-modulated residual blocks, final layer, latent update) runs as ONE cluster-chain launch; teacher-forced
    latents and the waveform against the oracle, and against the default launch chain of the same model."""
    from pocket_tts_mlx_b200 import TTSModel, _native
    n, frames = 40, 4
    rng = np.random.Generator(np.random.PCG64(3030 + lsd))
    ids = [rng.integers(0, 4000, size=int(rng.integers(8, 20))).astype(np.int32) for _ in range(n)]
    noise = rng.standard_normal((1 + frames, n, 32)).astype(np.float32)
    sample = [0, 17, 39]
    orc, st = _oracle(weights, cfg, voices("alba")[0], eos_threshold=1e30, lsd_decode_steps=lsd)
    refs = {b: orc.generate(st, ids[b], noise[:, b, :], frames_after_eos=3, max_frames=frames) for b in sample}
    got = {}
    for chain in ("1", "0"):
        monkeypatch.setenv("PTTS_CHAIN", chain)
        m = TTSModel.load_model(str(bundle), lsd_decode_steps=lsd, eos_threshold=1e30, precision="bf16", kv_pool_tokens=32768)
        try:
            state = m.get_state_for_audio_prompt("alba")
            # different reservations per mode: the arena (and its op lists) of the other mode must not be recycled
            batch = _native.Batch(m._ctx, [state["voice_id"]] * n,
                                  [state["prompt_len"] + len(t) + frames + (4 if chain == "1" else 40) for t in ids])
            launches0 = m._ctx.launch_count(reset=True)
            batch.warmup_mimi(1)
            batch.prefill_text(ids)
            lat, logit, aud = _teacher_forced(batch, noise, refs, frames, pipelined=False)
            got[chain] = (lat, m._ctx.launch_count())
            batch.close()
        finally:
            m.close()
        for b in sample:
            for f in range(frames):
                e = rel_l2(lat[b][f], refs[b]["latents"][f])
                assert e < 1e-2, (chain, b, f, e)
            assert snr_db(np.concatenate(aud[b]), refs[b]["audio"]) > 30.0, (chain, b)
    # one launch instead of 23 per Euler step and frame
    assert got["0"][1] - got["1"][1] >= frames * (22 * lsd) - 4, (got["0"][1], got["1"][1])
    for b in sample:
        assert rel_l2(np.stack(got["1"][0][b]), np.stack(got["0"][0][b])) < 6e-3
