"""Pin the NumPy oracle to golden vectors produced by the reference's own Python
(oracle/gen_golden.py ran /root/reference over oracle/mlx_shim).  CPU only."""

import json

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2, snr_db
from oracle.ptts_oracle import Oracle, postprocess_audio_start


def _voice_state(orc, voices, name, z):
    st = orc.new_flow_state()
    orc.prefill_audio(st, voices(name)[0], z=z)
    return st


def test_hello_eos_pipeline(cfg, weights, voices):
    """BASELINE config 1: 'Hello from MLX!', voice alba, live EOS: frame count, flags, latents, audio."""
    g = np.load(GOLDEN / "ref_hello_eos.npz")
    orc = Oracle(weights, cfg, dtype=np.float32, eos_threshold=float(g["eos_threshold"]))
    st = _voice_state(orc, voices, "alba", g["voice_noise"][0])
    # "Hello from MLX!" has 3 words -> frames_after_eos guess 3 (+2)
    res = orc.generate(st, g["tokens"], g["noise"], frames_after_eos=5)
    assert res["n_frames"] == int(g["n_frames"]) == int(g["eos_first_frame"]) + 5
    n_calls = len(g["step_eos"])
    flags = res["eos_logits"] > float(g["eos_threshold"])
    assert flags.tolist() == g["step_eos"].tolist()
    assert len(res["eos_logits"]) == n_calls
    for f in range(res["n_frames"]):
        assert rel_l2(res["latents"][f], g["step_latents"][f]) < 1e-4
    assert snr_db(res["audio"], g["audio"]) > 80.0
    post = postprocess_audio_start(res["audio"], 24000, trim_start_ms=20, fade_in_ms=15)
    assert post.shape == g["audio_post"].shape
    assert snr_db(post, g["audio_post"]) > 80.0


def test_long40_free_running(cfg, weights, voices):
    """40 free-running frames (Mimi ring buffer wraps at frame 15), EOS disabled."""
    g = np.load(GOLDEN / "ref_long40.npz")
    orc = Oracle(weights, cfg, dtype=np.float32, eos_threshold=1e30)
    st = _voice_state(orc, voices, "marius", g["voice_noise"][0])
    res = orc.generate(st, g["tokens"], g["noise"], frames_after_eos=3, max_frames=40)
    assert res["n_frames"] == 40
    worst = max(rel_l2(res["latents"][f], g["step_latents"][f]) for f in range(40))
    assert worst < 1e-4, worst
    for f in (0, 14, 15, 16, 39):
        a = res["audio"][f * 1920:(f + 1) * 1920]
        assert snr_db(a, g["audio"][f * 1920:(f + 1) * 1920]) > 80.0


def test_knobs_two_lsd_steps_clamp_temp(cfg, weights, voices, bundle):
    """Non-default sampling knobs: temp 0.9, 2 LSD steps, noise clamp 1.0, warmup_frames 2."""
    from pathlib import Path
    from pocket_tts_mlx_b200.text import SentencePieceTokenizer, split_into_best_sentences
    g = np.load(GOLDEN / "ref_knobs.npz")
    tok = SentencePieceTokenizer(4000, Path(bundle).parent / "tokenizer.model")
    text3 = "First sentence here. Second one follows! Is this the third? Yes it is."
    chunk0 = split_into_best_sentences(tok, text3, 8)[0]
    ids = tok.encode(chunk0)
    orc = Oracle(weights, cfg, dtype=np.float32, eos_threshold=1e30, temp=0.9, lsd_decode_steps=2,
                 noise_clamp=1.0)
    st = _voice_state(orc, voices, "jean", g["voice_noise"][0])
    res = orc.generate(st, ids, g["noise"], frames_after_eos=2, warmup_frames=2, max_frames=9)
    assert res["n_frames"] == 9
    worst = max(rel_l2(res["latents"][f], g["step_latents"][f]) for f in range(9))
    assert worst < 1e-4, worst
    assert snr_db(res["audio"], g["audio"]) > 80.0


def test_fp64_oracle_bounds_fp32_rounding(cfg, weights, voices):
    """The fp64 variant bounds the fp32 oracle's own rounding (teacher-forced on golden latents)."""
    g = np.load(GOLDEN / "ref_hello_eos.npz")
    out = {}
    for dt in (np.float32, np.float64):
        orc = Oracle(weights, cfg, dtype=dt, eos_threshold=1e30)
        st = _voice_state(orc, voices, "alba", g["voice_noise"][0])
        out[dt] = orc.generate(st, g["tokens"], g["noise"], frames_after_eos=5, max_frames=6,
                               teacher_latents=g["step_latents"])
    assert rel_l2(out[np.float32]["latents"], out[np.float64]["latents"]) < 1e-5
    assert snr_db(out[np.float32]["audio"], out[np.float64]["audio"]) > 90.0
    assert np.allclose(out[np.float64]["eos_logits"][:6], g["oracle64_eos_logits"][:6], atol=1e-3)


def test_text_split_matches_reference(bundle):
    from pathlib import Path
    from pocket_tts_mlx_b200.text import SentencePieceTokenizer, prepare_text_prompt, split_into_best_sentences
    tok = SentencePieceTokenizer(4000, Path(bundle).parent / "tokenizer.model")
    cases = json.loads((GOLDEN / "ref_text_split.json").read_text())
    assert len(cases) >= 20
    for c in cases:
        chunks = split_into_best_sentences(tok, c["text"], c["max_tokens"])
        assert chunks == c["chunks"], c["text"]
        assert list(prepare_text_prompt(c["text"])) == c["prepared"]
        assert [tok.encode(ch).tolist() for ch in chunks] == c["ids"]
    with pytest.raises(ValueError, match="Text prompt cannot be empty"):
        prepare_text_prompt("   ")


def test_voice_clone_encoder_matches_reference(cfg, weights):
    """Voice-cloning branch: the oracle's Mimi encoder (SEANet encoder, encoder transformer with its 250-step causal
    window, replicate-padded stride-16 downsample, speaker projection) against the reference's own
    MimiModel.encode_to_latent run over the NumPy MLX shim on a 71 000-sample waveform (600 encoder steps, 37
    frames).  The golden file also records that upstream `TTSModel._encode_audio` itself raises (it transposes a
    3-D array with two axes), so the last two lines are pinned by intent, not by execution."""
    g = np.load(GOLDEN / "ref_voice_clone.npz")
    assert not bool(g["upstream_encode_audio_runs"])
    orc = Oracle(weights, cfg, dtype=np.float32)
    cond = orc.encode_audio(g["audio"])
    assert cond.shape == g["conditioning"].shape == (37, 1024)
    assert rel_l2(cond, g["conditioning"]) < 1e-5


def test_multichunk_full_run(cfg, weights, voices, bundle):
    """Whole `generate_audio` of a 4-chunk text (max_tokens=8): every frame is an EOS (threshold -1e30) so each chunk
    stops `guess + 2` frames in (tts_model.py:346-361,402-412), one noise stream runs across the chunks (prefill draw,
    one draw per frame incl. the one whose frame is dropped), trim + fade at the end.  Golden = the reference's run."""
    from pathlib import Path
    from pocket_tts_mlx_b200.text import SentencePieceTokenizer, prepare_text_prompt, split_into_best_sentences
    g = np.load(GOLDEN / "ref_multichunk.npz")
    tok = SentencePieceTokenizer(4000, Path(bundle).parent / "tokenizer.model")
    text = "First sentence here. Second one follows! Is this the third? Yes it is."
    chunks = split_into_best_sentences(tok, text, 8)
    assert len(chunks) == int(g["n_chunks"]) == 4
    orc = Oracle(weights, cfg, dtype=np.float32, eos_threshold=-1e30)
    voice = _voice_state(orc, voices, "cosette", g["voice_noise"][0])
    cursor, audio, lats = 0, [], []
    for ch in chunks:
        _, guess = prepare_text_prompt(ch)
        res = orc.generate(voice, tok.encode(ch), g["noise"][cursor:], frames_after_eos=guess + 2)
        assert res["n_frames"] == guess + 2
        cursor += 1 + len(res["eos_logits"])             # prefill draw + one per flow call (the last one's frame is dropped)
        audio.append(res["audio"])
        lats.append(res["latents"])
    assert cursor == g["noise"].shape[0]
    assert sum(len(l) for l in lats) == int(g["n_frames"])
    audio = np.concatenate(audio)
    assert snr_db(audio, g["audio"]) > 80.0
    post = postprocess_audio_start(audio, 24000, trim_start_ms=10, fade_in_ms=25)
    assert post.shape == g["audio_post"].shape and snr_db(post, g["audio_post"]) > 80.0


def test_streaming_wav_writer_matches_reference_bytes():
    """`stream_audio_chunks` / `StreamingWAVWriter` (placeholder-length header, int16 = clip * 32767
    truncated, 0.2 s trailing silence) byte for byte against the stream the reference's data/audio.py wrote for the same
    chunks; int16 chunks (what the GPU produces with pcm16=True) pass through unchanged."""
    import io
    from pocket_tts_mlx_b200.audio import StreamingWAVWriter, stream_audio_chunks, to_pcm16
    g = np.load(GOLDEN / "ref_stream_wav.npz")

    class Keep(io.BytesIO):
        def close(self):
            self.final = self.getvalue()
            super().close()

    ref = g["wav_bytes"].tobytes()
    sink = Keep()
    stream_audio_chunks(sink, iter(g["chunks"]), 24000)
    assert sink.final == ref
    sink = Keep()
    stream_audio_chunks(sink, (to_pcm16(c) for c in g["chunks"]), 24000)
    assert sink.final == ref
    assert (np.abs(g["chunks"]) > 1).any()               # the fixture does exercise the clip
    pcm = to_pcm16(np.array([-2.0, -1.0, -0.5, 0.0, 0.99999, 1.0, 3.0], dtype=np.float32))
    assert pcm.tolist() == [-32767, -32767, -16383, 0, 32766, 32767, 32767]
    stream_audio_chunks(None, iter(g["chunks"]), 24000)   # path None only drains the generator
    w = StreamingWAVWriter(Keep(), 24000)
    w.write_header()
    assert w.wave_writer.getnframes() == 0 and w.first_chunk_buffer == []


def test_shim_primitives_against_torch():
    """Second, independent pin of the MLX restatement that generated the goldens: the shim's conv1d / conv_transpose1d
    (incl. stride, groups / depthwise), LayerNorm, gelu, softmax, ELU and SiLU against torch's CPU implementations on
    the same tensors, with MLX's documented layouts (NLC activations, [C_out, K, C_in/groups] weights)."""
    import sys
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    sys.path.insert(0, str(GOLDEN.parents[1] / "oracle" / "mlx_shim"))
    import mlx.core as mx
    import mlx.nn as nn
    rng = np.random.Generator(np.random.PCG64(17))
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))

    def conv_cases():
        yield dict(n=2, l=37, cin=8, cout=12, k=7, stride=1, groups=1)
        yield dict(n=1, l=64, cin=16, cout=16, k=3, stride=1, groups=1)
        yield dict(n=2, l=96, cin=6, cout=10, k=8, stride=4, groups=1)       # SEANet encoder downsampling shape
        yield dict(n=1, l=50, cin=12, cout=12, k=32, stride=16, groups=12)   # depthwise (Mimi resample)
        yield dict(n=1, l=40, cin=8, cout=4, k=5, stride=2, groups=2)

    for c in conv_cases():
        x = rng.standard_normal((c["n"], c["l"], c["cin"]))
        w = rng.standard_normal((c["cout"], c["k"], c["cin"] // c["groups"]))
        y = np.asarray(mx.conv1d(mx.array(x), mx.array(w), stride=c["stride"], groups=c["groups"]))
        ref = F.conv1d(t(x).permute(0, 2, 1), t(w).permute(0, 2, 1), stride=c["stride"], groups=c["groups"]).permute(0, 2, 1)
        assert rel_l2(y, ref.numpy()) < 1e-6, ("conv1d", c)
        # transposed: MLX weight [C_out, K, C_in/g]; torch wants [C_in, C_out/g, K]
        xt = rng.standard_normal((c["n"], c["l"], c["cout"]))                 # input channels = conv's outputs
        g_ = c["groups"]
        og, ig = c["cin"] // g_, c["cout"] // g_                               # per-group outputs / inputs of the transpose
        wt = rng.standard_normal((c["cin"], c["k"], ig))                       # [C_out_t, K, C_in_t/g]
        yt = np.asarray(mx.conv_transpose1d(mx.array(xt), mx.array(wt), stride=c["stride"], groups=g_))
        w_pt = torch.zeros(c["cout"], og, c["k"], dtype=torch.float64)
        for gi in range(g_):
            blk = t(wt[gi * og:(gi + 1) * og])                                 # [og, K, ig]
            w_pt[gi * ig:(gi + 1) * ig] = blk.permute(2, 0, 1)                 # [ig, og, K]
        ref_t = F.conv_transpose1d(t(xt).permute(0, 2, 1), w_pt, stride=c["stride"], groups=g_).permute(0, 2, 1)
        assert yt.shape == tuple(ref_t.shape) and rel_l2(yt, ref_t.numpy()) < 1e-6, ("conv_transpose1d", c)

    x = rng.standard_normal((5, 9, 48)).astype(np.float32)
    ln = nn.LayerNorm(48, eps=1e-5)
    ln.weight = mx.array(rng.standard_normal(48).astype(np.float32))
    ln.bias = mx.array(rng.standard_normal(48).astype(np.float32))
    ref = F.layer_norm(torch.from_numpy(x), (48,), torch.from_numpy(np.asarray(ln.weight)), torch.from_numpy(np.asarray(ln.bias)), 1e-5)
    assert rel_l2(np.asarray(ln(mx.array(x))), ref.numpy()) < 1e-6
    ln0 = nn.LayerNorm(48, eps=1e-6, affine=False)
    assert rel_l2(np.asarray(ln0(mx.array(x))), F.layer_norm(torch.from_numpy(x), (48,), None, None, 1e-6).numpy()) < 1e-6
    xs = (rng.standard_normal((7, 33)) * 3).astype(np.float32)
    tx = torch.from_numpy(xs)
    assert rel_l2(np.asarray(nn.gelu(mx.array(xs))), F.gelu(tx).numpy()) < 1e-6             # exact (erf) form
    assert rel_l2(np.asarray(mx.softmax(mx.array(xs), axis=-1)), F.softmax(tx, dim=-1).numpy()) < 1e-6
    assert rel_l2(np.asarray(nn.ELU()(mx.array(xs))), F.elu(tx).numpy()) < 1e-6
    assert rel_l2(np.asarray(nn.SiLU()(mx.array(xs))), F.silu(tx).numpy()) < 1e-6
