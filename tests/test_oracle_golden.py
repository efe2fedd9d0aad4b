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
