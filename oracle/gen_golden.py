"""Generate tests/golden/*.npz by running the REFERENCE's own Python over oracle/mlx_shim.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py

What is executed unmodified from /root/reference: TTSModel.load_model (YAML + safetensors walk +
conv weight transposes), get_state_for_audio_prompt (predefined-voice branch), generate_audio /
generate_audio_stream (sentence splitting, KV sizing, Mimi warm-up, text prefill, frame loop, EOS
rule, trim/fade).  What is substituted: `mlx` (NumPy shim, oracle/mlx_shim), `soundfile` (absent;
only the CLI uses it) and `huggingface_hub.hf_hub_download` (no network: resolves to the synthetic
local bundle).  `TTSModel._run_flow_lm_and_increment_step` is wrapped by a recorder that calls the
original and stores its outputs (latents, EOS flags); noise draws are recorded by the shim's
`mx.random`.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REPO / "oracle" / "mlx_shim"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO))

BUNDLE = Path("/tmp/ptts_golden_bundle")


def _install_stubs():
    sys.modules.setdefault("soundfile", types.ModuleType("soundfile"))
    hub = types.ModuleType("huggingface_hub")

    def hf_hub_download(repo_id, filename, revision=None):
        return str(BUNDLE / filename)

    hub.hf_hub_download = hf_hub_download
    sys.modules["huggingface_hub"] = hub


def _record_calls(model):
    calls = []
    orig = model._run_flow_lm_and_increment_step

    def wrapped(*a, **k):
        out = orig(*a, **k)
        calls.append((np.asarray(out[0]).reshape(-1).copy(), bool(np.asarray(out[1]).reshape(-1)[0])))
        return out

    model._run_flow_lm_and_increment_step = wrapped
    return calls


def run_reference(yaml_path, text, voice, seed, max_frames=None, frames_after_eos=None, max_tokens=50,
                  warmup_frames=1, trim_start_ms=0, fade_in_ms=0, **load_kw):
    import mlx.core as mx
    from pocket_tts_mlx import TTSModel

    model = TTSModel.load_model(str(yaml_path), **load_kw)
    mx.random.seed(seed)
    state = model.get_state_for_audio_prompt(voice)
    voice_draws = [d.copy() for d in mx.random.draws]
    mx.random.draws.clear()
    calls = _record_calls(model)
    tokens = np.asarray(model.flow_lm.conditioner.prepare(text).tokens).reshape(-1)
    chunks = []
    for i, c in enumerate(model.generate_audio_stream(state, text, max_tokens=max_tokens,
                                                      frames_after_eos=frames_after_eos,
                                                      warmup_frames=warmup_frames)):
        chunks.append(np.asarray(c, dtype=np.float32))
        if max_frames is not None and len(chunks) >= max_frames:
            break
    audio = np.concatenate(chunks)
    post = np.asarray(model._postprocess_audio_start(mx.array(audio), trim_start_ms, fade_in_ms))
    draws = np.stack([d.reshape(-1) for d in mx.random.draws])
    lat = np.stack([c[0] for c in calls[1:]])          # calls[0] is the text prefill
    flags = np.array([c[1] for c in calls[1:]], dtype=np.bool_)
    return dict(tokens=tokens.astype(np.int32), voice_noise=np.stack([d.reshape(-1) for d in voice_draws]),
                noise=draws.astype(np.float32), step_latents=lat.astype(np.float32), step_eos=flags,
                audio=audio.astype(np.float32), audio_post=post.astype(np.float32),
                n_frames=np.int32(len(chunks)))


def gen_voice_clone():
    """Golden vector for the voice-cloning branch: the reference's own `_encode_audio` (Mimi encoder, encoder
    transformer, downsample, speaker projection; tts_model.py:271-276) on a synthetic 3-second waveform, and the
    FlowLM state it prompts (frame count).  3 s = 600 encoder steps > the 250-step attention window."""
    _install_stubs()
    import mlx.core as mx
    from pocket_tts_mlx import TTSModel
    from pocket_tts_mlx_b200.synthetic import write_synthetic_bundle

    yml = write_synthetic_bundle(BUNDLE, seed=0)
    out = REPO / "tests" / "golden"
    model = TTSModel.load_model(str(yml))
    assert model.has_voice_cloning
    rng = np.random.Generator(np.random.PCG64(99))
    t = np.arange(71000) / 24000.0                                  # not a multiple of the 1920-sample frame
    audio = (0.3 * np.sin(2 * np.pi * 220.0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
             + 0.05 * rng.standard_normal(t.shape[0])).astype(np.float32)
    # `TTSModel._encode_audio` (tts_model.py:271-276) calls mx.transpose(encoded, (-1, -2)) on a 3-D array, which
    # MLX (and the shim) reject: the upstream voice-cloning branch cannot run as written.  The part that does run
    # unmodified is MimiModel.encode_to_latent; the two remaining lines (swap the last two axes, multiply by
    # speaker_proj_weight^T) are applied here as they were evidently meant.
    try:
        model._encode_audio(mx.array(audio[None, :])[None, ...])
        upstream_ok = True
    except ValueError:
        upstream_ok = False
    encoded = np.asarray(model.mimi.encode_to_latent(mx.array(audio[None, :])[None, ...]))     # [1, 512, T_v]
    latents = np.swapaxes(encoded, -1, -2).astype(np.float32)[0]                               # [T_v, 512]
    cond = latents @ np.asarray(model.flow_lm.speaker_proj_weight).T
    np.savez(out / "ref_voice_clone.npz", audio=audio, encoded=latents, conditioning=cond.astype(np.float32),
             upstream_encode_audio_runs=np.bool_(upstream_ok))
    print("voice clone:", audio.shape[0], "samples ->", cond.shape, "| upstream _encode_audio runs:", upstream_ok)


TEXT_MULTI = "First sentence here. Second one follows! Is this the third? Yes it is."


def gen_multichunk():
    """Full multi-chunk run of the public generate_audio: max_tokens=8 cuts the text into several chunks, every
    frame counts as EOS (threshold -1e30) so each chunk stops after its own `frames_after_eos` guess + 2 frames
    (tts_model.py:346-361, 402-412), the noise stream runs on across chunks, trim + fade at the end."""
    _install_stubs()
    from pocket_tts_mlx_b200.synthetic import write_synthetic_bundle
    yml = write_synthetic_bundle(BUNDLE, seed=0)
    out = REPO / "tests" / "golden"
    g = run_reference(yml, TEXT_MULTI, "cosette", seed=13, eos_threshold=-1e30, max_tokens=8, frames_after_eos=None,
                      trim_start_ms=10, fade_in_ms=25)
    from pocket_tts_mlx.models.tts_model import split_into_best_sentences
    from pocket_tts_mlx import TTSModel
    tok = TTSModel.load_model(str(yml)).flow_lm.conditioner.tokenizer
    g["n_chunks"] = np.int32(len(split_into_best_sentences(tok, TEXT_MULTI, 8)))
    np.savez(out / "ref_multichunk.npz", **g)
    print("multichunk:", int(g["n_chunks"]), "chunks,", int(g["n_frames"]), "frames,", g["noise"].shape[0], "noise draws")


def gen_stream_wav():
    """Bytes the reference's StreamingWAVWriter / stream_audio_chunks (data/audio.py:48-130) produce for seeded float
    chunks, incl. samples beyond [-1, 1] (clipped) and the 0.2 s trailing silence."""
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("ref_audio", str(REF / "pocket_tts_mlx" / "data" / "audio.py"))
    ra = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ra)
    rng = np.random.Generator(np.random.PCG64(5))
    chunks = [(rng.standard_normal(1920) * 0.7).astype(np.float32) for _ in range(4)]

    class Keep(io.BytesIO):
        def close(self):
            self.final = self.getvalue()
            super().close()

    sink = Keep()
    ra.stream_audio_chunks(sink, iter(chunks), 24000)
    np.savez(REPO / "tests" / "golden" / "ref_stream_wav.npz", chunks=np.stack(chunks),
             wav_bytes=np.frombuffer(sink.final, dtype=np.uint8))
    print("stream wav:", len(sink.final), "bytes")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "voice_clone":
        return gen_voice_clone()
    if len(sys.argv) > 1 and sys.argv[1] == "multichunk":
        return gen_multichunk()
    if len(sys.argv) > 1 and sys.argv[1] == "stream_wav":
        return gen_stream_wav()
    _install_stubs()
    from oracle.ptts_oracle import Oracle
    from pocket_tts_mlx_b200.config import load_config
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors
    from pocket_tts_mlx_b200.synthetic import write_synthetic_bundle

    yml = write_synthetic_bundle(BUNDLE, seed=0)
    out = REPO / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    cfg = load_config(yml)
    weights = read_safetensors(cfg.weights_path)
    voice = read_safetensors(BUNDLE / "embeddings" / "alba.safetensors")["audio_prompt"]

    # --- case 1: BASELINE config 1 ("Hello from MLX!", voice alba, max_tokens=200) with a live EOS --
    text = "Hello from MLX!"
    free = run_reference(yml, text, "alba", seed=3, max_tokens=200, eos_threshold=1e30, max_frames=24)
    # choose an EOS threshold with a wide margin from the fp64 oracle's logit trajectory
    orc = Oracle(weights, cfg, dtype=np.float64, eos_threshold=1e30)
    vs = orc.new_flow_state()
    orc.prefill_audio(vs, voice[0], z=free["voice_noise"][0])
    res = orc.generate(vs, free["tokens"], free["noise"], frames_after_eos=5, decode_audio=False, max_frames=24)
    lg = res["eos_logits"]
    best = None
    for k in range(3, 16):
        gap = lg[k] - lg[:k].max()
        if best is None or gap > best[1]:
            best = (k, gap)
    k, gap = best
    assert gap > 0.2, f"no safe EOS threshold (best gap {gap})"
    theta = float((lg[k] + lg[:k].max()) / 2)
    print(f"EOS: first crossing at frame {k}, threshold {theta:.4f}, margin {gap / 2:.3f}")
    hello = run_reference(yml, text, "alba", seed=3, max_tokens=200, eos_threshold=theta,
                          trim_start_ms=20, fade_in_ms=15)
    hello["eos_threshold"] = np.float64(theta)
    hello["eos_first_frame"] = np.int32(k)
    hello["oracle64_eos_logits"] = lg
    np.savez(out / "ref_hello_eos.npz", **hello)
    print("hello:", int(hello["n_frames"]), "frames,", len(hello["tokens"]), "tokens")

    # --- case 2: longer free-running utterance, EOS disabled, 40 frames (Mimi ring wraps at frame 15) --
    text2 = ("The quick brown fox jumps over the lazy dog. "
             "Streaming synthesis keeps state across frames, so every frame matters!")
    long = run_reference(yml, text2, "marius", seed=7, eos_threshold=1e30, max_frames=40, max_tokens=200)
    np.savez(out / "ref_long40.npz", **long)
    print("long:", int(long["n_frames"]), "frames,", len(long["tokens"]), "tokens")

    # --- case 3: non-default sampling knobs (2 LSD steps, noise clamp, temp) + multi-chunk text ---
    text3 = "First sentence here. Second one follows! Is this the third? Yes it is."
    knobs = run_reference(yml, text3, "jean", seed=11, eos_threshold=1e30, max_frames=9, max_tokens=8,
                          warmup_frames=2, temp=0.9, lsd_decode_steps=2, noise_clamp=1.0,
                          frames_after_eos=2)
    np.savez(out / "ref_knobs.npz", **knobs)
    print("knobs:", int(knobs["n_frames"]), "frames,", len(knobs["tokens"]), "tokens")

    # --- text splitting vectors (host logic, tts_model.py:521-593) ---
    from pocket_tts_mlx.models.tts_model import prepare_text_prompt, split_into_best_sentences
    from pocket_tts_mlx import TTSModel
    model = TTSModel.load_model(str(yml))
    tok = model.flow_lm.conditioner.tokenizer
    cases = ["Hello from MLX!", "hello world", text2, text3, "one two three four five six",
             "A. B! C? D... E", "  leading and trailing  \n new line ", "no punctuation at the end of this"]
    import json
    rec = []
    for t in cases:
        for mt in (8, 20, 50):
            rec.append({"text": t, "max_tokens": mt, "chunks": split_into_best_sentences(tok, t, mt),
                        "prepared": list(prepare_text_prompt(t)),
                        "ids": [np.asarray(tok(c).tokens).reshape(-1).tolist()
                                for c in split_into_best_sentences(tok, t, mt)]})
    (out / "ref_text_split.json").write_text(json.dumps(rec, indent=1))
    print("wrote", sorted(p.name for p in out.iterdir()))


if __name__ == "__main__":
    main()
