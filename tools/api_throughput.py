"""Throughput of the public batch API (TTSModel.generate_audio_batch) on BASELINE config 4."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
model, _ = load_model(0, 256 * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, 256, 60))
for pipelined in (True, False, True):
    t0 = time.perf_counter()
    waves = model.generate_audio_batch([state] * 256, ids, seed=3, pipelined=pipelined)
    dt = time.perf_counter() - t0
    sec = sum(len(w) for w in waves) / 24000.0
    print(f"generate_audio_batch pipelined={pipelined}: {sec:.0f} audio-s in {dt:.3f} s -> {sec / dt:.0f} audio-s/s")
