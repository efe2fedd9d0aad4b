"""cProfile of TTSModel.generate_audio_batch on BASELINE config 4 (second call): python tools/api_profile.py"""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
model, _ = load_model(0, 256 * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, 256, 60))
model.generate_audio_batch([state] * 256, ids, seed=3)
t0 = time.perf_counter()
w = model.generate_audio_batch([state] * 256, ids, seed=3)
print("plain second call: %.3f s" % (time.perf_counter() - t0))
del w
pr = cProfile.Profile()
pr.enable()
waves = model.generate_audio_batch([state] * 256, ids, seed=3)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
