"""Voice-cloning encode time: GPU (ptts_encode_audio) vs the NumPy oracle, 10-second prompt."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from oracle.ptts_oracle import Oracle
from pocket_tts_mlx_b200.safetensors_io import read_safetensors

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
model, _ = load_model(0, 8192)
rng = np.random.Generator(np.random.PCG64(3))
t = np.arange(int(24000 * secs)) / 24000.0
audio = (0.3 * np.sin(2 * np.pi * 200 * t) + 0.05 * rng.standard_normal(t.shape[0])).astype(np.float32)
model.encode_audio(audio)                                  # warm-up (module load, allocations)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); cond = model.encode_audio(audio); ts.append(time.perf_counter() - t0)
print(f"GPU encode of {secs:.0f} s of audio -> {cond.shape}: median {1e3 * np.median(ts):.1f} ms (host call, incl. H2D/D2H and temp allocations)")
t0 = time.perf_counter(); st = model.get_state_for_conditioning(cond); t1 = time.perf_counter()
print(f"prompt prefill of {cond.shape[0]} frames: {1e3 * (t1 - t0):.1f} ms")
orc = Oracle(read_safetensors(model.config.weights_path), model.config, dtype=np.float32)
t0 = time.perf_counter(); ref = orc.encode_audio(audio); t1 = time.perf_counter()
print(f"NumPy oracle encode: {t1 - t0:.2f} s; rel-L2 GPU vs oracle {np.linalg.norm(cond - ref) / np.linalg.norm(ref):.2e}")
