"""Batch-1 per-frame latency (BASELINE config 2) under the current environment: python tools/lat_probe.py [frames]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import latency_bs1, load_model

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
model, _ = load_model(0, 16384)
state = model.get_state_for_audio_prompt("alba")
rng = np.random.Generator(np.random.PCG64(1))
print(latency_bs1(model, state, rng, frames=frames))
model.close()
