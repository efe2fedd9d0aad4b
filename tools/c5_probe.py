"""Config-5 utterances through generate_audio_sharded on one GPU, several times in one process (python tools/c5_probe.py [n] [reps])."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model, workload5, VOICE_FRAMES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, _ = load_model(0, 256 * (VOICE_FRAMES + 174 + 750 + 40) + 4096)
state = model.get_state_for_audio_prompt("alba")
workload5(model, state, 1, 0, n_total=256, n_tok=174, max_frames=24)
for r in range(reps):
    audio, dt, _ = workload5(model, state, 1, 0, n_total=n, n_tok=174)
    print(f"rep {r}: {audio:.0f} audio-s in {dt:.2f} s -> {audio / dt:.0f} audio-s/s", flush=True)
model.close()
