"""Continuous batching vs lock-step waves on ragged utterance lengths
(python tools/continuous_probe.py [n_jobs] [slots] [min_admit,min_admit,...])."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 768
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model, _ = load_model(0, slots * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
rng = np.random.Generator(np.random.PCG64(5))
n_tok = rng.integers(15, 91, size=n_jobs)                    # 88 .. 400 frames per utterance
ids = [rng.integers(0, 4000, size=int(k)).astype(np.int32) for k in n_tok]
frames = [model._estimate_max_gen_len(int(k)) for k in n_tok]
audio_s = 0.08 * sum(frames)
print(f"{n_jobs} utterances, {min(frames)}..{max(frames)} frames, {audio_s:.0f} audio-seconds in total")

admits = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [None]
for ma in admits:
    for rep in range(2):
        t0 = time.perf_counter()
        waves = model.generate_audio_continuous([state] * n_jobs, ids, slots=slots, seed=1, min_admit=ma)
        dt = time.perf_counter() - t0
        assert [len(w) // 1920 for w in waves] == frames
        print(f"continuous, {slots} slots, min_admit {ma}: {dt:.2f} s -> {audio_s / dt:.0f} audio-s/s")
if len(sys.argv) > 3:
    sys.exit(0)

for pipelined in (True, True, False):
    t0 = time.perf_counter()
    done = 0
    for w0 in range(0, n_jobs, slots):
        sel = list(range(w0, min(n_jobs, w0 + slots)))
        out = model.generate_audio_batch([state] * len(sel), [ids[j] for j in sel], seed=1, pipelined=pipelined)
        done += sum(len(w) // 1920 for w in out)
    dt = time.perf_counter() - t0
    assert done == sum(frames)
    print(f"lock-step waves of {slots} (arrival order, pipelined={pipelined}): {dt:.2f} s -> {audio_s / dt:.0f} audio-s/s")
