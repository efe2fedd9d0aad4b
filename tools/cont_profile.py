"""Where the continuous scheduler's host time goes (python tools/cont_profile.py [n_jobs] [slots]): wall time inside the
native calls of one generate_audio_continuous run."""
import sys, time, collections
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model, _ = load_model(0, slots * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
rng = np.random.Generator(np.random.PCG64(5))
n_tok = rng.integers(15, 91, size=n_jobs)
ids = [rng.integers(0, 4000, size=int(k)).astype(np.int32) for k in n_tok]
acc, cnt = collections.Counter(), collections.Counter()
for name in ("reset_seqs", "prefill_text", "staged_wait", "step_staged_async", "set_active", "flush"):
    orig = getattr(_native.Batch, name)
    def wrap(self, *a, _o=orig, _n=name, **k):
        t0 = time.perf_counter()
        try:
            return _o(self, *a, **k)
        finally:
            acc[_n] += time.perf_counter() - t0
            cnt[_n] += 1
    setattr(_native.Batch, name, wrap)
for rep in range(2):
    acc.clear(); cnt.clear()
    t0 = time.perf_counter()
    waves = model.generate_audio_continuous([state] * n_jobs, ids, slots=slots, seed=1)
    dt = time.perf_counter() - t0
    audio = sum(len(w) for w in waves) / 24000
    print(f"rep {rep}: {dt:.3f} s, {audio / dt:.0f} audio-s/s; " + ", ".join(f"{k} {acc[k]*1e3:.0f} ms / {cnt[k]}" for k in acc))
model.close()
