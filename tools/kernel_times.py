"""Eager per-kernel CUDA-event times of one frame (python tools/kernel_times.py B [filter])."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flt = sys.argv[2] if len(sys.argv) > 2 else ""
start = 137
model, _ = load_model(0, B * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, B, 60 + start))
batch = _native.Batch(model._ctx, [state["voice_id"]] * B, [state["prompt_len"] + 60 + start + 64] * B)
batch.warmup_mimi(1)
batch.prefill_text(ids)
for _ in range(3):
    batch.step_device()
batch.profile_step()
rows = batch.profile_step()
tot = sum(r["ms"] for r in rows)
for r in rows:
    if flt in r["kernel"]:
        print(f"{r['kernel']:28s} n={r['launches']:3d} {1e3*r['ms']:8.1f} us {1e3*r['ms']/r['launches']:7.1f} us/launch "
              f"{r['bytes']/r['ms']/1e6 if r['ms'] else 0:8.0f} GB/s(model)")
print("total", round(1e3 * tot, 1), "us (eager, event overhead included)")
batch.close()
