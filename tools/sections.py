"""In-graph section timing for a batch (python tools/sections.py B [kv_start])."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
start = int(sys.argv[2]) if len(sys.argv) > 2 else 137
model, _ = load_model(0, B * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, B, 60 + start))
batch = _native.Batch(model._ctx, [state["voice_id"]] * B, [state["prompt_len"] + 60 + start + 64] * B)
batch.warmup_mimi(1)
batch.prefill_text(ids)
for _ in range(3):
    batch.step_device()
model._ctx.sync()
sec = batch.profile_sections()
print({k: round(v * 1000, 1) for k, v in sec.items()}, "us; parts sum", round(1000 * sum(list(sec.values())[:4]), 1))
batch.close()
