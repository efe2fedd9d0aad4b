"""Batch-N job with a handful of frames, for ncu launch lists (python tools/frame_probe.py B frames)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 6
start = int(sys.argv[3]) if len(sys.argv) > 3 else 137      # KV length offset: emulate the middle of a 275-frame job
model, _ = load_model(0, B * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, B, 60 + start))            # longer text stands in for already generated frames
batch = _native.Batch(model._ctx, [state["voice_id"]] * B, [state["prompt_len"] + 60 + start + frames + 8] * B)
batch.warmup_mimi(1)
batch.prefill_text(ids)
for _ in range(frames):
    batch.step_device()
model._ctx.sync()
print("lengths", batch.lengths()[:4], "launches", model._ctx.launch_count())
batch.close()
