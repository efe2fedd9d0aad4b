"""Run the Mimi decoder for B sequences x F frames (python tools/tail_probe.py B F) - exercises seanet_tail.cu."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
B = int(sys.argv[1]); F = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model, _ = load_model(0, B * 200 + 4096)
state = model.get_state_for_audio_prompt("alba")
batch = _native.Batch(model._ctx, [state["voice_id"]] * B, [state["prompt_len"] + 16] * B)
batch.warmup_mimi(1)
rng = np.random.Generator(np.random.PCG64(0))
a = batch.mimi_decode(rng.standard_normal((B, F, 32)).astype(np.float32))
print("B", B, "ok", a.shape, float(np.abs(a).mean()))
batch.close()
