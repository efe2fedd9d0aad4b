"""How long does the decode attention take INSIDE the frame graph?  PTTS_ATTN_DBG=1 python tools/attn_in_graph.py
Runs a batch-256 job in sequential and in pipelined mode and prints, per FlowLM layer, the span from the first CTA's
start to the last CTA's end of the attention launch of the last frame (globaltimer)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids

B, start = 256, 137
model, _ = load_model(0, B * 700 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, B, 60 + start))
for pipelined in (False, True):
    batch = _native.Batch(model._ctx, [state["voice_id"]] * B, [state["prompt_len"] + 60 + start + 40] * B)
    batch.set_pipelined(pipelined)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(8):
        batch.step_device()
    model._ctx.sync()
    model._ctx.attention_stamps()          # reset
    batch.step_device()
    model._ctx.sync()
    st = model._ctx.attention_stamps()
    spans = [(int(e) - int(s)) / 1e3 for s, e in st]
    gaps = [(int(st[i + 1][0]) - int(st[i][1])) / 1e3 for i in range(len(st) - 1)]
    print("pipelined" if pipelined else "sequential", "attention span per layer (us):", [round(x, 1) for x in spans],
          "| gap to the next layer's attention (us):", [round(x, 1) for x in gaps])
    batch.close()
model.close()
