"""Where does a batch-256 job spend its time? (setup / prefill / frames via graph, device vs host I/O)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
model, _ = load_model(0, B * 600 + 4096)
state = model.get_state_for_audio_prompt("alba")
ids = list(synthetic_token_ids(2, B, 60))
ctx = model._ctx
for rep in range(2):
    t0 = time.perf_counter()
    batch = _native.Batch(ctx, [state["voice_id"]] * B, [state["prompt_len"] + 60 + 2 * frames + 16] * B)
    ctx.sync(); t1 = time.perf_counter()
    batch.warmup_mimi(1); ctx.sync(); t2 = time.perf_counter()
    batch.prefill_text(ids); ctx.sync(); t3 = time.perf_counter()
    batch.step_device(); ctx.sync(); t4 = time.perf_counter()       # graph capture + first replay
    ctx.timer_begin()
    for _ in range(frames):
        batch.step_device()
    ms_dev = ctx.timer_end(); t5 = time.perf_counter()
    rng = np.random.Generator(np.random.PCG64(0))
    z = rng.standard_normal((B, 32), dtype=np.float32)
    batch.step(z); t6 = time.perf_counter()
    for _ in range(frames - 1):
        batch.step(z)
    t7 = time.perf_counter()
    batch.close(); t8 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f} ms, warmup {1e3*(t2-t1):.1f}, prefill {1e3*(t3-t2):.1f}, first step(capture) "
          f"{1e3*(t4-t3):.1f}, {frames} device steps: {ms_dev:.1f} ms by events ({ms_dev/frames:.3f}/frame), wall {1e3*(t5-t4):.1f}; "
          f"first host step {1e3*(t6-t5):.1f}, {frames-1} host steps {1e3*(t7-t6):.1f} ({1e3*(t7-t6)/(frames-1):.3f}/frame), close {1e3*(t8-t7):.1f}")
