"""Does the GPU overlap two half-batches? Two contexts (two streams) x B/2 sequences vs one context x B."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import load_model
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.synthetic import synthetic_token_ids

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 120
start = 137


def make(n):
    model, _ = load_model(0, n * 700 + 4096)
    state = model.get_state_for_audio_prompt("alba")
    ids = list(synthetic_token_ids(2, n, 60 + start))
    batch = _native.Batch(model._ctx, [state["voice_id"]] * n, [state["prompt_len"] + 60 + start + 3 * frames + 16] * n)
    batch.warmup_mimi(1)
    batch.prefill_text(ids)
    for _ in range(3):
        batch.step_device()
    model._ctx.sync()
    return model, batch


def run(pairs, label):
    for m, _ in pairs: m._ctx.sync()
    t0 = time.perf_counter()
    for _ in range(frames):
        for _, b in pairs:
            b.step_device()
    for m, _ in pairs: m._ctx.sync()
    dt = time.perf_counter() - t0
    n = sum(b.B for _, b in pairs) if hasattr(pairs[0][1], "B") else None
    print(f"{label}: {1e6 * dt / frames:.1f} us per frame of all sequences")


one = make(B)
run([one], f"1 ctx x {B}")
one[1].close()
two = [make(B // 2), make(B // 2)]
run(two, f"2 ctx x {B // 2}")
run(two[:1], f"1 ctx x {B // 2}")
