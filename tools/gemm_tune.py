"""Sweep (N tile, ring stages, split-K, persistent) for every GEMM shape of the batch-256 frame + prefill and
print the measured ranking (kernel-level, L2 flushed).  Results feed the table in gemm_tc.cu."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config

cfg = load_config(Path(__file__).resolve().parents[1] / "pocket_tts_mlx_b200" / "config" / "b6369a24.yaml")
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = {  # name: (nb, T, taps, C, N, epi, max_splits)
    "flow.qkv": (1, B, 1, 1024, 3072, 0, 1), "flow.out": (1, B, 1, 1024, 1024, 0, 8),
    "flow.ff1": (1, B, 1, 1024, 4096, 1, 1), "flow.ff2": (1, B, 1, 4096, 1024, 0, 8),
    "head.ada": (1, B, 1, 512, 10240, 0, 1), "head.m1": (1, B, 1, 512, 512, 1, 1),
    "mimi.qkv": (B, 16, 1, 512, 1536, 0, 1), "mimi.out": (B, 16, 1, 512, 512, 0, 1),
    "mimi.ff1": (B, 16, 1, 512, 2048, 1, 1), "mimi.ff2": (B, 16, 1, 2048, 512, 0, 1),
    "sn.conv0": (B, 16, 7, 512, 512, 1, 1), "sn.ct0": (B, 16, 2, 512, 1536, 2, 1),
    "sn.r3_0": (B, 96, 3, 256, 128, 1, 1), "sn.r1_0": (B, 96, 1, 128, 256, 5, 1),
    "sn.ct1": (B, 96, 2, 256, 640, 2, 1), "sn.r3_1": (B, 480, 3, 128, 64, 1, 1),
    "sn.r1_1": (B, 480, 1, 64, 128, 5, 1), "sn.ct2": (B, 480, 2, 128, 256, 2, 1),
    "sn.r3_2": (B, 1920, 3, 64, 32, 1, 1), "sn.r1_2": (B, 1920, 1, 32, 64, 5, 1),
    "prefill.qkv": (1, 15360, 1, 1024, 3072, 0, 1), "prefill.ff2": (1, 15360, 1, 4096, 1024, 0, 1),
}
out = {}
for name, (nb, t, taps, c, n, epi, ms) in shapes.items():
    res = []
    base, chosen = ctx.gemm_bench(nb, t, taps, c, n, epi)
    for bn in (128, 64, 32):
        if n % bn:
            continue
        for st in (2, 4, 6, 8):
            if c % 64 and st != 2:
                continue
            for sp in (1, 2, 4, 8):
                if sp > ms:
                    continue
                for pe in (0, 1):
                    try:
                        us, ch = ctx.gemm_bench(nb, t, taps, c, n, epi, force=(bn, st, sp, pe), reps=3)
                    except _native.PttsError:
                        continue
                    res.append((us, bn, st, sp, pe))
    res.sort()
    flops = 2.0 * nb * t * n * taps * c
    out[name] = {"planner": [base, chosen], "best": res[:4]}
    print(f"{name:12s} planner {base:7.1f} us {chosen}  best " + "  ".join(f"{u:6.1f}:{b},{s},{k},{p}" for u, b, s, k, p in res[:4])
          + f"   [{flops / res[0][0] / 1e6:.0f} TF/s best]", flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/gemm_tune_b{B}.json").write_text(json.dumps(out, indent=1))
