"""Flow-head GEMM shapes at batch B under forced tile configurations (graph-replay per-launch us)."""
import sys
sys.path.insert(0, '/root/repo')
from pathlib import Path
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = load_config(Path('/root/repo/pocket_tts_mlx_b200/config/b6369a24.yaml'))
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
shapes = {"head.ada": (1, B, 1, 512, 10240, 0), "head.cond": (1, B, 1, 1024, 512, 1), "head.m1": (1, B, 1, 512, 512, 1),
          "head.m2": (1, B, 1, 512, 512, 0), "flow.qkv": (1, B, 1, 1024, 3072, 0), "flow.ff1": (1, B, 1, 1024, 4096, 1)}
for name, (nb, t, taps, c, n, epi) in shapes.items():
    for force in (None, (32, 8, 1, 0), (64, 8, 1, 0), (64, 8, 1, 1), (128, 6, 1, 0), (128, 6, 1, 1)):
        try:
            us, ch = ctx.gemm_bench(nb, t, taps, c, n, epi, force=force, reps=-1020)
            usc, _ = ctx.gemm_bench(nb, t, taps, c, n, epi, force=force, reps=5)
            print(f"{name:9s} force={force} graph {us:6.2f} us cold {usc:6.1f} chosen={ch}")
        except Exception as e:
            print(name, force, "n/a")
