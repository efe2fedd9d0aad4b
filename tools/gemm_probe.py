"""Run a few tcgen05 GEMM shapes through ptts_debug_linear (used under ncu)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config

cfg = load_config(Path(__file__).resolve().parents[1] / "pocket_tts_mlx_b200" / "config" / "b6369a24.yaml")
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
shapes = [(64, 480, 1, 64, 128), (64, 480, 2, 128, 256), (256, 16, 1, 2048, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
rng = np.random.Generator(np.random.PCG64(0))
for nb, t, taps, c, n in shapes:
    a = rng.standard_normal((nb, t + taps - 1, c)).astype(np.float32)
    w = (rng.standard_normal((n, taps * c)) / np.sqrt(taps * c)).astype(np.float32)
    for _ in range(2):
        y = ctx.debug_linear(a, w, np.zeros(n, np.float32), taps=taps, path=3)
    print((nb, t, taps, c, n), float(np.abs(y).mean()))
