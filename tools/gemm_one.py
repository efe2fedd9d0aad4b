"""One GEMM shape through the kernel-level benchmark (python tools/gemm_one.py nb,T,taps,C,N,epi ...) - ncu target."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config
cfg = load_config(Path(__file__).resolve().parents[1] / "pocket_tts_mlx_b200" / "config" / "b6369a24.yaml")
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
for a in sys.argv[1:]:
    nb, t, taps, c, n, epi = (int(x) for x in a.split(","))
    us, ch = ctx.gemm_bench(nb, t, taps, c, n, epi, reps=3)
    print((nb, t, taps, c, n, epi), round(us, 1), "us cold", ch)
