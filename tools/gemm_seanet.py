"""Kernel-level times of the Mimi / SEANet GEMM shapes at batch B (cold = L2 flushed, graph = replayed back to back)."""
import sys
sys.path.insert(0, '/root/repo')
from pathlib import Path
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = load_config(Path('/root/repo/pocket_tts_mlx_b200/config/b6369a24.yaml'))
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
shapes = {
    "mimi.qkv": (B, 16, 1, 512, 1536, 0), "mimi.out": (B, 16, 1, 512, 512, 0),
    "mimi.ff1": (B, 16, 1, 512, 2048, 1), "mimi.ff2": (B, 16, 1, 2048, 512, 0),
    "sn.conv0": (B, 16, 7, 512, 512, 1), "sn.ct0": (B, 16, 2, 512, 1536, 2),
    "sn.r3_0": (B, 96, 3, 256, 128, 1), "sn.r1_0": (B, 96, 1, 128, 256, 5),
    "sn.ct1": (B, 96, 2, 256, 640, 2), "sn.r3_1": (B, 480, 3, 128, 64, 1),
    "sn.r1_1": (B, 480, 1, 64, 128, 5), "sn.ct2": (B, 480, 2, 128, 256, 2),
}
tot_c = tot_g = 0.0
for name, (nb, t, taps, c, n, epi) in shapes.items():
    us, ch = ctx.gemm_bench(nb, t, taps, c, n, epi, reps=7)
    us2, _ = ctx.gemm_bench(nb, t, taps, c, n, epi, reps=-1020)
    tot_c += us; tot_g += us2
    print(f"{name:9s} cold {us:6.1f} us  graph {us2:6.1f} us  plan bn={ch[0]} stages={ch[1]} persist={ch[3] & 1} stg_sets={2 if ch[3] & 2 else 1}")
print(f"sum cold {tot_c:.1f} graph {tot_g:.1f}")
