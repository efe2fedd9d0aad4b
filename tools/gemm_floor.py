import sys
sys.path.insert(0, '/root/repo')
from pathlib import Path
from pocket_tts_mlx_b200 import _native
from pocket_tts_mlx_b200.config import load_config
cfg = load_config(Path('/root/repo/pocket_tts_mlx_b200/config/b6369a24.yaml'))
ctx = _native.Context(_native.make_config(cfg, 0.7, 1, None, -4.0, "bf16", 4096))
for (nb,t,taps,c,n,epi) in [(1,256,1,64,3072,0),(1,256,1,256,3072,0),(1,256,1,1024,3072,0),(1,256,1,64,512,1),(1,256,1,512,512,1),(1,256,1,4096,1024,0),(1,256,1,1024,1024,0)]:
    for force in (None,(32,8,1,0)):
        try:
            us, ch = ctx.gemm_bench(nb,t,taps,c,n,epi,force=force,reps=9)
            us2, _ = ctx.gemm_bench(nb,t,taps,c,n,epi,force=force,reps=-200)
            us3, _ = ctx.gemm_bench(nb,t,taps,c,n,epi,force=force,reps=-1200)
            print((nb,t,taps,c,n,epi), force, 'cold', round(us,1), 'stream back-to-back', round(us2,2), 'graph', round(us3,2), ch)
        except Exception as e:
            pass
