#!/bin/bash
# usage: tools/sweep_env.sh <tag> "ENV1=a ENV2=b" ["ENV..." ...]   -> gpurun_out/<tag>_<i>.log (bench lines, value only runs)
tag=$1; shift
i=0
for envs in "$@"; do
  env $envs python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-latency --skip-parity > gpurun_out/${tag}_$i.log 2> gpurun_out/${tag}_$i.err
  echo "$i [$envs] rc=$?"
  i=$((i+1))
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_*.log')):
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(l['value']), 'e2e', round(l['e2e']['value']), 'frame', l['branches']['pipelined_frame_us'], l['frame_sections_us'], 'attn', round(l['roofline']['avg_launch_ms']*1e3,1), round(l['roofline']['frac'],3))
    except Exception as e:
        print(f, 'ERR', e)
PY
