#!/bin/bash
# bench.py under several values of one tuning knob: tools/sweep_env.sh VAR v1 v2 ...   (dev tool)
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_$v.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/sweep_$v.json').read())
print('$VAR=$v', round(d['value'],1), 'it/s  k_zstat live', round(d['roofline']['avg_launch_ms'],4), 'ms')"
done
