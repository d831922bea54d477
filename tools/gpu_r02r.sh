#!/bin/bash
# last call of round 2: HEAD as the driver runs it (GPU suite, smoke, the default bench line)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02r.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c3_driver_r02r.err > gpurun_out/bench_c3_driver_r02r.json
python -c "
import json; j=json.load(open('gpurun_out/bench_c3_driver_r02r.json')); r=j['roofline']; print('c3 (driver flags)', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), 'warm', round(j['value_l2_warm'],1), 'issue', r['issue']['frac'], r['issue']['thread_insts_per_pick'], 'z', r['avg_launch_ms'], r['launch_ms_kernel_alone'])"
