#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --workload c3-exome --precision f32 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c3-exome_f32_r02s.err > gpurun_out/bench_c3-exome_f32_r02s.json
python -c "
import json; j=json.load(open('gpurun_out/bench_c3-exome_f32_r02s.json')); r=j['roofline']; print('c3-exome f32', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), 'warm', round(j['value_l2_warm'],1), 'roof', r['kernel'], r['frac'], r.get('frac_kernel_alone'), 'z', r['avg_launch_ms'], r.get('launch_ms_kernel_alone'), j['kernels_ms_per_step'])"
