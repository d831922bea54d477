#!/bin/bash
# round 2, last 1-GPU call: whole GPU suite, smoke, every workload, reference arm, ncu launch list, ncu --set full of k_zstat
mkdir -p gpurun_out
TAG=r02q
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
show() { python -c "
import json; j=json.load(open('gpurun_out/bench_$1_$TAG.json')); r=j['roofline']; print('$1', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), 'warm', round(j.get('value_l2_warm',0),1), 'roof', r['kernel'], r['frac'], (r.get('issue') or r.get('fp64') or {}).get('frac'), j.get('mh_phase',{}).get('value'), j.get('kernels_ms_per_step'))"; }
timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 --cpu-budget 10 2> gpurun_out/bench_c3_$TAG.err > gpurun_out/bench_c3_$TAG.json; show c3
for W in c3-exome c1 c2 c4 c5; do
  timeout 300 python bench.py --workload $W --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_${W}_$TAG.err > gpurun_out/bench_${W}_$TAG.json; show $W
done
timeout 300 python bench.py --workload c3 --precision f32 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c3_f32_$TAG.err > gpurun_out/bench_c3_f32_$TAG.json; show c3_f32
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2> gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.json; cut -c1-200 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
tail -1 gpurun_out/ncu_launch_$TAG.log
timeout 300 python tools/prof_z.py 4000 100000 > gpurun_out/plain_z_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zstat -s 3 -c 1 -f -o gpurun_out/prof_zstat_$TAG python tools/prof_z.py 4000 100000 > gpurun_out/ncu_z_$TAG.log 2>&1
tail -2 gpurun_out/ncu_z_$TAG.log
