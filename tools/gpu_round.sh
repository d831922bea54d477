#!/bin/bash
# One gpurun call: GPU tests, smoke, the bench lines, then the ncu launch list and one full
# capture of the latent-count kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps 100 --warmup 5 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json
timeout 600 python bench.py --steps 100 --warmup 5 --workload c3-exome --no-cpu-baseline 2>> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_exome_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 300 $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zstat -s 4 -c 2 -f -o gpurun_out/prof_zstat_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
