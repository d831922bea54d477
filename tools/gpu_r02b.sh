#!/bin/bash
# round 2, call b (1 GPU): new tests, every workload through bench.py, the reference arm
mkdir -p gpurun_out
TAG=r02b
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
for W in c3 c1 c2 c4 c5 c3-exome; do
  timeout 900 python bench.py --workload $W --steps 20 --warmup 5 2> gpurun_out/bench_${W}_$TAG.err > gpurun_out/bench_${W}_$TAG.json
  tail -c 1500 gpurun_out/bench_${W}_$TAG.json; echo; tail -2 gpurun_out/bench_${W}_$TAG.err
done
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.json; cat gpurun_out/bench_ref_$TAG.json
