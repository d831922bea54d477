#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c4_tc.err > gpurun_out/bench_c4_tc.json
python -c "
import json; j=json.load(open('gpurun_out/bench_c4_tc.json')); print('c4 TC:', round(j['value'],1), 'it/s', j['kernels_ms_per_step'])"
BNMF_TC=0 timeout 600 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c4_notc.err > gpurun_out/bench_c4_notc.json
python -c "
import json; j=json.load(open('gpurun_out/bench_c4_notc.json')); print('c4 fp64:', round(j['value'],1), 'it/s', j['kernels_ms_per_step'])"
