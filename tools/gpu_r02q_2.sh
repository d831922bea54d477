#!/bin/bash
# round 2, final 2-GPU call: the whole GPU suite with nothing skipped, the sharded bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_r02q_2.txt
timeout 500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -6 > gpurun_out/pytest_gpu_2gpus_r02q.log; cat gpurun_out/gpus_r02q_2.txt >> gpurun_out/pytest_gpu_2gpus_r02q.log; tail -4 gpurun_out/pytest_gpu_2gpus_r02q.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_c3_n2_r02q.err > gpurun_out/bench_c3_n2_r02q.json
python -c "
import json; j=json.load(open('gpurun_out/bench_c3_n2_r02q.json')); print('c3 N=2', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['ms_per_step'])"
