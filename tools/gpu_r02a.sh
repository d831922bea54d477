#!/bin/bash
# round 2, call a (2 GPUs): the whole GPU suite with nothing skipped, the timings of every BASELINE config,
# ncu launch lists of the sweep configurations
mkdir -p gpurun_out
TAG=r02a
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -40 > gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python tools/time_configs.py > gpurun_out/time_configs_$TAG.log 2>&1
cat gpurun_out/time_configs_$TAG.log
for W in c4 c5; do
  BNMF_GRAPH=0 timeout 300 python tools/prof_c5.py $W > gpurun_out/plain_$W_$TAG.log 2>&1 &&
  BNMF_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${W}_$TAG.csv python tools/prof_c5.py $W > gpurun_out/ncu_${W}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${W}_$TAG.log
done
