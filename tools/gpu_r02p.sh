#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_poisson.py -q -x -m gpu 2>&1 | tail -2
(timeout 200 python tools/shard_ab.py 12500 300
 BNMF_HYPER_THREADS=128 timeout 200 python tools/shard_ab.py 12500 300 | head -1
 BNMF_HYPER_THREADS=512 timeout 200 python tools/shard_ab.py 12500 300 | head -1
 timeout 300 python tools/shard_ab.py 100000 60) 2>&1 | tee gpurun_out/shard_ab_r02p.log
