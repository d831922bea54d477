#!/bin/bash
mkdir -p gpurun_out
(timeout 200 python tools/shard_ab.py 12500 300; timeout 200 python tools/shard_ab.py 25000 200; timeout 300 python tools/shard_ab.py 100000 60) 2>&1 | tee gpurun_out/shard_ab_r02o.log
