#!/bin/bash
# round 2, final 4-GPU call: the sharded bench at N = 4
mkdir -p gpurun_out
TAG=r02q
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu-baseline --workload c3 2> gpurun_out/bench_c3_n4_$TAG.err > gpurun_out/bench_c3_n4_$TAG.json
python - <<PY
import json
j=json.load(open("gpurun_out/bench_c3_n4_$TAG.json")); print("c3 N=4:", round(j["value"],1), "it/s  e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],4), "warm", round(j.get("value_l2_warm",0),1), "z alone", j["roofline"].get("launch_ms_kernel_alone"), "z in-step", j["roofline"].get("avg_launch_ms"))
PY
