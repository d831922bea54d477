#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02f}
run() { # N, workload, extra env
  env $3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 50 --warmup 5 --no-cpu-baseline --workload $2 2> gpurun_out/bench_$2_n$1_$TAG.err > gpurun_out/bench_$2_n$1_$TAG.json
  python - <<PY
import json; j=json.load(open("gpurun_out/bench_$2_n$1_$TAG.json")); print("$2 N=$1:", round(j["value"],1), "it/s  e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],4), j["kernels_ms_per_step"], "z alone", j["roofline"].get("launch_ms_kernel_alone"))
PY
}
run 8 c3 A=1
run 4 c3 A=1
