#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02c}
run() { # N, extra env, label, workload
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 50 --warmup 5 --no-cpu-baseline --workload $4 2> gpurun_out/bench_$4_n$1_$3_$TAG.err > gpurun_out/bench_$4_n$1_$3_$TAG.json
  python - <<PY
import json; j=json.load(open("gpurun_out/bench_$4_n$1_$3_$TAG.json")); print("$4 N=$1 $3:", round(j["value"],1), "it/s  e2e", round(j["e2e"]["value"],1), "ms/step", round(j["ms_per_step"],4), j["kernels_ms_per_step"], "z alone", j["roofline"].get("launch_ms_kernel_alone"))
PY
}
run 8 BNMF_XCHG=1 xchg c3
run 8 BNMF_XCHG=0 nccl c3
run 4 BNMF_XCHG=1 xchg c3
run 8 BNMF_XCHG=1 chains c5
