#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02l.log
timeout 300 python tools/z_ab.py 4000 25000 - -@BNMF_ZR=16 -@BNMF_ZR=4 2>&1 | tee gpurun_out/z_ab_r02l.log
timeout 300 python tools/z_ab.py 4000 12500 -@BNMF_ZR=8 -@BNMF_ZR=16 -@BNMF_ZR=8,BNMF_ZR_B=4 2>&1 | tee -a gpurun_out/z_ab_r02l.log
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c3_r02l.err > gpurun_out/bench_c3_r02l.json
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_c3_r02l.json").read()); print("c3", round(j["value"],1), "it/s e2e", round(j["e2e"]["value"],1), j["ms_per_step"], j.get("kernels_ms_per_step"), j["roofline"]["avg_launch_ms"], j["roofline"].get("launch_ms_kernel_alone"), j.get("value_l2_warm"))
PY
