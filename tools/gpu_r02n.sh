#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -3
timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02n.log
for S in 1 0; do
BNMF_SIDES=$S timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/tmp_bench.json
python - <<PY
import json
j=json.loads(open("gpurun_out/tmp_bench.json").read()); r=j["roofline"]
print("sides=$S", "flushed", round(j["value"],1), "ms", round(j["ms_per_step"],4), "warm", round(j.get("value_l2_warm",0),1), "z in-step", round(r["avg_launch_ms"],4), "alone", round(r.get("launch_ms_kernel_alone",0),4), "e2e", round(j["e2e"]["value"],1), "launches", j["gpu_launches"])
PY
done 2>&1 | tee gpurun_out/bench_ab_r02n.log
