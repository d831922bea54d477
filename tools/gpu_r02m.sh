#!/bin/bash
mkdir -p gpurun_out
for L in "" exp/lib_nomt.so exp/lib_nomnext.so exp/lib_oldfix.so exp/lib_nomtmn.so; do
BNMF_LIB=$L timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/tmp_bench.json
python - <<PY
import json
j=json.loads(open("gpurun_out/tmp_bench.json").read()); r=j["roofline"]
print("lib=${L:-main}", "flushed", round(j["value"],1), "ms", round(j["ms_per_step"],4), "warm", round(j.get("value_l2_warm",0),1), "z in-step", round(r["avg_launch_ms"],4), "alone", round(r.get("launch_ms_kernel_alone",0),4), "e2e", round(j["e2e"]["value"],1))
PY
done 2>&1 | tee gpurun_out/bench_ab_r02m.log
