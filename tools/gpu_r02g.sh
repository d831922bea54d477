#!/bin/bash
# round 2, call g: item prologue with every load in flight + k_begin_iter folded into k_pside, A/B against the old prologue
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02g.log
for L in "" exp/lib_old.so; do
  for G in 100000 12500; do
    echo "lib=${L:-new} G=$G"; BNMF_LIB=$L timeout 120 python tools/prof_z.py 4000 $G 2>&1 | tail -1
  done
done
echo "lib=new exome"; timeout 120 python tools/prof_z.py 100 100000 2>&1 | tail -1
for F in 1 0; do
BNMF_FOLD_BEGIN=$F timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_c3_r02g_$F.err > gpurun_out/bench_c3_r02g_$F.json
python - <<PY
import json
j=json.loads(open("gpurun_out/bench_c3_r02g_$F.json").read()); print("c3 fold=$F", round(j["value"],1), "it/s e2e", round(j["e2e"]["value"],1), j["ms_per_step"], j.get("kernels_ms_per_step"))
PY
done
