#!/bin/bash
mkdir -p gpurun_out
for G in 100000 12500; do
  timeout 300 python tools/prof_z.py 4000 $G > gpurun_out/plain_z_$G.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zstat -s 3 -c 1 -f -o gpurun_out/prof_zstat_r02_$G python tools/prof_z.py 4000 $G > gpurun_out/ncu_z_$G.log 2>&1
  tail -2 gpurun_out/ncu_z_$G.log
done
