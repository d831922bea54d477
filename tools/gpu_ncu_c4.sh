#!/bin/bash
mkdir -p gpurun_out
BNMF_GRAPH=0 timeout 300 python tools/prof_c5.py c4 > gpurun_out/plain_c4.log 2>&1 &&
BNMF_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_e_gram -s 1 -c 1 -f -o gpurun_out/prof_egram python tools/prof_c5.py c4 > gpurun_out/ncu_egram.log 2>&1
tail -2 gpurun_out/ncu_egram.log
