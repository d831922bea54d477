#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/z_ab.py 4000 12500,100000 - exp/lib_cg.so exp/lib_cg10.so exp/lib_cg20.so exp/lib_u2.so 2>&1 | tee gpurun_out/z_ab_r02k.log
