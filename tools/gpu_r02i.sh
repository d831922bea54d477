#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r02i.log
timeout 500 python tools/z_ab.py 4000 100000,12500 - exp/lib_mnext.so exp/lib_efull.so exp/lib_nomt.so 2>&1 | tee gpurun_out/z_ab_r02i.log
timeout 100 python tools/z_ab.py 100 100000 - exp/lib_mnext.so 2>&1 | tee -a gpurun_out/z_ab_r02i.log
