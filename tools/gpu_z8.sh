#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -3
for G in 12500 25000 50000 100000; do echo -n "G=$G default: "; timeout 60 python tools/prof_z.py 4000 $G; done
for I in 4 6 12 16; do echo -n "G=12500 items/warp=$I: "; BNMF_Z_ITEMS=$I timeout 60 python tools/prof_z.py 4000 12500; echo -n "G=100000 items/warp=$I: "; BNMF_Z_ITEMS=$I timeout 60 python tools/prof_z.py 4000 100000; done
echo -n "exome: "; timeout 60 python tools/prof_z.py 100 100000
