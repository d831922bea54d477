#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -3
for G in 12500 100000; do for ZR in 1 2 4 8 16; do echo -n "G=$G ZR=$ZR: "; BNMF_ZR=$ZR python tools/prof_z.py 4000 $G; done; done
echo -n "G=12500 default: "; python tools/prof_z.py 4000 12500
echo -n "G=25000 default: "; python tools/prof_z.py 4000 25000
echo -n "G=50000 default: "; python tools/prof_z.py 4000 50000
echo -n "exome: "; python tools/prof_z.py 100 100000
