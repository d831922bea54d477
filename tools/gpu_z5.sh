#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -3
for G in 12500 25000 50000 100000; do echo -n "G=$G default: "; python tools/prof_z.py 4000 $G; done
for G in 12500 100000; do for B in 1 2 4; do for C in 50 100 200; do echo -n "G=$G ZR_B=$B ctB=$C: "; BNMF_ZR_B=$B BNMF_Z_CTB=$C python tools/prof_z.py 4000 $G; done; done; done
echo -n "G=12500 ZR=16: "; BNMF_ZR=16 python tools/prof_z.py 4000 12500
echo -n "G=12500 ZR=4: "; BNMF_ZR=4 python tools/prof_z.py 4000 12500
echo -n "exome: "; python tools/prof_z.py 100 100000
