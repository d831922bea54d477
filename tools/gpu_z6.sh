#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -q -x -m gpu 2>&1 | tail -3
for G in 12500 25000 50000 100000; do echo -n "G=$G default: "; python tools/prof_z.py 4000 $G; done
for ZR in 2 4 8; do echo -n "G=12500 ZR=$ZR: "; BNMF_ZR=$ZR python tools/prof_z.py 4000 12500; done
echo -n "exome: "; python tools/prof_z.py 100 100000
echo -n "f32: "; python tools/prof_z.py 4000 100000 f32
