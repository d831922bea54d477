#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py -q -x -m gpu 2>&1 | tail -2
for G in 12500 100000; do echo -n "G=$G default: "; timeout 60 python tools/prof_z.py 4000 $G; done
echo -n "exome: "; timeout 60 python tools/prof_z.py 100 100000
echo -n "mu=500: "; timeout 60 python tools/prof_z.py 500 100000
