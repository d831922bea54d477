#!/bin/bash
for L in "$@"; do
  export BNMF_LIB=$PWD/bayesnmf_b200/$L
  echo "== $L"
  timeout 600 python -m pytest tests/test_gpu_poisson.py -q -x -m gpu 2>&1 | tail -2
  echo -n "c3 wgs: "; python tools/prof_z.py 4000 100000
  echo -n "12.5k: "; python tools/prof_z.py 4000 12500
  echo -n "exome: "; python tools/prof_z.py 100 100000
done
