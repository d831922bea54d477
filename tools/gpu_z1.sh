#!/bin/bash
mkdir -p gpurun_out
echo "per_sm=2:"; python tools/prof_z.py 4000 100000
echo "per_sm=1:"; BNMF_Z_PER_SM=1 python tools/prof_z.py 4000 100000
echo "shard 12.5k per_sm=2:"; python tools/prof_z.py 4000 12500
echo "exome:"; python tools/prof_z.py 100 100000
echo "f32:"; python tools/prof_z.py 4000 100000 f32
