#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/z_ab.py 4000 12500 - -@BNMF_ZR_B=1 -@BNMF_ZR_B=1,BNMF_Z_CTB=100 -@BNMF_Z_CTB=100 -@BNMF_Z_CTB=300 -@BNMF_ZR=8 -@BNMF_ZR=8,BNMF_ZR_B=1 -@BNMF_ZR=2 -@BNMF_ZR_B=3 2>&1 | tee gpurun_out/z_ab_r02j.log
timeout 600 python tools/z_ab.py 4000 100000 - -@BNMF_ZR_B=1 -@BNMF_ZR_B=1,BNMF_Z_CTB=100 -@BNMF_Z_CTB=100 -@BNMF_Z_CTB=400 -@BNMF_ZR=8 2>&1 | tee -a gpurun_out/z_ab_r02j.log
