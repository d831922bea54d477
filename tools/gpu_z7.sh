#!/bin/bash
for cfg in "16 2 400" "16 2 800" "8 2 200" "8 2 400" "16 4 400" "16 4 800" "8 4 800" "32 2 200"; do set -- $cfg; echo -n "G=100000 ZR=$1 ZR_B=$2 ctB=$3: "; BNMF_ZR=$1 BNMF_ZR_B=$2 BNMF_Z_CTB=$3 python tools/prof_z.py 4000 100000; done
for cfg in "4 2 100" "4 2 195" "8 2 195" "4 1 100" "2 2 0" "4 4 0"; do set -- $cfg; echo -n "G=12500 ZR=$1 ZR_B=$2 ctB=$3: "; BNMF_ZR=$1 BNMF_ZR_B=$2 BNMF_Z_CTB=$3 python tools/prof_z.py 4000 12500; done
