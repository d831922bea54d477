#!/bin/bash
for G in 12500 25000 50000; do for ZR in 2 4 8 16; do echo -n "G=$G ZR=$ZR: "; BNMF_ZR=$ZR python tools/prof_z.py 4000 $G; done; done
