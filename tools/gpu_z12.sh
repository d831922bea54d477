#!/bin/bash
export BNMF_LIB=$PWD/bayesnmf_b200/libv_r112.so
for HT in 128 256; do
export BNMF_HYPER_THREADS=$HT
timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('regs112 HT=$HT c3', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['ms_per_step'], j['roofline']['avg_launch_ms'], j['roofline']['launch_ms_kernel_alone'])"
done
unset BNMF_LIB
export BNMF_HYPER_THREADS=128
timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('regs96 HT=128 c3', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['ms_per_step'], j['roofline']['avg_launch_ms'], j['roofline']['launch_ms_kernel_alone'])"
