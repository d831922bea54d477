#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_gpu_e2e.py -q -x -m gpu 2>&1 | tail -2
for W in c1 c3; do
timeout 200 python bench.py --workload $W --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$W', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['ms_per_step'], j['kernels_ms_per_step'])"
done
for S in 1 2 4 8; do echo -n "c1 split=$S: "; BNMF_Z_SPLIT=$S timeout 100 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value'],1), j['kernels_ms_per_step']['k_zstat'])"; done
