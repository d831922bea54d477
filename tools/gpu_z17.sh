#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_fullsize.py tests/test_gpu_e2e.py tests/test_gpu_map.py tests/test_gpu_replicas.py -q -x -m gpu 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for F in 1 0; do echo -n "c1 fork=$F: "; BNMF_FORK_HYPER=$F timeout 100 python bench.py --workload c1 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['ms_per_step'])"; done
