#!/bin/bash
timeout 250 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_e2e.py tests/test_gpu_map.py -q -x -m gpu 2>&1 | tail -3
timeout 120 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('c2', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), 'mh', j['mh_phase']['value'], j['kernels_ms_per_step'])"
