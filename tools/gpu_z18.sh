#!/bin/bash
timeout 300 python -m pytest tests -q -x -m gpu 2>&1 | tail -2
for W in c2 c4 c1; do
timeout 200 python bench.py --workload $W --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$W', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['ms_per_step'])"
done
