timeout 200 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_fullsize.py tests/test_gpu_e2e.py tests/test_gpu_init_prior.py -m gpu -q -x 2>&1 | tail -6
timeout 120 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value'],1), round(j['e2e']['value'],1), j['kernels_ms_per_step'])"
