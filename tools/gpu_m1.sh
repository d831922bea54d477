#!/bin/bash
# 2 GPUs: sharded == whole with both exchange paths, bench at N=2 both ways
mkdir -p gpurun_out
BNMF_TRACE=1 timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_replicas.py -q -x -m gpu -rs 2>&1 | grep -v "bnmf_create\|k_p_rows" | tail -12
for X in 1 0; do
BNMF_XCHG=$X timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_n2_x$X.err > gpurun_out/bench_n2_x$X.json
python - <<PY
import json; j=json.load(open("gpurun_out/bench_n2_x$X.json")); print("XCHG=$X N=2:", round(j["value"],1), "it/s e2e", round(j["e2e"]["value"],1), j["kernels_ms_per_step"])
PY
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_n1.err > gpurun_out/bench_n1.json
python -c "
import json; j=json.load(open('gpurun_out/bench_n1.json')); print('N=1:', round(j['value'],1), 'it/s e2e', round(j['e2e']['value'],1), j['kernels_ms_per_step'], j['roofline']['launch_ms_kernel_alone'])"
