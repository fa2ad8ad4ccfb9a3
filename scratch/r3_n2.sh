#!/bin/bash
# 2 GPUs: the peer-memory exchange against NCCL — tests, then the N = 2 bench lines of both
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q -m gpu 2>&1 | tail -5 | cut -c1-250
for ex in peer nccl peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --exchange $ex > gpurun_out/r3_bench_c1_n2_$ex.json 2> gpurun_out/r3_bench_c1_n2_$ex.err
  python -c "
import json; d=json.load(open('gpurun_out/r3_bench_c1_n2_$ex.json')); print('$ex', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('exchange','')[:40], d.get('exchange_error'))" || tail -5 gpurun_out/r3_bench_c1_n2_$ex.err
done
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('n1', round(d['value']), round(d['ms_per_step'],3))"
