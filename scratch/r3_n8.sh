#!/bin/bash
# 8 GPUs: the peer-memory exchange at world 8 (kernel test first, bounded), then the bench lines peer / nccl
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m pytest tests/test_data_parallel_gpu.py -x -q -m gpu -k "every_gpu" 2>&1 | tail -4 | cut -c1-250
for ex in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 30 --warmup 5 --exchange $ex 2> gpurun_out/r3_bench_c1_n8_$ex.err | grep '^{' > gpurun_out/r3_bench_c1_n8_$ex.json
  python -c "
import json; d=json.load(open('gpurun_out/r3_bench_c1_n8_$ex.json')); print('$ex', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('exchange','')[:40], d.get('exchange_error'))" || tail -5 gpurun_out/r3_bench_c1_n8_$ex.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 30 --warmup 5 --exchange peer 2> gpurun_out/r3_bench_c1_n4_peer.err | grep '^{' > gpurun_out/r3_bench_c1_n4_peer.json
python -c "
import json; d=json.load(open('gpurun_out/r3_bench_c1_n4_peer.json')); print('n4 peer', round(d['value']), round(d['ms_per_step'],3))"
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('n1', round(d['value']), round(d['ms_per_step'],3))"
