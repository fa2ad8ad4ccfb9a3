#!/bin/bash
export CTN_PEER_TIMEOUT_S=30
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-250
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 30 --warmup 5 2> gpurun_out/r3_final_n2.err | grep '^{' > gpurun_out/r3_final_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r3_final_n2.json')); print('n2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('exchange','')[:50], d.get('exchange_error'), d['gpu_launches'])" || tail -5 gpurun_out/r3_final_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --config 3 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c3 n2', round(d['value']), round(d['ms_per_step'],3), d.get('exchange','')[:30])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>/dev/null | grep '^{' | cut -c1-160
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('n1', round(d['value']), round(d['ms_per_step'],3))"
