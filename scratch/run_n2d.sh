#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q 2>&1 | tail -3 | cut -c1-300
for g in 1 2 3 6; do
CTN_DP_GRAPHS=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('dp_graphs=$g n2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
