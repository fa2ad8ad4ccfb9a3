#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -3
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q 2>&1 | tail -6 | cut -c1-300
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench_c1.out 2> gpurun_out/n2_bench_c1.err
grep '^{' gpurun_out/n2_bench_c1.out > gpurun_out/r2_bench_c1_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c1_n2.json')); print('c1 n2', d['value'], d['ms_per_step'], d['e2e']['value'])"
grep -c "NCCL INFO" gpurun_out/n2_bench_c1.out gpurun_out/n2_bench_c1.err | head; grep -h "nranks" gpurun_out/n2_bench_c1.out gpurun_out/n2_bench_c1.err | head -3 | cut -c1-200
