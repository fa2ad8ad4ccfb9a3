#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index --format=csv,noheader | wc -l
run() { # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 "$@" > gpurun_out/n8_$name.out 2> gpurun_out/n8_$name.err
  grep '^{' gpurun_out/n8_$name.out | tail -1 > gpurun_out/r2_bench_${name}_n8.json
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${name}_n8.json')); print('$name n8', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['config']['parallelism'])" || tail -3 gpurun_out/n8_$name.err
}
run c1 --steps 20 --warmup 5
run c4 --config 4 --steps 10 --warmup 3
run c3 --config 3 --steps 10 --warmup 3
run c1fwd --config 1 --mode fwd --steps 20 --warmup 5
