#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "fused or wgrad or separator or paper or adam_state" 2>&1 | tail -8
for v in "fused:" "nofuse:CTN_NO_APPLY_FUSION=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r8_$tag.json 2> gpurun_out/r8_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r8_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['launch_us'])" || tail -5 gpurun_out/r8_$tag.err
done
CTN_TIMING=1 timeout 300 python scratch/insitu_timing.py 2>&1 | grep -E "ctn timing|gemm|wgrad|gln|dwconv|norm_bwd"
