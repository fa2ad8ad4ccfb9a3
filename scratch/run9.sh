#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "dwconv or paper or golden or bn or batchnorm" 2>&1 | tail -8
for v in "staged:" "staged_nofuse:CTN_NO_APPLY_FUSION=1" "unstaged:CTN_DW_UNSTAGED=1 CTN_NO_APPLY_FUSION=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r9_$tag.json 2> gpurun_out/r9_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/r9_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['gpu_launches'], d['fwd'])" || tail -5 gpurun_out/r9_$tag.err
done
CTN_TIMING=1 timeout 300 python scratch/insitu_timing.py 2>&1 | grep -E "ctn timing|gemm|wgrad|gln|dwconv|norm_bwd"
