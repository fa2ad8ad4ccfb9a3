#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "bf16 or conv1x1 or dwconv_fwd or row_stats or golden" 2>&1 | tail -12 | cut -c1-250
for d in f32 bf16; do
timeout 600 python bench.py --config 2 --dtype $d --steps 10 --no-cpu-baseline > gpurun_out/r11_c2_$d.json 2> gpurun_out/r11_c2_$d.err; python -c "
import json; d=json.load(open('gpurun_out/r11_c2_$d.json')); print('c2 $d', d['value'], d['ms_per_step'], d['e2e']['value'], d['dtype'], d['step_algorithmic']['frac_of_hbm_bound'])" || tail -5 gpurun_out/r11_c2_$d.err
done
timeout 600 python bench.py --config 4 --dtype bf16 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c4 bf16', d['value'], d['ms_per_step'])"
