#!/bin/bash
timeout 300 python -m pytest tests -x -q -m gpu -k "row_stats or dwconv_fwd or bf16 or golden or paper_width" 2>&1 | tail -3 | cut -c1-200
timeout 200 python scratch/insitu_c2.py 2>&1 | grep -E "===|row_stats|dwconv|ctn timing"
for c in 2; do timeout 300 python bench.py --config $c --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c$c', round(d['value']), round(d['ms_per_step'],3))"; done
timeout 300 python bench.py --config 2 --dtype f32 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c2 f32', round(d['value']), round(d['ms_per_step'],3))"
