#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/half_stress.py 2>&1 | tail -3
CTN_TS_MASK=7 timeout 200 python scratch/half_stress.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r26_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/r26_tests.txt; tail -4 gpurun_out/r26_tests.txt | cut -c1-300
for c in 1 2 3; do timeout 600 python bench.py --config $c --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c$c', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"; done
F=51184 CTN_NO_PDL=1 timeout 120 python scratch/ts_time.py 2>&1 | tail -4
