#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f4_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/f4_tests.txt; tail -4 gpurun_out/f4_tests.txt | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/f4_bench_c1.json 2> gpurun_out/f4_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/f4_bench_c1.json')); print('c1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks'])"
timeout 300 python bench.py --no-graph --no-cpu-baseline --steps 10 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('eager (PDL)', round(d['value']), round(d['ms_per_step'],3))"
