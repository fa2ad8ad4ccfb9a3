#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_data_parallel_gpu.py -x -q 2>&1 | tail -4 | cut -c1-300
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r25_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/r25_tests.txt; tail -4 gpurun_out/r25_tests.txt | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c1', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['traffic'], d['gpu_launches'])"
