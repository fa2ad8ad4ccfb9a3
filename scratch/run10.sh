#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r10_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/r10_tests.txt; tail -6 gpurun_out/r10_tests.txt | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r10_bench.json')); print('c1', d['value'], d['ms_per_step'], d['e2e']['value'], d['fwd'], d['roofline']['launch_us'], d['roofline']['traffic'])" || tail -5 gpurun_out/r10_bench.err
