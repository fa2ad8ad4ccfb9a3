#!/bin/bash
# full GPU validation: tests, smoke, benches of every config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/full_tests.txt; tail -15 gpurun_out/full_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/full_bench_c1.json 2> gpurun_out/full_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_c1.json')); print('c1', d['value'], d['ms_per_step'], d['e2e']['value'], d['fwd'], d['roofline']['launch_us'])" || tail -5 gpurun_out/full_bench_c1.err
for c in 0 2 3 4; do
  timeout 900 python bench.py --config $c --steps 10 > gpurun_out/full_bench_c$c.json 2> gpurun_out/full_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_c$c.json')); print('c$c', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('fwd'), d['roofline']['launch_us'], d['step_algorithmic']['frac_of_hbm_bound'], d.get('cpu_baseline',{}).get('value'))" || tail -5 gpurun_out/full_bench_c$c.err
done
