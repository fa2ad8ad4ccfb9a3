#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f6_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/f6_tests.txt; tail -3 gpurun_out/f6_tests.txt | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for c in 1 0 2 3 4; do
  timeout 600 python bench.py --config $c > gpurun_out/f6_bench_c$c.json 2> gpurun_out/f6_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/f6_bench_c$c.json')); r=d['roofline']; print('c$c', d['dtype'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), r['bound'], round(r['frac'],3), round(r['other_roof']['frac'],3), round(r['launch_us'],1), round(d['step_algorithmic']['frac_of_hbm_bound'],3), d.get('cpu_baseline',{}).get('value'))" || tail -5 gpurun_out/f6_bench_c$c.err
done
timeout 300 python bench.py --config 2 --dtype f32 --no-cpu-baseline > gpurun_out/f6_bench_c2_f32.json 2>/dev/null
