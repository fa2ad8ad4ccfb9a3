#!/bin/bash
# full GPU validation: tests, smoke, benches of every config, launch list + full ncu capture of the top kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/f3_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/f3_tests.txt; tail -6 gpurun_out/f3_tests.txt | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for c in 1 0 2 3 4; do
  timeout 600 python bench.py --config $c > gpurun_out/f3_bench_c$c.json 2> gpurun_out/f3_bench_c$c.err; python -c "
import json; d=json.load(open('gpurun_out/f3_bench_c$c.json')); print('c$c', d['dtype'], round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('fwd',{}).get('value'), round(d['roofline']['launch_us'],1), round(d['step_algorithmic']['frac_of_hbm_bound'],3), d.get('cpu_baseline',{}).get('value'))" || tail -5 gpurun_out/f3_bench_c$c.err
done
timeout 300 python bench.py --config 2 --dtype f32 --no-cpu-baseline > gpurun_out/f3_bench_c2_f32.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f3_reference_c1.json 2>/dev/null; cut -c1-200 gpurun_out/f3_reference_c1.json
# ncu: launch list of two steps, then full captures of the top kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 800 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/f3_ncu1.log 2>&1; tail -2 gpurun_out/f3_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad -s 20 -c 2 -o gpurun_out/r3_wgrad python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/f3_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|ts_gemm" -s 60 -c 6 -o gpurun_out/r3_gemm python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/f3_ncu3.log 2>&1
for n in wgrad gemm; do ncu -i gpurun_out/r3_$n.ncu-rep --page raw --csv > gpurun_out/r3_${n}_full_raw.csv 2>/dev/null; done
ls -la gpurun_out/r3_*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv_fwd_bulk|dwconv_bwd_bulk|gln_bwd_apply|norm_bwd_reduce" -s 12 -c 5 -o gpurun_out/r3_elementwise python bench.py --steps 1 --warmup 3 --profile-only > gpurun_out/f3_ncu4.log 2>&1
ncu -i gpurun_out/r3_elementwise.ncu-rep --page raw --csv > gpurun_out/r3_elementwise_full_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r3_*
