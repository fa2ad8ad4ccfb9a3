#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv1x1" 2>&1 | tail -3
TAG=TS timeout 200 python scratch/ts_diag.py 2>&1 | tail -5; rm -f gpurun_out/est_*.pt
for cl in 1 2 4; do echo "== CTN_TS_CL=$cl"; CTN_TS_CL=$cl timeout 120 python scratch/ts_time.py 2>&1 | tail -5; done
echo "== SS"; CTN_GEMM_SS=1 timeout 120 python scratch/ts_time.py 2>&1 | tail -5
CTN_TS_CL=2 CTN_B200_LIB=/root/repo/scratch/variants/lib_tstrace.so timeout 200 python scratch/ts_trace.py 2>&1 | grep -v "^ CTA 13\|^ CTA 0" | grep -A40 "tf32 up" | tail -45
CTN_TIMING=1 timeout 300 python scratch/insitu_timing.py 2>&1 | grep -E "ctn timing|gemm|wgrad|gln|dwconv|norm_bwd"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/ts4_bench.json 2> gpurun_out/ts4_bench.err; cat gpurun_out/ts4_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['fwd'])" || tail -5 gpurun_out/ts4_bench.err
