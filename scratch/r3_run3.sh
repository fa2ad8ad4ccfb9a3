#!/bin/bash
mkdir -p gpurun_out
echo "== parity CTN_TC_CL=2 (kernel tests first, each bounded)"
CTN_TC_CL=2 CTN_GEMM_SS=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "conv1x1" 2>&1 | tail -3 | cut -c1-300
CTN_TC_CL=2 timeout 400 python -m pytest tests -x -q -m gpu -k "conv1x1 or golden or paper or smoke" 2>&1 | tail -3 | cut -c1-300
echo "== step A/B"
for v in 1 2; do CTN_TC_CL=$v timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/TC_CL=$v /"; done
CTN_TC_CL=2 CTN_GEMM_SS=1 timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/TC_CL=2 SS /"
CTN_TC_CL=1 CTN_GEMM_SS=1 timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/TC_CL=1 SS /"
echo "== in situ cl=2"
CTN_TC_CL=2 INSITU=1 CTN_NO_PDL=1 timeout 200 python scratch/variant_bench.py 2>&1 | grep -E "ctn timing|tc_|ts_"
