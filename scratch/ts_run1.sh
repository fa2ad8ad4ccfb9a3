#!/bin/bash
# first GPU run of the frame-major (TS) GEMM: kernel parity, model parity, in-place timing, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv1x1" > gpurun_out/ts1_kern.txt 2>&1; echo "kern exit $?" >> gpurun_out/ts1_kern.txt
tail -15 gpurun_out/ts1_kern.txt
timeout 600 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/ts1_model.txt 2>&1; echo "model exit $?" >> gpurun_out/ts1_model.txt
tail -15 gpurun_out/ts1_model.txt
CTN_TIMING=1 timeout 300 python scratch/insitu_timing.py > gpurun_out/ts1_insitu.txt 2>&1; tail -40 gpurun_out/ts1_insitu.txt
timeout 600 python bench.py > gpurun_out/ts1_bench.json 2> gpurun_out/ts1_bench.err; cat gpurun_out/ts1_bench.json
CTN_GEMM_SS=1 timeout 600 python bench.py > gpurun_out/ts1_bench_ss.json 2> gpurun_out/ts1_bench_ss.err; cat gpurun_out/ts1_bench_ss.json
