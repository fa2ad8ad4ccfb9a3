#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/variant_bench.py 2>&1 | tail -1
for l in f8 f12 f24 f32 b8 b12; do VARIANT=scratch/variants/lib_$l.so timeout 200 python scratch/variant_bench.py 2>&1 | tail -1; done
