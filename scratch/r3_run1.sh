#!/bin/bash
# round-2 session 3, call 1: parity of the bulk-staged depthwise kernels + A/B of the graph-replayed step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== parity (CTN_DW_BULK=3)"
CTN_DW_BULK=3 timeout 600 python -m pytest tests -x -q -m gpu -k "dwconv or golden or paper or bf16 or half or causal" 2>&1 | tail -4 | cut -c1-300
echo "== step A/B"
for v in 0 1 2 3; do CTN_DW_BULK=$v timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/BULK=$v /"; done
for l in dwf3 dwb3 dwfb3; do VARIANT=scratch/variants/lib_$l.so timeout 200 python scratch/variant_bench.py 2>&1 | tail -1; done
echo "== in situ default"
CTN_DW_BULK=0 INSITU=1 CTN_NO_PDL=1 timeout 200 python scratch/variant_bench.py 2>&1 | grep -E "ctn timing|dwconv|gln_bwd|norm_bwd|tc_|ts_"
echo "== in situ bulk"
CTN_DW_BULK=3 INSITU=1 CTN_NO_PDL=1 timeout 200 python scratch/variant_bench.py 2>&1 | grep -E "ctn timing|dwconv|gln_bwd|norm_bwd"
