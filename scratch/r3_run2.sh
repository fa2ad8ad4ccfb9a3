#!/bin/bash
mkdir -p gpurun_out
echo "== parity (CTN_DW_BULK=31, fusion on)"
CTN_DW_BULK=31 CTN_APPLY_FUSION=1 timeout 600 python -m pytest tests -x -q -m gpu -k "dwconv or norm_bwd or golden or paper or fused or causal or smoke" 2>&1 | tail -4 | cut -c1-300
echo "== step A/B"
for v in 3 11 19; do CTN_DW_BULK=$v timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/BULK=$v /"; done
CTN_DW_BULK=7 CTN_APPLY_FUSION=1 timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/BULK=7+fusion /"
CTN_DW_BULK=23 CTN_APPLY_FUSION=1 timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/BULK=23+fusion /"
echo "== in situ all on"
CTN_DW_BULK=31 CTN_APPLY_FUSION=1 INSITU=1 CTN_NO_PDL=1 timeout 200 python scratch/variant_bench.py 2>&1 | grep -E "ctn timing|dwconv|gln_bwd|norm_bwd"
