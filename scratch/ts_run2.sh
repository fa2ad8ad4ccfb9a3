#!/bin/bash
mkdir -p gpurun_out
for cfg in "SS:CTN_GEMM_SS=1" "TS:CTN_TS_MASK=7" "TSnoPDL:CTN_TS_MASK=7 CTN_NO_PDL=1" "TS1:CTN_TS_MASK=1" "TS2:CTN_TS_MASK=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "=== $tag ($envs)"
  env $envs TAG=$tag timeout 200 python scratch/ts_diag.py 2>&1 | tail -6
done
python - <<'PY'
import torch
a = torch.load("gpurun_out/est_SS.pt")
for t in ["TS", "TSnoPDL", "TS1", "TS2"]:
    try:
        b = torch.load(f"gpurun_out/est_{t}.pt")
        print(t, "vs SS: max abs diff", (a - b).abs().max().item(), "rel", ((a - b).abs().max() / a.abs().max()).item())
    except Exception as e:
        print(t, "missing", e)
PY
rm -f gpurun_out/est_*.pt
# cycle trace of the TS kernel (library variant built with -DCTN_TS_TRACE)
CTN_B200_LIB=/root/repo/scratch/variants/lib_tstrace.so timeout 200 python scratch/ts_trace.py 2>&1 | tail -80
