#!/bin/bash
export CTN_NO_PDL=1
F=51184 timeout 120 python scratch/ts_time.py > /dev/null 2>&1 || exit 1
F=51184 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ts_gemm -s 4 -c 4 -o gpurun_out/ts_f51k python scratch/ts_time.py > gpurun_out/ts6_ncu.log 2>&1
tail -3 gpurun_out/ts6_ncu.log
ncu -i gpurun_out/ts_f51k.ncu-rep --page raw --csv > gpurun_out/ts_f51k_raw.csv 2>/dev/null
ls -la gpurun_out/ts_f51k*
