#!/bin/bash
export CTN_NO_PDL=1
for F in 9597 51184; do
echo "######## F=$F"
echo "== SS"; F=$F CTN_GEMM_SS=1 timeout 120 python scratch/ts_time.py 2>&1 | tail -4
for cl in 1 2; do for ncap in 96 128 160; do
  echo "== TS cl=$cl ncap=$ncap"; F=$F CTN_TS_CL=$cl CTN_TS_NCAP=$ncap timeout 120 python scratch/ts_time.py 2>&1 | tail -4
done; done
done
