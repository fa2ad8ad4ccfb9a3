#!/bin/bash
CTN_NO_PDL=1 CTN_B200_LIB=/root/repo/scratch/variants/lib_tctrace.so timeout 100 python scratch/wg_trace2.py 2>&1 | tail -4
