#!/bin/bash
MODE=4 KD=256 O=256 F=102368 timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python scratch/half_probe.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in " | head -60 | cut -c1-220
