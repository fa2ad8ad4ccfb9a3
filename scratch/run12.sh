#!/bin/bash
for m in 0 2 3 4; do MODE=$m timeout 40 python scratch/half_probe.py 2>&1 | tail -2; echo "exit $?"; done
MODE=4 KD=512 O=256 timeout 40 python scratch/half_probe.py 2>&1 | tail -2
MODE=3 F=102368 timeout 40 python scratch/half_probe.py 2>&1 | tail -2
