#!/bin/bash
echo "== graph cl=1 steps 10"; CTN_TS_CL=1 timeout 60 python bench.py --config 2 --dtype bf16 --steps 10 --no-cpu-baseline 2>&1 | cut -c1-160 | tail -2; echo "exit $?"
echo "== no-graph steps 10"; timeout 60 python bench.py --config 2 --dtype bf16 --steps 10 --no-cpu-baseline --no-graph 2>&1 | cut -c1-160 | tail -2; echo "exit $?"
echo "== graph steps 10 M=3 (config 1 fwd bf16)"; timeout 60 python bench.py --config 1 --mode fwd --dtype bf16 --steps 10 --no-cpu-baseline 2>&1 | cut -c1-160 | tail -2; echo "exit $?"
