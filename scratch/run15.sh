#!/bin/bash
timeout 120 python -m pytest tests/test_model_gpu.py -x -q -k "bf16" 2>&1 | tail -5 | cut -c1-300
echo "== bench c2 bf16 no-graph"; timeout 90 python bench.py --config 2 --dtype bf16 --steps 3 --no-cpu-baseline --no-graph 2>&1 | cut -c1-200 | tail -3; echo "exit $?"
echo "== bench c2 bf16 graph"; timeout 90 python bench.py --config 2 --dtype bf16 --steps 3 --no-cpu-baseline 2>&1 | cut -c1-200 | tail -3; echo "exit $?"
echo "== bench c2 bf16 graph debug"; CTN_DEBUG_LAUNCH=1 timeout 60 python bench.py --config 2 --dtype bf16 --steps 3 --no-cpu-baseline --no-graph 2>&1 | tail -4 | cut -c1-200
