#!/bin/bash
timeout 200 python -m pytest tests/test_model_gpu.py -x -q -k "bf16" 2>&1 | tail -5 | cut -c1-300
echo "== bench c2 bf16 steps 10"; BENCH_WATCHDOG=45 timeout 80 python bench.py --config 2 --dtype bf16 --steps 10 --no-cpu-baseline 2>&1 | cut -c1-220 | tail -30; echo "exit $?"
