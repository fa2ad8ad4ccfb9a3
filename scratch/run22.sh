#!/bin/bash
timeout 200 python scratch/half_stress.py 2>&1 | tail -32
timeout 200 python -m pytest tests/test_model_gpu.py -x -q -k "bf16" 2>&1 | tail -4 | cut -c1-300
timeout 100 python bench.py --config 2 --dtype bf16 --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c2 bf16', d['value'], d['ms_per_step'], d['e2e']['value'])"
