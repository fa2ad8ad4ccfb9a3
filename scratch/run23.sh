#!/bin/bash
timeout 200 python -m pytest tests/test_model_gpu.py -x -q -k "bf16" --tb=line 2>&1 | tail -6 | cut -c1-400
