#!/bin/bash
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q 2>&1 | grep -v "^\[rank\|NCCL\|^$" | grep -B2 -A12 "Error\|error\|assert" | head -70 | cut -c1-260
