#!/bin/bash
timeout 600 python -m pytest tests/test_data_parallel_gpu.py -x -q 2>&1 | tail -3 | cut -c1-260
