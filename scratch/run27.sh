#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r27_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/r27_tests.txt; tail -4 gpurun_out/r27_tests.txt | cut -c1-300; grep -n "Error" gpurun_out/r27_tests.txt | head -3 | cut -c1-200
