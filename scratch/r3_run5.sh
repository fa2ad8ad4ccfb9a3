#!/bin/bash
mkdir -p gpurun_out
echo "== peer kernel, 2 ranks on one GPU"
timeout 300 python -m pytest tests/test_data_parallel_gpu.py -x -q -m gpu -k "peer_all_reduce_kernel_two_ranks" 2>&1 | tail -15 | cut -c1-250
echo "== DP with peer exchange, 2 ranks on one GPU"
timeout 400 python -m pytest tests/test_data_parallel_gpu.py -x -q -m gpu -k "with_peer_exchange_two_ranks" 2>&1 | tail -15 | cut -c1-250
echo "== flaky test x4"
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_data_parallel_gpu.py -q -m gpu -k "gloo" 2>&1 | tail -1; done
echo "== NRED fusion A/B"
timeout 200 python scratch/variant_bench.py 2>&1 | tail -1
CTN_NRED_FUSION=1 timeout 200 python scratch/variant_bench.py 2>&1 | tail -1 | sed "s/^/NRED /"
