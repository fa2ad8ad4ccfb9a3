#!/bin/bash
timeout 200 python scratch/variant_bench.py 2>&1 | tail -1
for l in ga8 nr8 ganr8 ga8tk8 ga8tk24 ga8mb ga6tk18; do VARIANT=scratch/variants/lib_$l.so timeout 200 python scratch/variant_bench.py 2>&1 | tail -1; done
timeout 200 python scratch/variant_bench.py 2>&1 | tail -1
