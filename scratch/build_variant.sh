#!/bin/bash
# usage: scratch/build_variant.sh NAME "-DCTN_DW_TJ=8 -DCTN_DW_U=8"   -> scratch/variants/lib_NAME.so (elementwise.cu rebuilt)
set -e
cd /root/repo
mkdir -p scratch/variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
SRC=${3:-elementwise}
/usr/local/cuda/bin/nvcc $FLAGS $2 -c conv_tasnet_b200/csrc/$SRC.cu -o scratch/variants/${SRC}_$1.o
OBJS=$(ls conv_tasnet_b200/build/*.o | grep -v "/$SRC.o")
/usr/local/cuda/bin/nvcc -shared -o scratch/variants/lib_$1.so $OBJS scratch/variants/${SRC}_$1.o -lcuda
echo built scratch/variants/lib_$1.so
