#!/bin/bash
# Recompiles only fused_rows.cu (+ api.cu with "api" as the first argument) and relinks libdeepfm_b200.so: the inner loop of kernel experiments.
set -e
cd "$(dirname "$0")/../recommender_tensorflow_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
if [ "$1" = "api" ]; then $NVCC $FLAGS -c api.cu -o build/api.o & fi
$NVCC $FLAGS -Xptxas -v -c fused_rows.cu -o build/fused_rows.o 2>&1 | grep -E "Compiling entry.*fused_rows_kernelILi16ELi16ELi2|spill|registers" | grep -A2 "ELi2ELb0" | tail -2
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libdeepfm_b200.so build/prims.o build/api.o build/csv.o build/fused_rows.o -lcudart
echo relinked
