#!/bin/bash
# Builds libdeepfm_b200.so in-tree for sm_100a (called by __graft_entry__.build()).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p build
rm -f build/prims.o build/api.o build/csv.o build/fused_rows.o
$NVCC $FLAGS -c prims.cu -o build/prims.o &
p1=$!
$NVCC $FLAGS -c api.cu -o build/api.o &
p2=$!
$NVCC $FLAGS -c csv.cu -o build/csv.o &
p3=$!
$NVCC $FLAGS -c fused_rows.cu -o build/fused_rows.o &
p4=$!
wait $p1    # a bare `wait` would swallow a failed compile and link stale objects
wait $p2
wait $p3
wait $p4
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libdeepfm_b200.so build/prims.o build/api.o build/csv.o build/fused_rows.o -lcudart
echo "built $(cd .. && pwd)/libdeepfm_b200.so"
