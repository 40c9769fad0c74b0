#!/bin/bash
# Builds libggd_b200.so (trainer + LPS kernels, C ABI in include/*.h) for sm_100a, in-tree.
set -e -o pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++"
mkdir -p build
pids=()
for f in gemm_tc dw_persist dw_wide dp_factor kernels ggd_train lps; do
  if [ ! -f build/$f.o ] || [ csrc/$f.cu -nt build/$f.o ] || [ -n "$(find csrc ../include -name '*.h' -newer build/$f.o -o -name '*.cuh' -newer build/$f.o)" ]; then
    $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart shared -ccbin /usr/bin/g++ -o libggd_b200.so build/gemm_tc.o build/dw_persist.o build/dw_wide.o build/dp_factor.o build/kernels.o build/ggd_train.o build/lps.o -lnccl
echo "built $(pwd)/libggd_b200.so"
