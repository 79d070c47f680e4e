#!/bin/bash
# development shortcut: recompile ONE translation unit of the library and relink (a header change normally rebuilds all 27;
# use this only while iterating on a kernel that a single .cu instantiates, then run `make` before committing).
set -e
cd "$(dirname "$0")/../rstsr_b200/csrc"
for tu in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -fmad=false \
    --compress-mode=size -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v -c $tu.cu -o build/$tu.o 2> build/$tu.ptxas.log || { cat build/$tu.ptxas.log; exit 1; }
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/librstsr_cuda.so build/*.o -ldl
