#!/bin/bash
# Variant builds of the block solver (chol.cu compiled with -DCHOL_VARIANT=n), linked against the other objects of the
# regular build: build/libdbslmm_b200_v<n>.so, selected at run time with DBSLMM_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python -m dbslmm_b200.build > /dev/null
for v in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden \
      -DCHOL_VARIANT=$v -c dbslmm_b200/csrc/chol.cu -o build/chol_v$v.o
  objs=$(ls build/*.o | grep -v "chol" | tr '\n' ' ')
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o build/libdbslmm_b200_v$v.so $objs build/chol_v$v.o -cudart static
  echo "built build/libdbslmm_b200_v$v.so"
done
