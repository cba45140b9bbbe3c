#!/bin/bash
# A/B builds of libbrawl_cuda.so with experiment macros: tools/xp_build.sh name1:"-DFOO=1" name2:"-DBAR=2 -DBAZ" ...
# -> tools/xp/lib<name>.so (select with BRAWL_CUDA_LIB).  Both translation units get the macros.
cd "$(dirname "$0")/../brawl_b200/csrc" || exit 1
mkdir -p ../../tools/xp
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc $F $defs -c -o /tmp/xp_$name.o brawl_cuda.cu &
  nvcc $F $defs -c -o /tmp/xp_${name}_byte.o byte_epoch_kernels.cu &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/xp/lib$name.so /tmp/xp_$name.o /tmp/xp_${name}_byte.o && echo built tools/xp/lib$name.so
done
