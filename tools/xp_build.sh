#!/bin/bash
# A/B builds of libbrawl_cuda.so with experiment macros: tools/xp_build.sh name1:"-DFOO=1" name2:"-DBAR=2 -DBAZ" ...
# -> tools/xp/lib<name>.so (select with BRAWL_CUDA_LIB).  The second translation unit is compiled once and shared.
cd "$(dirname "$0")/../brawl_b200/csrc" || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
newest=$(ls -t *.cuh *.inc *.cu | head -1)
[ -f /tmp/xp_byte_epoch.o ] && [ /tmp/xp_byte_epoch.o -nt "$newest" ] || nvcc $F -c -o /tmp/xp_byte_epoch.o byte_epoch_kernels.cu &
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc $F $defs -c -o /tmp/xp_$name.o brawl_cuda.cu &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/xp/lib$name.so /tmp/xp_$name.o /tmp/xp_byte_epoch.o && echo built tools/xp/lib$name.so
done
