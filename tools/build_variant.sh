#!/bin/bash
# developer tool: build a variant of libb200raster.so with extra -D flags:  tools/build_variant.sh <suffix> <flags...>
set -e
suffix=$1; shift
src=/root/repo/cpu_renderer_b200/csrc
out=/tmp/variant_$suffix; mkdir -p $out
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC"
for f in setup_kernel object_walk_kernel bin_kernels edge_table_kernels raster_kernel api; do
  nvcc $FL "$@" -c $src/$f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/cpu_renderer_b200/libb200raster_$suffix.so $out/*.o -lcudart
echo built libb200raster_$suffix.so
