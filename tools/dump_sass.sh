#!/bin/bash
# developer tool: SASS listings of the kernels the bench runs, from the built library -> profiles/r02_sass_*.txt
set -e
cd /root/repo
lib=cpu_renderer_b200/libb200raster.so
dump() { cuobjdump -sass -fun "$1" $lib 2>/dev/null | sed -n '/Function :/,$p' > "profiles/$2"; echo "$2: $(grep -c '/\*[0-9a-f]\{4\}\*/' profiles/$2) instructions"; }
dump _ZN5b200r12setup_kernelILb0ELb0ELb0ELb0EEEvNS_10ViewParamsENS_10MeshParamsENS_12SetupOutputsE r02_sass_setup_kernel_plain.txt
dump _ZN5b200r12setup_kernelILb0ELb0ELb0ELb1EEEvNS_10ViewParamsENS_10MeshParamsENS_12SetupOutputsE r02_sass_setup_kernel_split.txt
dump _ZN5b200r12setup_kernelILb0ELb0ELb1ELb0EEEvNS_10ViewParamsENS_10MeshParamsENS_12SetupOutputsE r02_sass_setup_kernel_listed.txt
dump _ZN5b200r14scatter_kernelENS_13ScatterParamsE r02_sass_scatter_kernel.txt
dump _ZN5b200r13select_kernelENS_10ViewParamsENS_10MeshParamsEPjS2_S2_ r02_sass_select_kernel.txt
for m in 0 1 2; do
  dump _ZN5b200r13raster_kernelILi128ELi8ELi4ELi${m}EEEvNS_12RasterParamsE r02_sass_raster_kernelILi128ELi8ELi4ELi${m}.txt
  dump _ZN5b200r13raster_kernelILi64ELi16ELi8ELi${m}EEEvNS_12RasterParamsE r02_sass_raster_kernelILi64ELi16ELi8ELi${m}.txt
done
