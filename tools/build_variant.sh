#!/bin/bash
# Build an experimental variant of libslamklt.so with extra nvcc flags for lk_tma.cu / lk_patch.cu / pyramid.cu:
#   tools/build_variant.sh NAME "-DLKT_MINB=20 -DLKT_TR=24"   ->  slam.jl_b200/csrc/variants/libslamklt_NAME.so
# Select it at run time with SLAMKLT_LIB=<path> (kernel experiments only).
set -e
cd "$(dirname "$0")/../slam.jl_b200/csrc"
name=$1; flags=$2
mkdir -p variants/obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin /usr/bin/g++"
for f in lk_tma lk_patch pyramid api lk; do sc=""; [ $f = pyramid ] && sc="--split-compile 4";
  $NV $flags $sc -Xptxas -v -c $f.cu -o variants/obj_$name/$f.o 2> variants/obj_$name/$f.log &
done
$NV $flags -fmad=false -Xptxas -v -c detect.cu -o variants/obj_$name/detect.o 2> variants/obj_$name/detect.log &
wait
$NV -shared -o variants/libslamklt_$name.so variants/obj_$name/*.o match.o brief.o host_pack.o -Xlinker --exclude-libs,ALL
grep -E "Used|spill" variants/obj_$name/lk_tma.log | head -4
