#!/bin/bash
# build_variant.sh NAME [extra nvcc -D flags...]  ->  build/variants/libmrslam_NAME.so  (select with MRSLAM_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC"
TU=${VARIANT_TU:-gicp}   # the translation unit the -D flags are meant for (the others are reused from build/obj)
for f in runtime gicp keyframe elevation elev_post bev radon; do
  if [ $f = $TU ] || [ ! -f build/obj/$f.o ]; then
    FLF="$FL"; [ $f = elev_post ] && FLF="${FL/-fmad=false/}"
    nvcc $FLF "$@" -c mr_slam_b200/csrc/$f.cu -o $out/$f.o &
  else
    cp build/obj/$f.o $out/$f.o
  fi
done
wait
nvcc -shared -o build/variants/libmrslam_$name.so $out/*.o
echo built build/variants/libmrslam_$name.so
