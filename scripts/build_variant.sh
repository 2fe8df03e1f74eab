#!/bin/bash
# build_variant.sh NAME [extra nvcc -D flags...]  ->  build/variants/libmrslam_NAME.so  (select with MRSLAM_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC"
for f in runtime gicp keyframe elevation elev_post bev radon; do
  if [ $f = gicp ] || [ ! -f build/obj/$f.o ]; then
    nvcc $FL "$@" -c mr_slam_b200/csrc/$f.cu -o $out/$f.o &
  else
    cp build/obj/$f.o $out/$f.o
  fi
done
wait
nvcc -shared -o build/variants/libmrslam_$name.so $out/*.o
echo built build/variants/libmrslam_$name.so
