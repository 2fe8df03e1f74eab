#!/bin/bash
# oracle/build_ref.sh -- compile the pieces of the reference that build from their own few source files, from where
# they lie under $REF, into oracle/_ref/ (git-ignored; travels to the GPU box).  No reference source is copied.
# What is NOT built and why: fast_gicp (needs PCL + Eigen, absent), the torch-radon python extension (needs the
# reference's make.py / libtorch ABI flags), ROS nodes.
set -u
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT"
CXX=/usr/bin/g++
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
LD=$REF/LoopDetection
ok=0; fail=0
run() { echo "+ $*" >> "$OUT/build.log"; if "$@" >> "$OUT/build.log" 2>&1; then ok=$((ok+1)); else fail=$((fail+1)); echo "[build_ref] FAILED: $*" ; fi; }
: > "$OUT/build.log"
# 1. DiSCO polar voxeliser, CPU twin (runs anywhere)
P=$LD/src/disco_ros/tools/multi-layer-polar-cpu/cython/src
run $CXX -O2 -fPIC -shared -w -I"$P" -o "$OUT/libpolar_cpu_ref.so" "$HERE/ref_wrap/polar_cpu_wrap.cpp" "$P/manager.cpp"
# 2. torch-radon analytic oracle (CPU)
T=$LD/torch-radon
run $CXX -O2 -fPIC -shared -w -std=gnu++17 -I"$T/include" -I/usr/local/cuda/include -include cuda_runtime.h -o "$OUT/libsymbolic_ref.so" -x c++ "$HERE/ref_wrap/symbolic_wrap.cpp" "$T/src/symbolic.cpp" "$T/src/parameter_classes.cu"
# 3. GPU voxelisers (reference kernels + host scatter), sm_100
V=$LD/generate_bev_cython_binary/src
run $NVCC -arch=sm_100 -O2 -w -Xcompiler -fPIC -shared -I"$V" -DREF_FN=ref_voxelocc_gpu -o "$OUT/libvoxelocc_ref.so" "$HERE/ref_wrap/voxel_gpu_wrap.cu" "$V/manager.cu"
G=$LD/src/disco_ros/tools/multi-layer-polar-gpu/cython/src
run $NVCC -arch=sm_100 -O2 -w -Xcompiler -fPIC -shared -I"$G" -DREF_FN=ref_polar_gpu -o "$OUT/libpolar_gpu_ref.so" "$HERE/ref_wrap/voxel_gpu_wrap.cu" "$G/manager.cu"
F=$LD/generate_bev_pointfeat_cython/src
run $NVCC -arch=sm_100 -O2 -w -Xcompiler -fPIC -shared -I"$F" -o "$OUT/libvoxelfeat_ref.so" "$HERE/ref_wrap/voxelfeat_gpu_wrap.cu" "$F/manager.cu"
# 4. torch-radon forward projection (reference CUDA kernel with hardware texture filtering), sm_100
run $NVCC -arch=sm_100 -O2 -w -std=c++17 -Xcompiler -fPIC -shared -I"$T/include" -o "$OUT/libradon_ref.so" "$HERE/ref_wrap/radon_wrap.cu" "$T/src/forward.cu" "$T/src/texture.cu" "$T/src/parameter_classes.cu" "$T/src/log.cpp"
# 5. elevation mapping kernels against the minimal Eigen shim (oracle/eigen_shim), sm_100
E=$REF/Mapping/src/elevation_mapping_periodical/elevation_mapping/cuda
if [ -f "$HERE/ref_wrap/elev_wrap.cu" ]; then
run $NVCC -arch=sm_100 -O2 -w -std=c++17 -Xcompiler -fPIC -shared -I"$HERE/eigen_shim" -o "$OUT/libelev_ref.so" "$HERE/ref_wrap/elev_wrap.cu" "$E/gpu_process.cu"
fi
echo "[build_ref] built $ok, failed $fail (log: $OUT/build.log)"
ls -la "$OUT"
