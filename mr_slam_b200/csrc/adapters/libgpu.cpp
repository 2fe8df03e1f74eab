// libgpu.so of the B200 drop-in: the nine free functions of the reference's elevation_mapping CUDA library
// (EM/cuda/gpu_process.cu:938-1316; declared at EM/src/ElevationMapping.cpp:44-50, sensor_processors/SensorProcessorBase.cpp:34,
// RobotMotionMapUpdater.cpp:18) emitted OUT OF LINE, C++-mangled, in the global namespace, over the C-ABI of libmrslam_b200.so.
#define ELEV_B200_FN
#include "elevation_b200/gpu_process.hpp"
#ifndef MRSLAM_HAVE_EIGEN
#error "libgpu.so needs <Eigen/Core> (or the ABI shim under tests/cpp/abi_shim) for Process_points' Eigen signature"
#endif
