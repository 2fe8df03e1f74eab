// elevation_b200/gpu_process.hpp -- the nine free functions of the reference's libgpu.so
// (Mapping/src/elevation_mapping_periodical/elevation_mapping/cuda/gpu_process.cu:938-1316, forward-declared at
// src/ElevationMapping.cpp:44-50, src/sensor_processors/SensorProcessorBase.cpp:34, src/RobotMotionMapUpdater.cpp:18) with the
// SAME names, global namespace and C++ signatures, implemented over the C-ABI.  Compile this header into a translation
// unit named libgpu and the ROS node links unchanged.  One hidden map per process, as in the reference.
//
// The Eigen-typed overload of Process_points is provided when <Eigen/Core> is found; a POD overload (row-major float
// arrays) always exists.  Map_feature / Raytracing run the traversability and ray-tracing kernels of
// csrc/elev_post.cu.
//
// Two ways to use it: header-only (the default: every function is `inline`), or as the real library --
// mr_slam_b200/csrc/adapters/libgpu.cpp defines ELEV_B200_FN empty and includes this file, which emits the nine functions
// out of line with the reference's mangled names into libgpu.so (built by __graft_entry__.build(); tests/test_cpp_adapters.py
// links a caller that only has the reference's forward declarations against it).
#pragma once
#ifndef ELEV_B200_FN
#define ELEV_B200_FN inline
#endif
#include <cstdio>

#include "../mrslam_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define MRSLAM_HAVE_EIGEN 1
#endif
#endif

namespace elevation_b200 {
inline mrslam_elev_t*& map() {
  static mrslam_elev_t* m = mrslam_elev_create();
  return m;
}
inline void report(int s, const char* fn) {
  if (s != MRSLAM_OK) std::fprintf(stderr, "%s failed: %s\n", fn, mrslam_last_error());  // reference: fprintf(stderr, ...) and continue
}
}  // namespace elevation_b200

ELEV_B200_FN void Init_GPU_elevationmap(int length, float resolution, float h_mahalanobisDistanceThreshold_, float h_obstacle_threshold) {
  elevation_b200::report(mrslam_elev_init(elevation_b200::map(), length, resolution, h_mahalanobisDistanceThreshold_, h_obstacle_threshold),
                         "Init_GPU_elevationmap");
}
ELEV_B200_FN void Move(float* current_Position, float resolution, int length, float* Central_coordinate, int* Start_indice,
                 float* alignedPositionShift) {
  elevation_b200::report(mrslam_elev_move(elevation_b200::map(), current_Position, resolution, length, Central_coordinate, Start_indice,
                                          alignedPositionShift), "Move");
}
// POD overload: T 4x4, sensorJacobian 1x3, rotationVariance 3x3, C_SB_transpose 3x3, P_mul_C_BM_transpose 1x3, B_r_BS_skew 3x3 (row-major)
ELEV_B200_FN int Process_points(int* map_index, float* point_x, float* point_y, float* point_z, float* point_var, float* point_x_ts,
                          float* point_y_ts, float* point_z_ts, const float* transform16, int point_num, double relativeLowerThreshold,
                          double relativeUpperThreshold, float min_r, float beam_a, float beam_c, const float* sensorJacobian3,
                          const float* rotationVariance9, const float* C_SB_transpose9, const float* P_mul_C_BM_transpose3,
                          const float* B_r_BS_skew9) {
  int s = mrslam_elev_process_points(elevation_b200::map(), map_index, point_x, point_y, point_z, point_var, point_x_ts, point_y_ts,
                                     point_z_ts, transform16, point_num, relativeLowerThreshold, relativeUpperThreshold, min_r, beam_a,
                                     beam_c, sensorJacobian3, rotationVariance9, C_SB_transpose9, P_mul_C_BM_transpose3, B_r_BS_skew9);
  elevation_b200::report(s, "Process_points");
  return 0;
}
#ifdef MRSLAM_HAVE_EIGEN
ELEV_B200_FN int Process_points(int* map_index, float* point_x, float* point_y, float* point_z, float* point_var, float* point_x_ts,
                          float* point_y_ts, float* point_z_ts, Eigen::Matrix4f transform, int point_num, double relativeLowerThreshold,
                          double relativeUpperThreshold, float min_r, float beam_a, float beam_c, Eigen::RowVector3f sensorJacobian,
                          Eigen::Matrix3f rotationVariance, Eigen::Matrix3f C_SB_transpose, Eigen::RowVector3f P_mul_C_BM_transpose,
                          Eigen::Matrix3f B_r_BS_skew) {
  // element access only (Eigen matrices are column-major, the C-ABI wants row-major arrays)
  float T[16], sj[3], rv[9], cs[9], pc[3], br[9];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) T[4 * i + j] = transform(i, j);
  for (int i = 0; i < 3; i++) {
    sj[i] = sensorJacobian(0, i);
    pc[i] = P_mul_C_BM_transpose(0, i);
    for (int j = 0; j < 3; j++) {
      rv[3 * i + j] = rotationVariance(i, j);
      cs[3 * i + j] = C_SB_transpose(i, j);
      br[3 * i + j] = B_r_BS_skew(i, j);
    }
  }
  return Process_points(map_index, point_x, point_y, point_z, point_var, point_x_ts, point_y_ts, point_z_ts, T, point_num, relativeLowerThreshold,
                        relativeUpperThreshold, min_r, beam_a, beam_c, sj, rv, cs, pc, br);
}
#endif
ELEV_B200_FN void Fuse(int length, int point_num, int* point_index, int* point_colorR, int* point_colorG, int* point_colorB,
                 float* point_intensity, float* point_height, float* point_var) {
  elevation_b200::report(mrslam_elev_fuse(elevation_b200::map(), length, point_num, point_index, point_colorR, point_colorG, point_colorB,
                                          point_intensity, point_height, point_var, nullptr), "Fuse");
}
ELEV_B200_FN void Mapvar_update(int length, float var_update) {
  elevation_b200::report(mrslam_elev_mapvar_update(elevation_b200::map(), length, var_update), "Mapvar_update");
}
ELEV_B200_FN void Map_optmove(float* opt_p, float height_update, float resolution, int length, float* opt_alignedPosition) {
  elevation_b200::report(mrslam_elev_map_optmove(elevation_b200::map(), opt_p, height_update, resolution, length, opt_alignedPosition),
                         "Map_optmove");
}
ELEV_B200_FN void Map_closeloop(float* update_position, float height_update, int length, float resolution) {
  elevation_b200::report(mrslam_elev_map_closeloop(elevation_b200::map(), update_position, height_update, length, resolution), "Map_closeloop");
}
ELEV_B200_FN void Map_feature(int length, float* elevation, float* var, int* colorR, int* colorG, int* colorB, float* rough, float* slope,
                        float* traver, float* intensity) {
  elevation_b200::report(mrslam_elev_map_feature(elevation_b200::map(), length, elevation, var, colorR, colorG, colorB, rough, slope, traver,
                                                 intensity), "Map_feature");
}
ELEV_B200_FN void Raytracing(int length) { elevation_b200::report(mrslam_elev_raytracing(elevation_b200::map(), length), "Raytracing"); }
