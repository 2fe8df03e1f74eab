// A caller that knows libgpu.so only through the reference's own forward declarations -- restated here as the boundary
// specification with their source lines -- and therefore links only if libgpu.so exports exactly those C++-mangled names.
// Built with -Wl,--no-undefined against mr_slam_b200/libgpu.so by tests/test_cpp_adapters.py; runs a frame when a GPU is present.
#include <Eigen/Core>
#include <cmath>
#include <cstdio>
#include <vector>

// EM/src/ElevationMapping.cpp:44-50
void Move(float* current_Position, float resolution, int length, float* h_central_coordinate, int* h_start_indice, float* position_shift);
void Init_GPU_elevationmap(int length, float resolution, float h_mahalanobisDistanceThreshold_, float h_obstacle_threshold);
void Map_closeloop(float* update_position, float height_update, int length, float resolution);
void Raytracing(int length_);
void Fuse(int length, int point_num, int* point_index, int* point_colorR, int* point_colorG, int* point_colorB, float* point_intensity,
          float* point_height, float* point_var);
void Map_feature(int length, float* elevation, float* var, int* point_colorR, int* point_colorG, int* point_colorB, float* rough, float* slope,
                 float* traver, float* intensity);
void Map_optmove(float* opt_p, float height_update, float resolution, int length, float* opt_alignedPosition);
// EM/src/sensor_processors/SensorProcessorBase.cpp:34
int Process_points(int* mapindex, float* point_x, float* point_y, float* point_z, float* point_var, float* point_x_ts, float* point_y_ts,
                   float* point_z_ts, Eigen::Matrix4f Transform, int point_num, double relativeLowerThreshold, double relativeUpperThreshold,
                   float min_r, float beam_a, float beam_c, Eigen::RowVector3f sensorJacobian, Eigen::Matrix3f rotationVariance,
                   Eigen::Matrix3f C_SB_transpose, Eigen::RowVector3f P_mul_C_BM_transpose, Eigen::Matrix3f B_r_BS_skew);
// EM/src/RobotMotionMapUpdater.cpp:18 (declared int there, defined void in gpu_process.cu:1144: the return type is not mangled)
int Mapvar_update(int length, float update_var);

extern "C" int mrslam_init(int);
extern "C" const char* mrslam_last_error(void);

int main() {
  if (mrslam_init(0) != 0) {
    std::printf("NO_GPU %s\n", mrslam_last_error());
    return 0;
  }
  const int L = 60, n = 2000;
  Init_GPU_elevationmap(L, 0.2f, 2.5f, 0.2f);
  float pos[3] = {0.4f, -0.2f, 0.f}, cc[2], sh[2];
  int si[2];
  Move(pos, 0.2f, L, cc, si, sh);
  unsigned s = 777;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  std::vector<float> x(n), y(n), z(n), var(n), xt(n), yt(n), zt(n), inten(n, 0.f);
  std::vector<int> mi(n), zero(n, 0);
  for (int i = 0; i < n; i++) { x[i] = rnd() * 10 - 5; y[i] = -1.5f - rnd() * 4; z[i] = -1.0f + 0.1f * rnd(); }
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  T(0, 3) = 0.4f;
  T(1, 3) = -0.2f;
  Eigen::RowVector3f sJ, P;
  sJ(0) = 0; sJ(1) = 0; sJ(2) = 1;
  P(0) = 0; P(1) = 0; P(2) = 1;
  Eigen::Matrix3f Z = Eigen::Matrix3f::Zero(), I = Eigen::Matrix3f::Identity();
  Process_points(mi.data(), x.data(), y.data(), z.data(), var.data(), xt.data(), yt.data(), zt.data(), T, n, -5.0, 5.0, 0.018f, 0.0006f, 0.0015f, sJ, Z,
                 I, P, Z);
  Fuse(L, n, mi.data(), zero.data(), zero.data(), zero.data(), inten.data(), zt.data(), var.data());
  Mapvar_update(L, 0.001f);
  std::vector<float> elev(L * L), v2(L * L), rough(L * L), slope(L * L), trav(L * L), in2(L * L);
  std::vector<int> R(L * L), G(L * L), B(L * L);
  Map_feature(L, elev.data(), v2.data(), R.data(), G.data(), B.data(), rough.data(), slope.data(), trav.data(), in2.data());
  Raytracing(L);
  float opt[2] = {0.5f, -0.1f}, al[2];
  Map_optmove(opt, 0.01f, 0.2f, L, al);
  float up[2] = {0.6f, 0.f};
  Map_closeloop(up, -0.01f, L, 0.2f);
  int valid = 0, cells = 0;
  bool shifted = true;
  for (int i = 0; i < n; i++) {
    valid += mi[i] >= 0;
    if (mi[i] >= 0) shifted = shifted && std::fabs(xt[i] - (x[i] + 0.4f)) < 1e-5f;  // column-major Eigen -> row-major C-ABI
  }
  for (int c = 0; c < L * L; c++) cells += elev[c] != 0.f && elev[c] != -10.f;
  std::printf("valid %d cells %d shifted %d\n", valid, cells, (int)shifted);
  const bool ok = valid > 1000 && cells > 100 && shifted;
  std::printf(ok ? "LIBGPU_OK\n" : "LIBGPU_FAIL\n");
  return ok ? 0 : 1;
}
