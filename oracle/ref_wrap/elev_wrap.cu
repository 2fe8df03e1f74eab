// extern "C" shim over the reference's elevation free functions (gpu_process.cu compiled UNMODIFIED against
// oracle/eigen_shim).  TEST INFRASTRUCTURE: GPU box only.  Eigen-typed arguments are filled from row-major float arrays.
#include <Eigen/Core>
#include <cuda_runtime.h>
void Init_GPU_elevationmap(int length, float resolution, float h_mahalanobisDistanceThreshold_, float h_obstacle_threshold);
void Move(float* current_Position, float resolution, int length, float* Central_coordinate, int* Start_indice, float* alignedPositionShift);
int Process_points(int* map_index, float* point_x, float* point_y, float* point_z, float* point_var, float* point_x_ts, float* point_y_ts,
                   float* point_z_ts, Eigen::Matrix4f transform, int point_num, double relativeLowerThreshold,
                   double relativeUpperThreshold, float min_r, float beam_a, float beam_c, Eigen::RowVector3f sensorJacobian,
                   Eigen::Matrix3f rotationVariance, Eigen::Matrix3f C_SB_transpose, Eigen::RowVector3f P_mul_C_BM_transpose,
                   Eigen::Matrix3f B_r_BS_skew);
void Fuse(int length, int point_num, int* point_index, int* point_colorR, int* point_colorG, int* point_colorB, float* point_intensity,
          float* point_height, float* point_var);
void Mapvar_update(int length, float var_update);
void Map_feature(int length, float* elevation, float* var, int* colorR, int* colorG, int* colorB, float* rough, float* slope, float* traver,
                 float* intensity);
void Raytracing(int length);

template <typename M>
static M fill(const float* p, int n) {
  M m;
  for (int i = 0; i < n; i++) m.d[i] = p[i];
  return m;
}
extern "C" {
void ref_elev_init(int length, float res, float mahal, float obst) {
  Init_GPU_elevationmap(length, res, mahal, obst);
  cudaDeviceSynchronize();
}
void ref_elev_move(float* pos3, float res, int length, float* central2, int* start2, float* shift2) {
  Move(pos3, res, length, central2, start2, shift2);
  cudaDeviceSynchronize();
}
int ref_elev_process_points(int* map_index, float* x, float* y, float* z, float* var, float* x_ts, float* y_ts, float* z_ts,
                            const float* T16, int n, double lower, double upper, float min_r, float beam_a, float beam_c,
                            const float* sJ, const float* rotVar, const float* CSBT, const float* PCBM, const float* BrBS) {
  return Process_points(map_index, x, y, z, var, x_ts, y_ts, z_ts, fill<Eigen::Matrix4f>(T16, 16), n, lower, upper, min_r, beam_a, beam_c,
                        fill<Eigen::RowVector3f>(sJ, 3), fill<Eigen::Matrix3f>(rotVar, 9), fill<Eigen::Matrix3f>(CSBT, 9),
                        fill<Eigen::RowVector3f>(PCBM, 3), fill<Eigen::Matrix3f>(BrBS, 9));
}
void ref_elev_fuse(int length, int n, int* idx, int* R, int* G, int* B, float* inten, float* h, float* v) {
  Fuse(length, n, idx, R, G, B, inten, h, v);
  cudaDeviceSynchronize();
}
void ref_elev_mapvar_update(int length, float upd) {
  Mapvar_update(length, upd);
  cudaDeviceSynchronize();
}
// the reference only exposes the map through Map_feature (elevation/variance/colour/intensity + derived features)
void ref_elev_get_map(int length, float* elevation, float* var, int* R, int* G, int* B, float* intensity) {
  int n = length * length;
  float* scratch = new float[3 * (size_t)n];
  Map_feature(length, elevation, var, R, G, B, scratch, scratch + n, scratch + 2 * n, intensity);
  cudaDeviceSynchronize();
  delete[] scratch;
}
// the full Map_feature output (rough / slope / traver are only defined where elevation != -10: the reference leaves the rest
// of its freshly cudaMalloc'ed output planes unwritten)
void ref_elev_map_feature(int length, float* elevation, float* var, int* R, int* G, int* B, float* rough, float* slope, float* traver,
                          float* intensity) {
  Map_feature(length, elevation, var, R, G, B, rough, slope, traver, intensity);
  cudaDeviceSynchronize();
}
void ref_elev_raytracing(int length) { Raytracing(length); }
}
