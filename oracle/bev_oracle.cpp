// oracle/bev_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the RING Cartesian voxeliser (python module `voxelocc`) and the DiSCO
// polar voxeliser (python module `gputransform`).  Paths relative to /root/reference/LoopDetection/:
//   generate_bev_cython_binary/src/kernel.cu:14-61     point2gridmap (index math, f64)
//   generate_bev_cython_binary/src/manager.cu:53-92    GPUTransformer::retreive (host scatter, point order)
//   src/disco_ros/tools/multi-layer-polar-gpu/cython/src/kernel.cu:24-80   xy2theta, point2gridmap (polar)
//   src/disco_ros/tools/multi-layer-polar-gpu/cython/src/manager.cu:53-81  retreive (first `enough_large` pts / voxel)
// Float/double promotion follows the PTX nvcc generates for those kernels (checked with nvcc -ptx):
// the Cartesian index is pure IEEE double math; the polar kernel takes atan in FLOAT (CUDA atanf),
// squares and sqrt in double, the 180/pi scaling in double with FMA contraction.
//
// PARITY PIN: the reference only has visual smoke tests for these modules (BEV/test.py) -- no
// assertions.  Pins used instead: (1) the reference's CPU twin of the polar voxeliser
// (multi-layer-polar-cpu/cython/src/*.cpp) compiled unmodified into oracle/_ref and run on the
// reference's own fixtures 1.bin/2.bin/test.bin; (2) the reference GPU classes compiled unmodified
// into oracle/_ref and run on the GPU box (-m gpu tests).
//
// Out-of-range voxel indices (x == 1.0 exactly, |z| > max_height in the polar variant, theta that
// rounds to 360) make the reference write out of bounds (undefined behaviour); here such points
// are skipped and counted in *n_dropped.
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {

// voxelocc: points SoA float[3*size] (x.., y.., z..); out float[num_x*num_y*num_height*enough_large*3], caller-zeroed.
// idx_out (optional) int[3*size]: x,y,height index per point (kernel output).
int orc_voxelocc(const float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_height,
                 int enough_large, float* out, int* idx_out, int* n_dropped) {
  (void)enough_large;
  float gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
  float gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
  float gap_h = (float)(2.0 * (double)(float)max_height / (double)(float)num_height);
  std::vector<double> max_h((size_t)num_x * num_y, 0.0);
  int dropped = 0;
  for (int i = 0; i < size; i++) {
    float x = point[i], y = point[i + size], z = point[i + 2 * (size_t)size];
    if (x == 0.0) x = 0.0001;
    if (y == 0.0) y = 0.0001;
    if (z == 0.0) z = 0.0001;
    if (x > 1.0) x = 0.9999;
    if (y > 1.0) y = 0.9999;
    if (z > 1.0) z = 0.9999;
    if (x < -1.0) x = -0.9999;
    if (y < -1.0) y = -0.9999;
    if (z < -1.0) z = -0.9999;
    int ix = (int)std::floor(((double)x + 1.0) / (double)gap_x);
    int iy = (int)std::floor(((double)y + 1.0) / (double)gap_y);
    int ih = (int)std::floor(((double)z + 1.0) / (double)gap_h);
    if (idx_out) {
      idx_out[i] = ix;
      idx_out[i + size] = iy;
      idx_out[i + 2 * (size_t)size] = ih;
    }
    if (ix < 0 || ix >= num_x || iy < 0 || iy >= num_y || ih < 0 || ih >= num_height) {
      dropped++;
      continue;
    }
    size_t slot = 3 * ((size_t)iy + (size_t)ix * num_y + (size_t)ih * num_y * num_x);
    out[slot + 0] = point[i];
    out[slot + 1] = point[i + size];
    // host scatter uses the ORIGINAL (unclamped) z from point_host (manager.cu:71-73)
    double zo = (double)point[i + 2 * (size_t)size];
    if (max_h[(size_t)iy + (size_t)ix * num_y] < zo) {
      out[slot + 2] = point[i + 2 * (size_t)size];
      max_h[(size_t)iy + (size_t)ix * num_y] = zo;
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return 0;
}

// theta as the reference GPU kernel computes it (see header); atan_f = float atan implementation to use
static inline float polar_theta(float x, float y) {
  const double k = 180 / M_PI;
  if (x >= 0 && y >= 0) return (float)(k * (double)atanf(y / x));
  else if (x < 0 && y >= 0) return (float)fma((double)atanf(y / (-x)), -k, 180.0);
  else if (x < 0 && y < 0) return (float)fma((double)atanf(y / x), k, 180.0);
  else return (float)fma((double)atanf((-y) / x), -k, 360.0);
}

// gputransform (polar): out float[num_ring*num_sector*num_height*enough_large*3], caller-zeroed.
int orc_polar(const float* point, int size, int max_length, int max_height, int num_ring, int num_sector, int num_height,
              int enough_large, float* out, int* idx_out, int* n_dropped) {
  float gap_ring = (float)max_length / (float)num_ring;
  float gap_sector = (float)(360.0 / (double)(float)num_sector);
  float gap_h = (float)(2.0 * (double)(float)max_height / (double)(float)num_height);
  size_t grid = (size_t)num_ring * num_sector * num_height;
  std::vector<int> counter(grid, 0);
  int dropped = 0;
  for (int i = 0; i < size; i++) {
    float x = point[i], y = point[i + size], z = point[i + 2 * (size_t)size];
    if (x == 0.0) x = 0.0001;
    if (y == 0.0) y = 0.0001;
    if (z == 0.0) z = 0.0001;
    float theta = polar_theta(x, y);
    float faraway = (float)std::sqrt((double)x * (double)x + (double)y * (double)y);
    int ir = (int)std::floor(faraway / gap_ring);
    int is = (int)std::floor(theta / gap_sector);
    int ih = (int)std::floor((z + (float)max_height) / gap_h);
    if (ir >= num_ring) ir = num_ring - 1;
    if (idx_out) {
      idx_out[i] = ir;
      idx_out[i + size] = is;
      idx_out[i + 2 * (size_t)size] = ih;
    }
    if (ir < 0 || is < 0 || is >= num_sector || ih < 0 || ih >= num_height) {
      dropped++;
      continue;
    }
    size_t cell = (size_t)is + (size_t)ir * num_sector + (size_t)ih * num_sector * num_ring;
    if (counter[cell] < enough_large) {
      size_t slot = 3 * (cell + (size_t)counter[cell] * grid);
      out[slot + 0] = point[i];
      out[slot + 1] = point[i + size];
      out[slot + 2] = 1;
      counter[cell]++;
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return 0;
}
}

// -----------------------------------------------------------------------------------------------------------------
// voxelfeat (RING++), LoopDetection/generate_bev_pointfeat_cython/src/kernel.cu:16-164, manager.cu:18-181
// -----------------------------------------------------------------------------------------------------------------
extern "C" {

// GPUFeatureExtractor::get_features: pointcloud AoS float[3*size] (xyzxyz), neighbours int[size*k], eigens float[size*5]
// (3 eigenvalues of the 3-D covariance, descending, then 2 of the 2-D one); feature float[size*featsize], featsize >= 13.
// Float/double promotion follows the PTX nvcc emits for kernel.cu:16-104: pow() and the 1/3 root in double, log in float.
void orc_voxelfeat_features(const float* pc, int size, int featsize, int k, const int* nbr, const float* eig, float* feature) {
  const float eps = 0.0f;
  for (int g = 0; g < size; g++) {
    float e3[3], e2[2];
    float s3 = 0.f, p3 = 1.f, s2 = 0.f;
    for (int i = 0; i < 5; i++) {
      if (i < 3) {
        e3[i] = eig[g * 5 + i];
        s3 += e3[i];
        p3 = p3 * e3[i];
      } else {
        e2[i - 3] = eig[g * 5 + i];
        s2 += e2[i - 3];
      }
    }
    float C_ = e3[2] / (s3 + eps);
    float O_ = (float)std::pow((double)p3 / (std::pow((double)s3, 3.0) + (double)eps), 1.0 / 3.0);
    float L_ = (e3[0] - e3[1]) / (e3[0] + eps);
    float E_ = 0.f;
    for (int i = 0; i < 3; i++) E_ += (e3[i] / (s3 + eps)) * logf(e3[i] / (s3 + eps));
    E_ = -E_;
    float P_ = (e3[1] - e3[2]) / e3[0];
    float S_ = e3[2] / e3[0];
    float A_ = (e3[0] - e3[2]) / e3[0];
    float X_ = s3;
    float D_ = (float)((double)(3 * k) / ((4 * M_PI * (double)p3) + (double)eps));
    float S_2 = s2;
    float L_2 = e2[1] / e2[0];
    float mean = 0.f, mn = 10000.0f;
    std::vector<float> z(k);
    for (int i = 0; i < k; i++) {
      z[i] = pc[(size_t)nbr[(size_t)g * k + i] * 3 + 2];
      mean += z[i];
      if (z[i] < mn) mn = z[i];
    }
    mean = mean / k;
    float dZ = -100000.0f, vZ = 0.f;
    for (int i = 0; i < k; i++) {
      float diff = z[i] - mn;
      vZ = (float)((double)vZ + std::pow((double)fabsf(z[i] - mean), 2.0));
      if (diff > dZ) dZ = diff;
    }
    vZ = vZ / k;
    float* f = feature + (size_t)g * featsize;
    f[0] = C_; f[1] = O_; f[2] = L_; f[3] = E_; f[4] = P_; f[5] = S_; f[6] = A_; f[7] = X_; f[8] = D_; f[9] = S_2; f[10] = L_2;
    f[11] = dZ; f[12] = vZ;
  }
}

// GPUTransformer (feat): point SoA float[featsize*size] (channel j at point[i + j*size]; channels 0..2 are x,y,z);
// out float[num_x*num_y*featsize] (caller-zeroed).  The reference update is a racy compare-then-write against a running
// maximum initialised to 0; its intended (and race-free) result is the per-cell, per-channel maximum of the positive
// values -- that is what is computed here (num_height == 1, the only configuration RING++ uses, RING/config.py:9).
int orc_voxelfeat_bev(const float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_height, int featsize,
                      float* out, int* n_dropped) {
  float gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
  float gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
  float gap_h = (float)(2.0 * (double)(float)max_height / (double)(float)num_height);
  int dropped = 0;
  for (int i = 0; i < size; i++) {
    float x = point[i], y = point[i + size], z = point[i + 2 * (size_t)size];
    if (x == 0.0) x = 0.0001;
    if (y == 0.0) y = 0.0001;
    if (z == 0.0) z = 0.0001;
    if (x > 1.0) x = 0.9999;
    if (y > 1.0) y = 0.9999;
    if (z > 1.0) z = 0.9999;
    if (x < -1.0) x = -0.9999;
    if (y < -1.0) y = -0.9999;
    if (z < -1.0) z = -0.9999;
    int ix = (int)std::floor(((double)x + 1.0) / (double)gap_x);
    int iy = (int)std::floor(((double)y + 1.0) / (double)gap_y);
    int ih = (int)std::floor(((double)z + 1.0) / (double)gap_h);
    if (ix < 0 || ix >= num_x || iy < 0 || iy >= num_y || ih != 0) {
      dropped++;
      continue;
    }
    for (int j = 0; j < featsize; j++) {
      float v = point[i + (size_t)j * size];
      float& o = out[(size_t)featsize * (iy + ix * num_y) + j];
      if (o < v) o = v;  // running maximum starting at 0: only positive values are ever written
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return 0;
}
}
