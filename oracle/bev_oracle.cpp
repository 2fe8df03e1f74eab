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
