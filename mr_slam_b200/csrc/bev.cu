// bev.cu -- RING Cartesian voxeliser (`voxelocc`) and DiSCO polar voxeliser (`gputransform`) for sm_100a.
//
// Reference (paths relative to LoopDetection/):
//   generate_bev_cython_binary/src/kernel.cu:14-61      point2gridmap   (index math in IEEE double)
//   generate_bev_cython_binary/src/manager.cu:53-92     retreive(): sequential HOST scatter in point order
//   src/disco_ros/tools/multi-layer-polar-gpu/cython/src/kernel.cu:24-80, manager.cu:53-81
// The reference computes indices on the GPU, copies them back and scatters on the CPU in a serial loop.
// Here the scatter runs on the GPU with order-independent formulations of the same result:
//   * channels 0/1 (x,y of the LAST point written to a voxel)        -> atomicMax on the point index, then gather
//   * channel 2, num_height == 1 (max z over the column, if > 0)     -> atomicMax on the int pattern of positive floats
//   * polar, enough_large == 1 (x,y of the FIRST point of a voxel)   -> atomicMin on the point index
//   * general num_height > 1 / enough_large > 1: stable radix sort by column/voxel key, then one thread walks
//     each run in point order (exactly the reference's sequential semantics).
// Points whose voxel index falls outside the grid (x == 1.0 exactly, |z| > max_height for the polar variant)
// make the reference write out of bounds; they are skipped here (documented in DESIGN.md).

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"

namespace mrslam {

struct VoxCfg {
  int size;
  float gap_x, gap_y, gap_h;
  int num_x, num_y, num_h, enough_large;
  float max_height;  // polar only
};

__device__ __forceinline__ bool cart_index_xyz(float x, float y, float z, const VoxCfg& c, int& ix, int& iy, int& ih);
__device__ __forceinline__ bool cart_index(const float* __restrict__ pt, int i, const VoxCfg& c, int& ix, int& iy, int& ih) {
  return cart_index_xyz(pt[i], pt[i + c.size], pt[i + 2 * (size_t)c.size], c, ix, iy, ih);
}
__device__ __forceinline__ bool cart_index_xyz(float x, float y, float z, const VoxCfg& c, int& ix, int& iy, int& ih) {
  if (x == 0.0f) x = 0.0001f;
  if (y == 0.0f) y = 0.0001f;
  if (z == 0.0f) z = 0.0001f;
  if (x > 1.0f) x = 0.9999f;
  if (y > 1.0f) y = 0.9999f;
  if (z > 1.0f) z = 0.9999f;
  if (x < -1.0f) x = -0.9999f;
  if (y < -1.0f) y = -0.9999f;
  if (z < -1.0f) z = -0.9999f;
  ix = (int)floor(((double)x + 1.0) / (double)c.gap_x);
  iy = (int)floor(((double)y + 1.0) / (double)c.gap_y);
  ih = (int)floor(((double)z + 1.0) / (double)c.gap_h);
  return ix >= 0 && ix < c.num_x && iy >= 0 && iy < c.num_y && ih >= 0 && ih < c.num_h;
}

// theta exactly as the reference GPU kernel evaluates it: float atan, double scaling with the FMA nvcc contracts
__device__ __forceinline__ float polar_theta(float x, float y) {
  const double k = 180 / 3.14159265358979323846;
  if (x >= 0 && y >= 0) return (float)(k * (double)atanf(y / x));
  else if (x < 0 && y >= 0) return (float)fma((double)atanf(y / (-x)), -k, 180.0);
  else if (x < 0 && y < 0) return (float)fma((double)atanf(y / x), k, 180.0);
  else return (float)fma((double)atanf((-y) / x), -k, 360.0);
}

__device__ __forceinline__ bool polar_index(const float* __restrict__ pt, int i, const VoxCfg& c, int& ir, int& is, int& ih) {
  float x = pt[i], y = pt[i + c.size], z = pt[i + 2 * (size_t)c.size];
  if (x == 0.0f) x = 0.0001f;
  if (y == 0.0f) y = 0.0001f;
  if (z == 0.0f) z = 0.0001f;
  const float theta = polar_theta(x, y);
  const float faraway = (float)sqrt((double)x * (double)x + (double)y * (double)y);
  ir = (int)floorf(faraway / c.gap_x);
  is = (int)floorf(theta / c.gap_y);
  ih = (int)floorf((z + c.max_height) / c.gap_h);
  if (ir >= c.num_x) ir = c.num_x - 1;
  return ir >= 0 && is >= 0 && is < c.num_y && ih >= 0 && ih < c.num_h;
}

// ---- Cartesian, num_height == 1 ------------------------------------------------------------------------
__global__ void k_vox_cart_scatter(const float* __restrict__ pt, VoxCfg c, int* __restrict__ last_idx, int* __restrict__ zmax_bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  int ix, iy, ih;
  if (!cart_index(pt, i, c, ix, iy, ih)) return;
  const int cell = iy + ix * c.num_y;  // ih == 0
  atomicMax(&last_idx[cell], i);
  const float z = pt[i + 2 * (size_t)c.size];  // ORIGINAL z is compared and stored (manager.cu:71-73)
  if (z > 0.0f) atomicMax(&zmax_bits[cell], __float_as_int(z));
}
__global__ void k_vox_cart_gather(const float* __restrict__ pt, VoxCfg c, const int* __restrict__ last_idx,
                                  const int* __restrict__ zmax_bits, float* __restrict__ out) {
  int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= c.num_x * c.num_y) return;
  const int li = last_idx[cell];
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  if (li >= 0) {
    o0 = pt[li];
    o1 = pt[li + c.size];
  }
  const int zb = zmax_bits[cell];
  if (zb > 0) o2 = __int_as_float(zb);
  out[3 * (size_t)cell + 0] = o0;
  out[3 * (size_t)cell + 1] = o1;
  out[3 * (size_t)cell + 2] = o2;
}

// batched device variant for the RING pipeline: only channel 2 (the max-height image) is produced
__global__ void k_vox_cart_bev_batch(const float* __restrict__ pts, const long long* __restrict__ offsets, float gap_x, float gap_y,
                                     float gap_h, int num_x, int num_y, int* __restrict__ bev_bits) {
  const int b = blockIdx.y;
  const long long o0 = offsets[b], o1 = offsets[b + 1];
  const int n = (int)(o1 - o0);
  VoxCfg c;
  c.size = n; c.gap_x = gap_x; c.gap_y = gap_y; c.gap_h = gap_h; c.num_x = num_x; c.num_y = num_y; c.num_h = 1; c.enough_large = 1;
  const float* pt = pts + 3 * o0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int ix, iy, ih;
    if (!cart_index(pt, i, c, ix, iy, ih)) continue;
    const float z = pt[i + 2 * (size_t)n];
    if (z > 0.0f) atomicMax(&bev_bits[(size_t)b * num_x * num_y + iy + ix * num_y], __float_as_int(z));
  }
}

// RING front end in one pass over the RAW scans (AoS rows of `stride` floats, x y z first): load_pc_infer
// (LoopDetection/src/RING_ros/util.py:91-112 -- keep |x|,|y| < xy_range, 0 < z < z_range, divide by the ranges) fused with the
// channel-2 voxeliser above.  The compaction the reference does (`pc[conditions]`) only changes the point order, which the
// max-height image does not depend on.
__global__ void k_ring_scan_bev_batch(const float* __restrict__ pts, const long long* __restrict__ offsets, int stride, float xy_range,
                                      float z_range, float gap_x, float gap_y, float gap_h, int num_x, int num_y, int* __restrict__ bev_bits) {
  const int b = blockIdx.y;
  const long long o0 = offsets[b], o1 = offsets[b + 1];
  const int n = (int)(o1 - o0);
  VoxCfg c;
  c.size = n; c.gap_x = gap_x; c.gap_y = gap_y; c.gap_h = gap_h; c.num_x = num_x; c.num_y = num_y; c.num_h = 1; c.enough_large = 1;
  const float* pt = pts + (size_t)o0 * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = pt + (size_t)i * stride;
    const float x = p[0], y = p[1], z = p[2];
    if (!(fabsf(x) < xy_range && fabsf(y) < xy_range && z < z_range && z > 0.0f)) continue;
    const float xs = x / xy_range, ys = y / xy_range, zs = z / z_range;
    int ix, iy, ih;
    if (!cart_index_xyz(xs, ys, zs, c, ix, iy, ih)) continue;
    if (zs > 0.0f) atomicMax(&bev_bits[(size_t)b * num_x * num_y + iy + ix * num_y], __float_as_int(zs));
  }
}

// ---- general (sorted) paths ----------------------------------------------------------------------------
__global__ void k_vox_keys(const float* __restrict__ pt, VoxCfg c, int polar, int* __restrict__ keys, int* __restrict__ vals,
                           int* __restrict__ layer) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  int a, b, h;
  bool ok = polar ? polar_index(pt, i, c, a, b, h) : cart_index(pt, i, c, a, b, h);
  const int ncol = c.num_x * c.num_y;
  // Cartesian: runs are COLUMNS (the running max is shared by all layers of a column); polar: runs are voxels
  int key = polar ? (b + a * c.num_y + h * ncol) : (b + a * c.num_y);
  keys[i] = ok ? key : (polar ? ncol * c.num_h : ncol);
  vals[i] = i;
  layer[i] = h;
}
__global__ void k_vox_cart_runs(const float* __restrict__ pt, VoxCfg c, const int* __restrict__ keys, const int* __restrict__ order,
                                const int* __restrict__ layer, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  const int ncol = c.num_x * c.num_y;
  const int col = keys[i];
  if (col >= ncol || (i > 0 && keys[i - 1] == col)) return;
  double maxh = 0.0;
  for (int j = i; j < c.size && keys[j] == col; j++) {
    const int p = order[j];
    const size_t slot = 3 * ((size_t)col + (size_t)layer[p] * ncol);
    out[slot + 0] = pt[p];
    out[slot + 1] = pt[p + c.size];
    const float z = pt[p + 2 * (size_t)c.size];
    if (maxh < (double)z) {
      out[slot + 2] = z;
      maxh = (double)z;
    }
  }
}
__global__ void k_vox_polar_runs(const float* __restrict__ pt, VoxCfg c, const int* __restrict__ keys, const int* __restrict__ order,
                                 float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  const int grid = c.num_x * c.num_y * c.num_h;
  const int vox = keys[i];
  if (vox >= grid || (i > 0 && keys[i - 1] == vox)) return;
  int cnt = 0;
  for (int j = i; j < c.size && keys[j] == vox && cnt < c.enough_large; j++, cnt++) {
    const int p = order[j];
    const size_t slot = 3 * ((size_t)vox + (size_t)cnt * grid);
    out[slot + 0] = pt[p];
    out[slot + 1] = pt[p + c.size];
    out[slot + 2] = 1.f;
  }
}
// polar fast path, enough_large == 1
__global__ void k_vox_polar_scatter(const float* __restrict__ pt, VoxCfg c, int* __restrict__ first_idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  int ir, is, ih;
  if (!polar_index(pt, i, c, ir, is, ih)) return;
  atomicMin(&first_idx[is + ir * c.num_y + ih * c.num_y * c.num_x], i);
}
__global__ void k_vox_polar_gather(const float* __restrict__ pt, VoxCfg c, const int* __restrict__ first_idx, float* __restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= c.num_x * c.num_y * c.num_h) return;
  const int fi = first_idx[v];
  float o0 = 0.f, o1 = 0.f, o2 = 0.f;
  if (fi != INT_MAX) {
    o0 = pt[fi];
    o1 = pt[fi + c.size];
    o2 = 1.f;
  }
  out[3 * (size_t)v] = o0;
  out[3 * (size_t)v + 1] = o1;
  out[3 * (size_t)v + 2] = o2;
}
__global__ void k_fill_i32(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- voxelfeat (RING++): generate_bev_pointfeat_cython/src/kernel.cu:16-104 (features) and :106-164 (feature BEV) ----
__global__ void k_voxelfeat_features(const float* __restrict__ pc, int size, int featsize, int k, const int* __restrict__ nbr,
                                     const float* __restrict__ eig, float* __restrict__ feature) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= size) return;
  const float eps = 0.0f;
  float e3[3], e2[2];
  float s3 = 0.f, p3 = 1.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (i < 3) {
      e3[i] = eig[(size_t)g * 5 + i];
      s3 += e3[i];
      p3 = p3 * e3[i];
    } else {
      e2[i - 3] = eig[(size_t)g * 5 + i];
      s2 += e2[i - 3];
    }
  }
  const float C_ = e3[2] / (s3 + eps);
  const float O_ = (float)pow((double)p3 / (pow((double)s3, 3.0) + (double)eps), 1.0 / 3.0);
  const float L_ = (e3[0] - e3[1]) / (e3[0] + eps);
  float E_ = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) E_ += (e3[i] / (s3 + eps)) * logf(e3[i] / (s3 + eps));
  E_ = -E_;
  const float P_ = (e3[1] - e3[2]) / e3[0];
  const float S_ = e3[2] / e3[0];
  const float A_ = (e3[0] - e3[2]) / e3[0];
  const float D_ = (float)((double)(3 * k) / ((4 * 3.14159265358979323846 * (double)p3) + (double)eps));
  const float L_2 = e2[1] / e2[0];
  float mean = 0.f, mn = 10000.0f;
  const int* nb = nbr + (size_t)g * k;
  for (int i = 0; i < k; i++) {
    const float z = pc[(size_t)nb[i] * 3 + 2];
    mean += z;
    if (z < mn) mn = z;
  }
  mean = mean / k;
  float dZ = -100000.0f, vZ = 0.f;
  for (int i = 0; i < k; i++) {
    const float z = pc[(size_t)nb[i] * 3 + 2];
    const float diff = z - mn;
    vZ = (float)((double)vZ + pow((double)fabsf(z - mean), 2.0));
    if (diff > dZ) dZ = diff;
  }
  vZ = vZ / k;
  float* f = feature + (size_t)g * featsize;
  f[0] = C_; f[1] = O_; f[2] = L_; f[3] = E_; f[4] = P_; f[5] = S_; f[6] = A_; f[7] = s3; f[8] = D_; f[9] = s2; f[10] = L_2;
  f[11] = dZ; f[12] = vZ;
}

// per-cell, per-channel maximum of the positive values (the race-free result the reference's compare-then-write intends)
__global__ void k_voxelfeat_bev(const float* __restrict__ pt, VoxCfg c, int featsize, int* __restrict__ out_bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.size) return;
  int ix, iy, ih;
  if (!cart_index(pt, i, c, ix, iy, ih) || ih != 0) return;
  int* o = out_bits + (size_t)featsize * (iy + ix * c.num_y);
  for (int j = 0; j < featsize; j++) {
    const float v = pt[i + (size_t)j * c.size];
    if (v > 0.0f) atomicMax(&o[j], __float_as_int(v));
  }
}

struct BevScratch {
  DevBuf<float> pts, out;
  DevBuf<int> a, b, keys, keys_alt, vals, vals_alt, layer;
  RadixScratch rsort;
  DevBuf<long long> offsets;
};
static BevScratch& bs() {
  static BevScratch s;
  return s;
}

static int sort_pairs(BevScratch& S, int n, int maxkey, const int** keys_out, const int** vals_out) {
  int bits = 1;
  while ((1ll << bits) <= maxkey) bits++;
  // the library's own stable radix sort (csrc/radix_sort.cuh)
  unsigned *ks = nullptr, *vs = nullptr;
  const int* d_n = nullptr;
  MR_CHECK(radix_sort_pairs(S.rsort, (unsigned*)S.keys.p, (unsigned*)S.vals.p, (unsigned*)S.keys_alt.p, (unsigned*)S.vals_alt.p, n, nullptr, 0,
                            ((bits + 7) / 8) * 8, 0u, 0, rt().stream, &ks, &vs, &d_n));
  *keys_out = (const int*)ks;
  *vals_out = (const int*)vs;
  return MRSLAM_OK;
}

static int transform_host(bool polar, const float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_h,
                          int enough_large, float* out) {
  MR_CHECK(ensure_init());
  if (size < 0 || num_x <= 0 || num_y <= 0 || num_h <= 0 || enough_large <= 0 || !out || (size > 0 && !point)) {
    set_error("GPUTransformer: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  BevScratch& S = bs();
  VoxCfg c;
  c.size = size;
  c.num_x = num_x; c.num_y = num_y; c.num_h = num_h; c.enough_large = enough_large;
  c.max_height = (float)max_height;
  if (!polar) {
    c.gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
    c.gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
    c.gap_h = (float)(2.0 * (double)(float)max_height / (double)(float)num_h);
  } else {
    c.gap_x = (float)max_length / (float)num_x;
    c.gap_y = (float)(360.0 / (double)(float)num_y);
    c.gap_h = (float)(2.0 * (double)(float)max_height / (double)(float)num_h);
  }
  const size_t grid = (size_t)num_x * num_y * num_h;
  const size_t out_n = grid * enough_large * 3;
  MR_CHECK(S.pts.reserve(std::max<size_t>((size_t)size * 3, 1)));
  MR_CHECK(S.out.reserve(out_n));
  if (size > 0) MR_CUDA(cudaMemcpyAsync(S.pts.p, point, sizeof(float) * 3 * (size_t)size, cudaMemcpyHostToDevice, st));
  const int nb = (size + 255) / 256;
  const bool fast = polar ? (enough_large == 1) : (num_h == 1);
  if (fast) {
    MR_CHECK(S.a.reserve(grid));
    MR_CHECK(S.b.reserve(grid));
    if (!polar) {
      k_fill_i32<<<((int)grid + 255) / 256, 256, 0, st>>>(S.a.p, (int)grid, -1);
      MR_CUDA(cudaMemsetAsync(S.b.p, 0, sizeof(int) * grid, st));
      if (size > 0) k_vox_cart_scatter<<<nb, 256, 0, st>>>(S.pts.p, c, S.a.p, S.b.p);
      k_vox_cart_gather<<<((int)grid + 255) / 256, 256, 0, st>>>(S.pts.p, c, S.a.p, S.b.p, S.out.p);
      count_launch(size > 0 ? 3 : 2);
      // slots beyond the first enough_large layer stay zero like the reference's np.zeros buffer
      if (enough_large > 1) MR_CUDA(cudaMemsetAsync(S.out.p + grid * 3, 0, sizeof(float) * (out_n - grid * 3), st));
    } else {
      k_fill_i32<<<((int)grid + 255) / 256, 256, 0, st>>>(S.a.p, (int)grid, INT_MAX);
      if (size > 0) k_vox_polar_scatter<<<nb, 256, 0, st>>>(S.pts.p, c, S.a.p);
      k_vox_polar_gather<<<((int)grid + 255) / 256, 256, 0, st>>>(S.pts.p, c, S.a.p, S.out.p);
      count_launch(size > 0 ? 3 : 2);
    }
  } else {
    MR_CUDA(cudaMemsetAsync(S.out.p, 0, sizeof(float) * out_n, st));
    if (size > 0) {
      size_t m = size;
      MR_CHECK(S.keys.reserve(m)); MR_CHECK(S.keys_alt.reserve(m)); MR_CHECK(S.vals.reserve(m)); MR_CHECK(S.vals_alt.reserve(m));
      MR_CHECK(S.layer.reserve(m));
      k_vox_keys<<<nb, 256, 0, st>>>(S.pts.p, c, polar ? 1 : 0, S.keys.p, S.vals.p, S.layer.p);
      const int *sk, *sv;
      MR_CHECK(sort_pairs(S, size, polar ? (int)grid : num_x * num_y, &sk, &sv));
      if (polar) k_vox_polar_runs<<<nb, 256, 0, st>>>(S.pts.p, c, sk, sv, S.out.p);
      else k_vox_cart_runs<<<nb, 256, 0, st>>>(S.pts.p, c, sk, sv, S.layer.p, S.out.p);
      count_launch(2);
    }
  }
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaMemcpyAsync(out, S.out.p, sizeof(float) * out_n, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}

}  // namespace mrslam

using namespace mrslam;

extern "C" {

int mrslam_voxelocc_transform(const float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_height,
                              int enough_large, float* out) {
  return transform_host(false, point, size, max_length, max_height, num_x, num_y, num_height, enough_large, out);
}
int mrslam_polar_transform(const float* point, int size, int max_length, int max_height, int num_ring, int num_sector, int num_height,
                           int enough_large, float* out) {
  return transform_host(true, point, size, max_length, max_height, num_ring, num_sector, num_height, enough_large, out);
}

int mrslam_voxelfeat_features(const float* point_aos, int size, int featsize, int k, const int* neighbors, const float* eigens,
                              float* feature_out) {
  MR_CHECK(ensure_init());
  if (size < 0 || featsize < 13 || k < 1 || k > 30 || !feature_out) {
    set_error("GPUFeatureExtractor: bad arguments (featsize >= 13, 1 <= k <= 30)");
    return MRSLAM_ERR_ARG;
  }
  if (size == 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  BevScratch& S = bs();
  MR_CHECK(S.pts.reserve((size_t)size * 3));
  MR_CHECK(S.out.reserve((size_t)size * featsize));
  MR_CHECK(S.keys.reserve((size_t)size * k));
  DevBuf<float> deig;
  MR_CHECK(deig.reserve((size_t)size * 5));
  MR_CUDA(cudaMemcpyAsync(S.pts.p, point_aos, sizeof(float) * 3 * (size_t)size, cudaMemcpyHostToDevice, st));
  MR_CUDA(cudaMemcpyAsync(S.keys.p, neighbors, sizeof(int) * (size_t)size * k, cudaMemcpyHostToDevice, st));
  MR_CUDA(cudaMemcpyAsync(deig.p, eigens, sizeof(float) * 5 * (size_t)size, cudaMemcpyHostToDevice, st));
  MR_CUDA(cudaMemsetAsync(S.out.p, 0, sizeof(float) * (size_t)size * featsize, st));
  k_voxelfeat_features<<<(size + 255) / 256, 256, 0, st>>>(S.pts.p, size, featsize, k, S.keys.p, deig.p, S.out.p);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaMemcpyAsync(feature_out, S.out.p, sizeof(float) * (size_t)size * featsize, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  deig.release();
  return MRSLAM_OK;
}

int mrslam_voxelfeat_transform(const float* point, int size, int max_length, int max_height, int num_x, int num_y, int num_height,
                               int featsize, float* out) {
  MR_CHECK(ensure_init());
  if (size < 0 || num_x <= 0 || num_y <= 0 || num_height != 1 || featsize < 3 || !out) {
    set_error("voxelfeat GPUTransformer: bad arguments (num_height must be 1, the only configuration RING++ uses)");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  BevScratch& S = bs();
  VoxCfg c;
  c.size = size; c.num_x = num_x; c.num_y = num_y; c.num_h = 1; c.enough_large = 1; c.max_height = (float)max_height;
  c.gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
  c.gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
  c.gap_h = (float)(2.0 * (double)(float)max_height / 1.0);
  const size_t out_n = (size_t)num_x * num_y * featsize;
  MR_CHECK(S.pts.reserve(std::max<size_t>((size_t)size * featsize, 1)));
  MR_CHECK(S.out.reserve(out_n));
  MR_CUDA(cudaMemsetAsync(S.out.p, 0, sizeof(float) * out_n, st));
  if (size > 0) {
    MR_CUDA(cudaMemcpyAsync(S.pts.p, point, sizeof(float) * (size_t)size * featsize, cudaMemcpyHostToDevice, st));
    k_voxelfeat_bev<<<(size + 255) / 256, 256, 0, st>>>(S.pts.p, c, featsize, (int*)S.out.p);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaMemcpyAsync(out, S.out.p, sizeof(float) * out_n, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}

int mrslam_voxelocc_bev_device(const float* d_points_soa, const long long* offsets_host, int batch, int max_length, int max_height,
                               int num_x, int num_y, float* d_bev) {
  MR_CHECK(ensure_init());
  if (batch <= 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  BevScratch& S = bs();
  MR_CHECK(S.offsets.reserve(batch + 1));
  MR_CUDA(cudaMemcpyAsync(S.offsets.p, offsets_host, sizeof(long long) * (batch + 1), cudaMemcpyHostToDevice, st));
  long long maxn = 0;
  for (int b = 0; b < batch; b++) maxn = std::max(maxn, offsets_host[b + 1] - offsets_host[b]);
  const float gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
  const float gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
  const float gap_h = (float)(2.0 * (double)(float)max_height / 1.0);
  MR_CUDA(cudaMemsetAsync(d_bev, 0, sizeof(float) * (size_t)batch * num_x * num_y, st));
  if (maxn > 0) {
    int gx = (int)std::min<long long>((maxn + 255) / 256, 64);
    k_vox_cart_bev_batch<<<dim3(gx, batch), 256, 0, st>>>(d_points_soa, S.offsets.p, gap_x, gap_y, gap_h, num_x, num_y, (int*)d_bev);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

int mrslam_ring_scans_to_bev_device(const float* d_points, const long long* offsets_host, int batch, int stride, float xy_range,
                                    float z_range, int max_length, int max_height, int num_x, int num_y, float* d_bev) {
  MR_CHECK(ensure_init());
  if (batch < 0 || stride < 3 || num_x <= 0 || num_y <= 0 || !(xy_range > 0.f) || !(z_range > 0.f) || (batch > 0 && (!offsets_host || !d_bev))) {
    set_error("ring_scans_to_bev: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  if (batch == 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  BevScratch& S = bs();
  MR_CHECK(S.offsets.reserve(batch + 1));
  MR_CUDA(cudaMemcpyAsync(S.offsets.p, offsets_host, sizeof(long long) * (batch + 1), cudaMemcpyHostToDevice, st));
  long long maxn = 0;
  for (int b = 0; b < batch; b++) maxn = std::max(maxn, offsets_host[b + 1] - offsets_host[b]);
  if (maxn > 0 && !d_points) {
    set_error("ring_scans_to_bev: null points");
    return MRSLAM_ERR_ARG;
  }
  const float gap_x = (float)(2.0 * (double)(float)max_length / (double)(float)num_x);
  const float gap_y = (float)(2.0 * (double)(float)max_length / (double)(float)num_y);
  const float gap_h = (float)(2.0 * (double)(float)max_height / 1.0);
  MR_CUDA(cudaMemsetAsync(d_bev, 0, sizeof(float) * (size_t)batch * num_x * num_y, st));
  if (maxn > 0) {
    int gx = (int)std::min<long long>((maxn + 255) / 256, 64);
    k_ring_scan_bev_batch<<<dim3(gx, batch), 256, 0, st>>>(d_points, S.offsets.p, stride, xy_range, z_range, gap_x, gap_y, gap_h, num_x, num_y,
                                                          (int*)d_bev);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  // the pageable offsets array was read by an asynchronous copy: do not return before it has been consumed
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}
}
