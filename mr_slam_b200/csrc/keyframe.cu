// keyframe.cu -- the front of ICPCheck: GlobalManager::mergeNearestKeyframes,
// Mapping/src/global_manager/src/global_manager.cpp:1894-1939 -- the 2*submap_size+1 keyframes around a loop candidate are
// transformed into the candidate's frame (pcl::transformPointCloud), concatenated, cropped to |x|,|y| <= 60 m
// (pcl::PassThrough twice) and reduced by a pcl::VoxelGrid centroid filter (leaf icp_filter_size); the result is the GICP
// input.  The reference does this on the CPU for both clouds of every ICPCheck (~1 M points in, ~100 k out).
//
// PCL 1.10 is a third-party dependency that is absent from /root/reference; its published algorithms are restated:
//   common/impl/transforms.hpp  Transformer::se3     x' = m00 x + (m01 y + (m02 z + m03)) per row (SSE mul/add, no FMA)
//   filters/impl/passthrough.hpp                     drop non-finite points, keep min <= v <= max
//   filters/impl/voxel_grid.hpp applyFilter          min/max of the cloud -> min_b = floor(min * inv_leaf), div_b, linear
//                                                    index (i - min_b) . (1, dx, dx dy); sort by index; one centroid
//                                                    (float sums of x, y, z, intensity divided by the count) per index,
//                                                    emitted in ascending index; if dx dy dz overflows int the input is
//                                                    returned unfiltered with a warning.
// PCL sorts with std::sort (not stable), so the summation order inside a voxel is unspecified there; here it is ascending
// input order (stable radix sort), which the CPU oracle mirrors -> oracle and GPU agree bit for bit, PCL to float rounding.
//
// Kernels: k_kf_transform (all keyframes of a merge in one launch, blockIdx.y = keyframe) -> k_kf_minmax -> k_kf_keys ->
// CUB radix sort (library) -> k_kf_heads + CUB exclusive scan -> k_kf_centroids (one thread per voxel run).  HBM bound:
// 16 B/pt in, 16 B/pt staged, 8 B/pt per sort pass, 16 B per output voxel.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"

namespace mrslam {

struct KfDesc {
  const float* src;  // device pointer of this keyframe's first point
  int n;
  int offset;  // position in the concatenated cloud
  float T[12];
};

constexpr int KF_INVALID = 0x7fffffff;

__global__ void k_kf_transform(const KfDesc* __restrict__ descs, int stride, float lim, float4* __restrict__ cat, int* __restrict__ valid) {
  const KfDesc D = descs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.n; i += gridDim.x * blockDim.x) {
    const float* p = D.src + (size_t)i * stride;
    const float x = p[0], y = p[1], z = p[2], w = stride > 3 ? p[3] : 0.0f;
    const float* T = D.T;
    const float xo = T[0] * x + (T[1] * y + (T[2] * z + T[3]));
    const float yo = T[4] * x + (T[5] * y + (T[6] * z + T[7]));
    const float zo = T[8] * x + (T[9] * y + (T[10] * z + T[11]));
    cat[D.offset + i] = make_float4(xo, yo, zo, w);
    const bool fin = isfinite(xo) && isfinite(yo) && isfinite(zo);
    valid[D.offset + i] = (fin && !(xo > lim || xo < -lim) && !(yo > lim || yo < -lim)) ? 1 : 0;
  }
}

__global__ void k_kf_minmax_init(unsigned* mm) {
  if (threadIdx.x < 3) {
    mm[threadIdx.x] = f2ord(INFINITY);
    mm[3 + threadIdx.x] = f2ord(-INFINITY);
  }
  if (threadIdx.x == 6) mm[6] = 0;  // number of valid points
}

__global__ void __launch_bounds__(256) k_kf_minmax(const float4* __restrict__ cat, const int* __restrict__ valid, int n, unsigned* mm) {
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ int s_cnt[8];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (valid[i]) {
      const float4 v = cat[i];
      mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
      mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
      cnt++;
    }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = warp_min(mn[a]);
    mx[a] = warp_max(mx[a]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    for (int a = 0; a < 3; a++) {
      s_mn[w][a] = mn[a];
      s_mx[w][a] = mx[a];
    }
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    float lo = s_mn[0][a], hi = s_mx[0][a];
    for (int k = 1; k < 8; k++) {
      lo = fminf(lo, s_mn[k][a]);
      hi = fmaxf(hi, s_mx[k][a]);
    }
    atomicMin(&mm[a], f2ord(lo));
    atomicMax(&mm[3 + a], f2ord(hi));
  }
  if (threadIdx.x == 3) {
    int c = 0;
    for (int k = 0; k < 8; k++) c += s_cnt[k];
    atomicAdd(&mm[6], (unsigned)c);
  }
}

struct KfGrid {
  float inv[3];
  int min_b[3];
  int mul[3];
};

__global__ void k_kf_keys(const float4* __restrict__ cat, const int* __restrict__ valid, int n, KfGrid g, int* __restrict__ keys,
                          int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int key = KF_INVALID;
  if (valid[i]) {
    const float4 v = cat[i];
    const int i0 = (int)(floorf(v.x * g.inv[0]) - (float)g.min_b[0]);
    const int i1 = (int)(floorf(v.y * g.inv[1]) - (float)g.min_b[1]);
    const int i2 = (int)(floorf(v.z * g.inv[2]) - (float)g.min_b[2]);
    key = i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2];
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void k_kf_heads(const int* __restrict__ keys, int n, int* __restrict__ head) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = keys[i];
  head[i] = (k != KF_INVALID && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// one thread per voxel run: float sums in ascending input order (stable sort), divided by the count
__global__ void k_kf_centroids(const int* __restrict__ keys, const int* __restrict__ order, const int* __restrict__ head,
                               const int* __restrict__ pos, int n, const float4* __restrict__ cat, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  const int k = keys[i];
  float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
  int cnt = 0;
  for (int j = i; j < n && keys[j] == k; j++) {
    const float4 v = cat[order[j]];
    sx += v.x; sy += v.y; sz += v.z; sw += v.w;
    cnt++;
  }
  const float c = (float)cnt;
  out[pos[i]] = make_float4(sx / c, sy / c, sz / c, sw / c);
}

// unfiltered fall-back (PCL's "leaf size too small" path): the valid points in input order
__global__ void k_kf_compact(const float4* __restrict__ cat, const int* __restrict__ valid, const int* __restrict__ pos, int n,
                             float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && valid[i]) out[pos[i]] = cat[i];
}

// ---- wire formats at the seams ---------------------------------------------------------------------
// sensor_msgs/PointCloud2 blob (dislam_msgs/SubMap.submap / .keyframePC) -> packed x,y,z,intensity on the device; replaces the
// per-point python loop over pc2.read_points(..., skip_nans=True) of LoopDetection/src/RING_ros/main_RING.py:251-254.
// Fields are little-endian float32 at arbitrary byte offsets inside a point_step-byte record (read bytewise: no alignment).
__device__ __forceinline__ float load_f32_unaligned(const unsigned char* p) {
  const unsigned u = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
  return __uint_as_float(u);
}
__global__ void k_pc2_decode(const unsigned char* __restrict__ data, int n, int point_step, int ox, int oy, int oz, int oi, int skip_nans,
                             float4* __restrict__ pts, int* __restrict__ keep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* r = data + (size_t)i * point_step;
  const float x = load_f32_unaligned(r + ox), y = load_f32_unaligned(r + oy), z = load_f32_unaligned(r + oz);
  const float w = oi >= 0 ? load_f32_unaligned(r + oi) : 0.0f;
  pts[i] = make_float4(x, y, z, w);
  // read_points(skip_nans=True) drops a point when ANY requested field is NaN (sensor_msgs/point_cloud2.py:82-88)
  keep[i] = (skip_nans && (isnan(x) || isnan(y) || isnan(z) || (oi >= 0 && isnan(w)))) ? 0 : 1;
}

struct KfScratch {
  DevBuf<KfDesc> d_desc;
  PinBuf<KfDesc> h_desc;
  StageFence fence;
  DevBuf<float4> cat;
  DevBuf<int> valid, keys, keys_alt, vals, vals_alt, head, pos;
  DevBuf<unsigned> mm;
  DevBuf<unsigned char> tmp;
  RadixScratch rsort;
  DevBuf<float> staging;  // host inputs
  DevBuf<unsigned char> bytes;
};
static KfScratch& ks() {
  static KfScratch s;
  return s;
}

}  // namespace mrslam

using namespace mrslam;

extern "C" {

int mrslam_keyframe_merge(const float* const* clouds, const int* n, int stride_floats, int count, const float* T16, float xy_limit,
                          float leaf, int on_device, float* d_out_xyzi, int capacity, int* n_out) {
  MR_CHECK(ensure_init());
  if (count < 0 || (count > 0 && (!clouds || !n || !T16)) || stride_floats < 3 || !(leaf > 0.f) || !n_out || (capacity > 0 && !d_out_xyzi)) {
    set_error("keyframe_merge: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  KfScratch& S = ks();
  *n_out = 0;
  long long total = 0;
  int max_n = 0;
  for (int i = 0; i < count; i++) {
    if (n[i] < 0 || (n[i] > 0 && !clouds[i])) {
      set_error("keyframe_merge: bad cloud %d", i);
      return MRSLAM_ERR_ARG;
    }
    total += n[i];
    max_n = std::max(max_n, n[i]);
  }
  if (total == 0) return MRSLAM_OK;
  if (total > INT_MAX / 2) {
    set_error("keyframe_merge: %lld points do not fit", total);
    return MRSLAM_ERR_ARG;
  }
  const int N = (int)total;
  MR_CHECK(S.fence.wait());
  MR_CHECK(S.h_desc.reserve(count));
  MR_CHECK(S.d_desc.reserve(count));
  if (!on_device) MR_CHECK(S.staging.reserve((size_t)N * stride_floats));
  int off = 0;
  for (int i = 0; i < count; i++) {
    KfDesc d;
    d.n = n[i];
    d.offset = off;
    for (int r = 0; r < 12; r++) d.T[r] = T16[(size_t)i * 16 + r];
    if (on_device) {
      d.src = clouds[i];
    } else {
      d.src = S.staging.p + (size_t)off * stride_floats;
      if (n[i] > 0)
        MR_CUDA(cudaMemcpyAsync(S.staging.p + (size_t)off * stride_floats, clouds[i], ((size_t)(n[i] - 1) * stride_floats + std::min(stride_floats, 4)) * sizeof(float),
                                cudaMemcpyHostToDevice, st));
    }
    S.h_desc.p[i] = d;
    off += n[i];
  }
  MR_CUDA(cudaMemcpyAsync(S.d_desc.p, S.h_desc.p, sizeof(KfDesc) * count, cudaMemcpyHostToDevice, st));
  MR_CHECK(S.fence.mark(st));
  MR_CHECK(S.cat.reserve(N)); MR_CHECK(S.valid.reserve(N)); MR_CHECK(S.keys.reserve(N)); MR_CHECK(S.keys_alt.reserve(N));
  MR_CHECK(S.vals.reserve(N)); MR_CHECK(S.vals_alt.reserve(N)); MR_CHECK(S.head.reserve(N)); MR_CHECK(S.pos.reserve(N));
  MR_CHECK(S.mm.reserve(8));
  const int gx = std::max(1, std::min((max_n + 255) / 256, 256));
  k_kf_transform<<<dim3(gx, count), 256, 0, st>>>(S.d_desc.p, stride_floats, xy_limit, S.cat.p, S.valid.p);
  k_kf_minmax_init<<<1, 32, 0, st>>>(S.mm.p);
  k_kf_minmax<<<std::max(1, std::min((N + 255) / 256, 592)), 256, 0, st>>>(S.cat.p, S.valid.p, N, S.mm.p);
  count_launch(3);
  unsigned h_mm[7];
  MR_CUDA(cudaMemcpyAsync(h_mm, S.mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  const int n_valid = (int)h_mm[6];
  if (n_valid == 0) return MRSLAM_OK;
  // voxel_grid.hpp: bounds, overflow check, divisions
  float mn[3], mx[3];
  auto ord_to_float = [](unsigned u) {  // host twin of ord2f (common.cuh)
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &v, sizeof(f));
    return f;
  };
  for (int a = 0; a < 3; a++) {
    mn[a] = ord_to_float(h_mm[a]);
    mx[a] = ord_to_float(h_mm[3 + a]);
  }
  const float inv = 1.0f / leaf;
  long long dxyz[3];
  for (int a = 0; a < 3; a++) dxyz[a] = (long long)((mx[a] - mn[a]) * inv) + 1;
  const bool overflow = (dxyz[0] * dxyz[1] * dxyz[2]) > (long long)INT_MAX;
  size_t tmp_bytes = 0;
  int result = 0;
  if (overflow) {
    // PCL: "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
    MR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, S.valid.p, S.pos.p, N, st));
    MR_CHECK(S.tmp.reserve(tmp_bytes));
    MR_CUDA(cub::DeviceScan::ExclusiveSum(S.tmp.p, tmp_bytes, S.valid.p, S.pos.p, N, st));
    result = n_valid;
    if (result > capacity) {
      set_error("keyframe_merge: output needs %d points, capacity %d", result, capacity);
      return MRSLAM_ERR_ARG;
    }
    k_kf_compact<<<(N + 255) / 256, 256, 0, st>>>(S.cat.p, S.valid.p, S.pos.p, N, reinterpret_cast<float4*>(d_out_xyzi));
    count_launch(1);
    MR_CUDA(cudaGetLastError());
    *n_out = result;
    return MRSLAM_OK;
  }
  KfGrid g;
  int div_b[3];
  for (int a = 0; a < 3; a++) {
    g.inv[a] = inv;
    g.min_b[a] = (int)std::floor(mn[a] * inv);
    const int max_b = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b - g.min_b[a] + 1;
  }
  g.mul[0] = 1;
  g.mul[1] = div_b[0];
  g.mul[2] = div_b[0] * div_b[1];
  k_kf_keys<<<(N + 255) / 256, 256, 0, st>>>(S.cat.p, S.valid.p, N, g, S.keys.p, S.vals.p);
  count_launch(1);
  // stable sort by voxel key with the library's own radix sort (csrc/radix_sort.cuh); the head scan stays a library scan
  unsigned *ks = nullptr, *vs = nullptr;
  const int* d_ns = nullptr;
  MR_CHECK(radix_sort_pairs(S.rsort, (unsigned*)S.keys.p, (unsigned*)S.vals.p, (unsigned*)S.keys_alt.p, (unsigned*)S.vals_alt.p, N, nullptr, 0, 32, 0u, 0,
                            st, &ks, &vs, &d_ns));
  const int* keys_sorted = (const int*)ks;
  const int* vals_sorted = (const int*)vs;
  size_t scan_bytes = 0;
  MR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, S.head.p, S.pos.p, N, st));
  MR_CHECK(S.tmp.reserve(std::max(tmp_bytes, scan_bytes)));
  k_kf_heads<<<(N + 255) / 256, 256, 0, st>>>(keys_sorted, N, S.head.p);
  MR_CUDA(cub::DeviceScan::ExclusiveSum(S.tmp.p, scan_bytes, S.head.p, S.pos.p, N, st));
  count_launch(1);
  // number of voxels = pos[N-1] + head[N-1]
  int tail[2] = {0, 0};
  MR_CUDA(cudaMemcpyAsync(&tail[0], S.pos.p + (N - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(&tail[1], S.head.p + (N - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  result = tail[0] + tail[1];
  if (result > capacity) {
    set_error("keyframe_merge: output needs %d points, capacity %d", result, capacity);
    return MRSLAM_ERR_ARG;
  }
  k_kf_centroids<<<(N + 127) / 128, 128, 0, st>>>(keys_sorted, vals_sorted, S.head.p, S.pos.p, N, S.cat.p, reinterpret_cast<float4*>(d_out_xyzi));
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  *n_out = result;
  return MRSLAM_OK;
}
/* gtsam-style symbol keys of the Loop / SubMap messages: ((97 + robot) << 56) + keyframe index + 1
 * (global_manager.cpp:2587-2609 robotID2Key / Key2robotID, call sites :1949-1951; RING/util.py:253-260 robotid_to_key) */
unsigned long long mrslam_key_pack(int robot, long long keyframe_index) {
  return ((unsigned long long)(unsigned char)(97 + robot) << 56) + (unsigned long long)(keyframe_index + 1);
}
int mrslam_key_robot(unsigned long long key) { return (int)(char)(key >> 56) - 97; }
long long mrslam_key_index(unsigned long long key) { return (long long)(key - ((key >> 56) << 56)) - 1; }

int mrslam_pointcloud2_unpack(const unsigned char* data, int on_device, int n_points, int point_step, int off_x, int off_y, int off_z,
                              int off_intensity, int skip_nans, float* d_out_xyzi, int* n_out) {
  MR_CHECK(ensure_init());
  if (n_points < 0 || point_step < 12 || off_x < 0 || off_y < 0 || off_z < 0 || off_x + 4 > point_step || off_y + 4 > point_step ||
      off_z + 4 > point_step || off_intensity + 4 > point_step || !n_out || (n_points > 0 && (!data || !d_out_xyzi))) {
    set_error("pointcloud2_unpack: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  *n_out = 0;
  if (n_points == 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  KfScratch& S = ks();
  const unsigned char* src = data;
  if (!on_device) {
    MR_CHECK(S.bytes.reserve((size_t)n_points * point_step));
    MR_CUDA(cudaMemcpyAsync(S.bytes.p, data, (size_t)n_points * point_step, cudaMemcpyHostToDevice, st));
    src = S.bytes.p;
  }
  MR_CHECK(S.cat.reserve(n_points)); MR_CHECK(S.valid.reserve(n_points)); MR_CHECK(S.pos.reserve(n_points));
  k_pc2_decode<<<(n_points + 255) / 256, 256, 0, st>>>(src, n_points, point_step, off_x, off_y, off_z, off_intensity, skip_nans, S.cat.p,
                                                      S.valid.p);
  size_t tmp_bytes = 0;
  MR_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, S.valid.p, S.pos.p, n_points, st));
  MR_CHECK(S.tmp.reserve(tmp_bytes));
  MR_CUDA(cub::DeviceScan::ExclusiveSum(S.tmp.p, tmp_bytes, S.valid.p, S.pos.p, n_points, st));
  k_kf_compact<<<(n_points + 255) / 256, 256, 0, st>>>(S.cat.p, S.valid.p, S.pos.p, n_points, reinterpret_cast<float4*>(d_out_xyzi));
  count_launch(2);
  MR_CUDA(cudaGetLastError());
  int tail[2] = {0, 0};
  MR_CUDA(cudaMemcpyAsync(&tail[0], S.pos.p + (n_points - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(&tail[1], S.valid.p + (n_points - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  *n_out = tail[0] + tail[1];
  return MRSLAM_OK;
}
}
