// gicp_kernels.cuh -- sm_100a kernels of the FastGICP path.
//
// Data layout in HBM (per cloud):
//   porig  float4[n]        original order (x,y,z,1)
//   psort  float4[npad]     Morton-sorted copy, w = bit pattern of the ORIGINAL index (tie-breaks);
//                           padded to a multiple of 32 with far-away sentinels (w = INT_MAX)
//   tb     float[6][ntp]    per 32-point tile bounding boxes (SoA: minx,miny,minz,maxx,maxy,maxz)
//   cov    double[npad][6]  regularised covariances in SORTED order (xx,xy,xz,yy,yz,zz)
//
//   sb     float[6][nsp]    bounding boxes of "super tiles" (32 consecutive tiles = 1024 points)
//
// Search scheme (tiled brute force with box culling and a warp-shuffle k-min; no tree): one warp
// owns the 32 spatially coherent queries of one tile and serves them one after the other.  For a
// query the 32 lanes evaluate the 32 points of a candidate tile in parallel (coalesced 512 B load,
// brute force inside the tile); the running k-best list lives one entry per lane, kept sorted by
// (distance, original index), and a qualifying candidate is merged with one shuffle-up step.
// A tile (or a whole super tile) is skipped only when its bounding box is provably farther from the
// query than the current k-th distance, so the result is the exact kNN under the (distance, index)
// order -- identical to an exhaustive scan.
#pragma once
#include "common.cuh"

namespace mrslam {

constexpr int TILE = 32;
constexpr int PAD_IDX = INT_MAX;
constexpr float PAD_COORD = 3.0e18f;
constexpr float CULL_SAFETY = 0.99999f;  // box distance is a lower bound only up to a few ulp
#ifndef MR_HOME_RADIUS
#define MR_HOME_RADIUS 1
#endif
#ifndef MR_PF
#define MR_PF 2
#endif
#ifndef MR_SGROUP
#define MR_SGROUP 2
#endif
constexpr int HOME_RADIUS = MR_HOME_RADIUS;  // tiles qt-R..qt+R are searched first (they seed the cull radius)
constexpr int NRED = 28;                 // 21 (H upper) + 6 (b) + 1 (cost)

// ---------------------------------------------------------------------------------------------
// cloud preparation
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_points(const float* __restrict__ raw, int n, int stride, float4* __restrict__ porig) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = raw + (size_t)i * stride;
  porig[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

struct PackDesc {
  const float* raw;
  float4* porig;
  int n, stride;
};
// the same for many clouds in one launch (blockIdx.y = cloud)
__global__ void k_pack_points_batch(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) {
    const float* p = d.raw + (size_t)i * d.stride;
    d.porig[i] = make_float4(p[0], p[1], p[2], 1.0f);
  }
}

// ---- batched cloud build: every kernel handles all clouds of a batch (blockIdx.y = cloud) ----
struct BuildDesc {
  const float4* porig;
  float4* psort;
  float* tb;
  float* sb;
  unsigned long long* skeys;  // [n] sorted curve keys of this cloud (binary-search seeds of the first NN-1 pass)
  float* qp;  // [4] quantisation frame of the Morton keys: min x,y,z and scale (kept for Morton look-ups)
  int n, ntiles, ntiles_pad, nsuper_pad;
  long long offset;  // position of this cloud's keys in the concatenated sort arrays
  const float* raw;  // != nullptr: a host upload (xyz with raw_stride floats per point) still to be unpacked into porig
  int raw_stride;
};

// unpack the pending host uploads of a batch (blockIdx.y = cloud): one launch, in front of the build kernels
__global__ void k_pack_pending(const BuildDesc* __restrict__ descs) {
  const BuildDesc d = descs[blockIdx.y];
  if (!d.raw) return;
  float4* porig = const_cast<float4*>(d.porig);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) {
    const float* p = d.raw + (size_t)i * d.raw_stride;
    porig[i] = make_float4(p[0], p[1], p[2], 1.0f);
  }
}

__global__ void k_bbox_init(unsigned* bb, int nclouds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nclouds * 6) return;
  bb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) k_bbox(const BuildDesc* __restrict__ descs, unsigned* bball) {
  const BuildDesc D = descs[blockIdx.y];
  unsigned* bb = bball + 6 * blockIdx.y;
  __shared__ float s_mn[8][3], s_mx[8][3];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.n; i += gridDim.x * blockDim.x) {
    float4 v = D.porig[i];
    if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
      mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
      mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = warp_min(mn[a]);
    mx[a] = warp_max(mx[a]);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      s_mn[w][a] = mn[a];
      s_mx[w][a] = mx[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {  // one atomic pair per block and axis (8x fewer than per warp: the six addresses are contended)
    const int a = threadIdx.x;
    float lo = s_mn[0][a], hi = s_mx[0][a];
    for (int k = 1; k < 8; k++) {
      lo = fminf(lo, s_mn[k][a]);
      hi = fmaxf(hi, s_mx[k][a]);
    }
    atomicMin(&bb[a], f2ord(lo));
    atomicMax(&bb[3 + a], f2ord(hi));
  }
}

__device__ __forceinline__ unsigned long long spread16(unsigned v) {  // 16 bits -> every 3rd bit
  unsigned long long x = v & 0xffffu;
  x = (x | x << 16) & 0x0000ff0000ffull;
  x = (x | x << 8) & 0x00f00f00f00full;
  x = (x | x << 4) & 0x0c30c30c30c3ull;
  x = (x | x << 2) & 0x249249249249ull;
  return x;
}

#ifndef MR_HILBERT
#define MR_HILBERT 1
#endif
// The batched radix sort skips the lowest KEY_DROP_BITS bits of the curve key (4 per axis: cells of extent/2^12, ~5 cm for
// a 200 m scan) so that (cloud id | key) fits FIVE 8-bit passes; ties keep input order (stable sort).  Any order is correct for
// the exact searches -- the curve only makes tiles compact; measured with 6 (six passes) against 12 dropped bits: cloud build
// 0.413 -> 0.384 ms per 16 clouds, kNN / NN-1 times unchanged (gpurun_out/r2y).
#ifndef MR_KEY_DROP_BITS
#define MR_KEY_DROP_BITS 12
#endif
constexpr int KEY_DROP_BITS = MR_KEY_DROP_BITS;
constexpr unsigned long long KEY48_SORTED_MASK = 0xffffffffffffull & ~((1ull << KEY_DROP_BITS) - 1ull);
// 48-bit space-filling-curve key of a point in the cloud's bounding cube (16 bits per axis): Hilbert order (Skilling's
// axes-to-transpose transform + bit interleave) keeps 32 consecutive points in tighter boxes than Morton order does.
__device__ __forceinline__ unsigned long long morton48(float x, float y, float z, float mnx, float mny, float mnz, float s) {
  float f0 = (x - mnx) * s, f1 = (y - mny) * s, f2 = (z - mnz) * s;
  f0 = isfinite(f0) ? fminf(fmaxf(f0, 0.0f), 65535.0f) : 65535.0f;
  f1 = isfinite(f1) ? fminf(fmaxf(f1, 0.0f), 65535.0f) : 65535.0f;
  f2 = isfinite(f2) ? fminf(fmaxf(f2, 0.0f), 65535.0f) : 65535.0f;
  unsigned X0 = (unsigned)f0, X1 = (unsigned)f1, X2 = (unsigned)f2;
#if MR_HILBERT
#pragma unroll
  for (unsigned Q = 0x8000u; Q > 1; Q >>= 1) {
    const unsigned P = Q - 1;
    if (X0 & Q) X0 ^= P;  // i = 0: invert (the exchange branch is a no-op for i = 0)
    if (X1 & Q) X0 ^= P;
    else { const unsigned t = (X0 ^ X1) & P; X0 ^= t; X1 ^= t; }
    if (X2 & Q) X0 ^= P;
    else { const unsigned t = (X0 ^ X2) & P; X0 ^= t; X2 ^= t; }
  }
  X1 ^= X0;
  X2 ^= X1;
  unsigned t = 0;
#pragma unroll
  for (unsigned Q = 0x8000u; Q > 1; Q >>= 1)
    if (X2 & Q) t ^= Q - 1;
  X0 ^= t; X1 ^= t; X2 ^= t;
  return (spread16(X0) << 2) | (spread16(X1) << 1) | spread16(X2);
#else
  return spread16(X0) | (spread16(X1) << 1) | (spread16(X2) << 2);
#endif
}

// key = (cloud << 48) | 48-bit Morton code of the point in its cloud's bounding cube; value = index inside the cloud
__global__ void k_morton(const BuildDesc* __restrict__ descs, const unsigned* __restrict__ bball, unsigned long long* keys, int* vals) {
  const BuildDesc D = descs[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.n) return;
  const unsigned* bb = bball + 6 * blockIdx.y;
  float mn[3] = {ord2f(bb[0]), ord2f(bb[1]), ord2f(bb[2])};
  float mx[3] = {ord2f(bb[3]), ord2f(bb[4]), ord2f(bb[5])};
  float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), fmaxf(mx[2] - mn[2], 1e-6f));
  float sc = 65535.0f / ext;
  if (i == 0) {
    D.qp[0] = mn[0]; D.qp[1] = mn[1]; D.qp[2] = mn[2]; D.qp[3] = sc;
  }
  float4 v = D.porig[i];
  keys[D.offset + i] = ((unsigned long long)blockIdx.y << 48) | morton48(v.x, v.y, v.z, mn[0], mn[1], mn[2], sc);
  vals[D.offset + i] = i;
}

// one warp per tile: gather the sorted points, pad, and reduce the tile's bounding box
__global__ void k_gather_tiles(const BuildDesc* __restrict__ descs, const int* __restrict__ order_all,
                               const unsigned long long* __restrict__ keys_sorted) {
  const BuildDesc D = descs[blockIdx.y];
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (t >= D.ntiles_pad) return;
  const int* order = order_all + D.offset;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (t < D.ntiles) {
    int s = t * TILE + lane;
    float4 v;
    if (s < D.n) {
      int o = order[s];
      v = D.porig[o];
      v.w = __int_as_float(o);
      D.skeys[s] = keys_sorted[D.offset + s] & KEY48_SORTED_MASK;
      mn[0] = mx[0] = v.x; mn[1] = mx[1] = v.y; mn[2] = mx[2] = v.z;
    } else {
      v = make_float4(PAD_COORD, PAD_COORD, PAD_COORD, __int_as_float(PAD_IDX));
    }
    D.psort[s] = v;
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = warp_min(mn[a]);
    mx[a] = warp_max(mx[a]);
  }
  if (lane == 0) {
    bool real = t < D.ntiles;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      D.tb[(size_t)a * D.ntiles_pad + t] = real ? mn[a] : PAD_COORD;
      D.tb[(size_t)(3 + a) * D.ntiles_pad + t] = real ? mx[a] : PAD_COORD;
    }
  }
}

// one warp per super tile: box of its (real) tiles
__global__ void k_super_boxes(const BuildDesc* __restrict__ descs) {
  const BuildDesc D = descs[blockIdx.y];
  int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (s >= D.nsuper_pad) return;
  int t = s * 32 + lane;
  bool real = t < D.ntiles;
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    mn[a] = warp_min(real ? D.tb[(size_t)a * D.ntiles_pad + t] : FLT_MAX);
    mx[a] = warp_max(real ? D.tb[(size_t)(3 + a) * D.ntiles_pad + t] : -FLT_MAX);
  }
  bool any = __any_sync(FULL, real);
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      D.sb[(size_t)a * D.nsuper_pad + s] = any ? mn[a] : PAD_COORD;
      D.sb[(size_t)(3 + a) * D.nsuper_pad + s] = any ? mx[a] : PAD_COORD;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// shared search pieces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_lt(float ad, int ai, float bd, int bi) { return ad < bd || (ad == bd && ai < bi); }

// squared distance from a point to box t of an SoA box table (stride = padded count)
__device__ __forceinline__ float pt_box_dist2(float x, float y, float z, const float* __restrict__ b, int stride, int t) {
  float gx = fmaxf(0.0f, fmaxf(b[t] - x, x - b[3 * stride + t]));
  float gy = fmaxf(0.0f, fmaxf(b[stride + t] - y, y - b[4 * stride + t]));
  float gz = fmaxf(0.0f, fmaxf(b[2 * stride + t] - z, z - b[5 * stride + t]));
  return gx * gx + gy * gy + gz * gz;
}

// squared distance between two boxes (a: query-warp box, b: box t of an SoA table)
__device__ __forceinline__ float box_box_dist2(const float a[6], const float* __restrict__ b, int stride, int t) {
  float gx = fmaxf(0.0f, fmaxf(b[t] - a[3], a[0] - b[3 * stride + t]));
  float gy = fmaxf(0.0f, fmaxf(b[stride + t] - a[4], a[1] - b[4 * stride + t]));
  float gz = fmaxf(0.0f, fmaxf(b[2 * stride + t] - a[5], a[2] - b[5 * stride + t]));
  return gx * gx + gy * gy + gz * gz;
}

// ascending sort of one 32-bit key per lane; every compare-exchange keeps the smaller key in the lower lane (first step of a
// merge: partner = mirror image inside the block, then partner = lane ^ stride), so a step's direction is one bit of the lane id
__device__ __forceinline__ unsigned warp_sort_u32(unsigned key, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
    {
      const unsigned other = __shfl_xor_sync(FULL, key, size - 1);
      key = (lane & (size >> 1)) == 0 ? min(key, other) : max(key, other);
    }
#pragma unroll
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
      const unsigned other = __shfl_xor_sync(FULL, key, stride);
      key = (lane & stride) == 0 ? min(key, other) : max(key, other);
    }
  }
  return key;
}

// symmetric 3x3 eigen-decomposition, cyclic Jacobi (double); eigenvalues descending; same operation
// order as the oracle's sym_eig3 so that covariances agree bit for bit.
__device__ __forceinline__ void sym_eig3(const double A[9], double w[3], double V[9]) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    a[i] = A[i];
    V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int q = p + 1; q < 3; q++) {
        double apq = a[3 * p + q];
        if (apq == 0.0) continue;
        double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double akp = a[3 * k + p], akq = a[3 * k + q];
          a[3 * k + p] = c * akp - s * akq;
          a[3 * k + q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double apk = a[3 * p + k], aqk = a[3 * q + k];
          a[3 * p + k] = c * apk - s * aqk;
          a[3 * q + k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  double e0 = a[0], e1 = a[4], e2 = a[8];
  // stable selection of the descending order (matches std::sort on 3 distinct values)
  int o0 = 0, o1 = 1, o2 = 2;
  double w0 = e0, w1 = e1, w2 = e2;
  if (w1 > w0) { double tw = w0; w0 = w1; w1 = tw; int to = o0; o0 = o1; o1 = to; }
  if (w2 > w1) { double tw = w1; w1 = w2; w2 = tw; int to = o1; o1 = o2; o2 = to; }
  if (w1 > w0) { double tw = w0; w0 = w1; w1 = tw; int to = o0; o0 = o1; o1 = to; }
  w[0] = w0; w[1] = w1; w[2] = w2;
  double V2[9];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    V2[3 * i + 0] = V[3 * i + o0];
    V2[3 * i + 1] = V[3 * i + o1];
    V2[3 * i + 2] = V[3 * i + o2];
  }
#pragma unroll
  for (int i = 0; i < 9; i++) V[i] = V2[i];
}

// PLANE regularisation (fast_gicp_impl.hpp:273-293: SVD of the 3x3 covariance, singular values replaced by (1, 1, 1e-3)) needs
// only the direction n of the SMALLEST eigenvalue: U diag(1,1,1e-3) U^T = I - (1 - 1e-3) n n^T.  Instead of a full Jacobi
// eigen-decomposition (the round-1 k_cov spent ~2500 double-precision instructions per point in it), the smallest root of the
// characteristic polynomial of the trace-normalised matrix is found by Newton's iteration from 0 (monotone from below for a
// positive semi-definite matrix) and n is the largest cross product of two rows of (B - lambda I).  Only +, -, *, /, sqrt in a
// fixed order, so the CUDA kernel and the CPU oracle produce the same bits.  Where the two smallest eigenvalues coincide the
// direction inside their plane is arbitrary -- as it is for the reference's JacobiSVD.
__device__ __forceinline__ void plane_regularize(const double a[6] /* xx xy xz yy yz zz */, double out[6]) {
  const double tr = a[0] + a[3] + a[5];
  double n0 = 0.0, n1 = 0.0, n2 = 1.0;
  if (tr > 0.0) {
    const double inv = 1.0 / tr;
    const double b00 = a[0] * inv, b01 = a[1] * inv, b02 = a[2] * inv, b11 = a[3] * inv, b12 = a[4] * inv, b22 = a[5] * inv;
    const double c2 = b00 + b11 + b22;
    const double m01 = b00 * b11 - b01 * b01, m02 = b00 * b22 - b02 * b02, m12 = b11 * b22 - b12 * b12;
    const double c1 = m01 + m02 + m12;
    const double c0 = b00 * m12 - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
    double l = 0.0;
    for (int it = 0; it < 12; it++) {
      const double q = ((l - c2) * l + c1) * l - c0;
      const double dq = (3.0 * l - 2.0 * c2) * l + c1;
      if (!(dq > 0.0)) break;
      const double step = q / dq;
      if (!(step == step)) break;
      l = l - step;
      if (fabs(step) <= fabs(l) * 1e-15) break;  // converged to the last bits (typically after 4-6 of the 12 steps)
    }
    const double r00 = b00 - l, r11 = b11 - l, r22 = b22 - l;
    // rows of B - l I: (r00 b01 b02), (b01 r11 b12), (b02 b12 r22); the null direction is orthogonal to all of them
    const double x0 = b01 * b12 - b02 * r11, x1 = b02 * b01 - r00 * b12, x2 = r00 * r11 - b01 * b01;      // row0 x row1
    const double y0 = b01 * r22 - b02 * b12, y1 = b02 * b02 - r00 * r22, y2 = r00 * b12 - b01 * b02;      // row0 x row2
    const double z0 = r11 * r22 - b12 * b12, z1 = b12 * b02 - b01 * r22, z2 = b01 * b12 - r11 * b02;      // row1 x row2
    const double nx = x0 * x0 + x1 * x1 + x2 * x2, ny = y0 * y0 + y1 * y1 + y2 * y2, nz = z0 * z0 + z1 * z1 + z2 * z2;
    double v0 = x0, v1 = x1, v2 = x2, nn = nx;
    if (ny > nn) { v0 = y0; v1 = y1; v2 = y2; nn = ny; }
    if (nz > nn) { v0 = z0; v1 = z1; v2 = z2; nn = nz; }
    if (nn > 0.0) {
      const double s = 1.0 / sqrt(nn);
      n0 = v0 * s; n1 = v1 * s; n2 = v2 * s;
    } else {
      // rank <= 1 exactly (all points on one line): any direction orthogonal to the line; the line is the largest row
      double u0 = r00, u1 = b01, u2 = b02, un = r00 * r00 + b01 * b01 + b02 * b02;
      const double w1 = b01 * b01 + r11 * r11 + b12 * b12, w2 = b02 * b02 + b12 * b12 + r22 * r22;
      if (w1 > un) { u0 = b01; u1 = r11; u2 = b12; un = w1; }
      if (w2 > un) { u0 = b02; u1 = b12; u2 = r22; un = w2; }
      if (un > 0.0) {
        // cross with the coordinate axis the line is least aligned with
        const double a0 = fabs(u0), a1 = fabs(u1), a2 = fabs(u2);
        double e0 = 1.0, e1 = 0.0, e2 = 0.0;
        if (a1 < a0 && a1 <= a2) { e0 = 0.0; e1 = 1.0; }
        else if (a2 < a0 && a2 < a1) { e0 = 0.0; e2 = 1.0; }
        const double t0 = u1 * e2 - u2 * e1, t1 = u2 * e0 - u0 * e2, t2 = u0 * e1 - u1 * e0;
        const double s = 1.0 / sqrt(t0 * t0 + t1 * t1 + t2 * t2);
        n0 = t0 * s; n1 = t1 * s; n2 = t2 * s;
      }
    }
  }
  const double g = 1.0 - 1e-3;
  const double g0 = g * n0, g1 = g * n1, g2 = g * n2;
  out[0] = 1.0 - g0 * n0; out[1] = 0.0 - g0 * n1; out[2] = 0.0 - g0 * n2;
  out[3] = 1.0 - g1 * n1; out[4] = 0.0 - g1 * n2; out[5] = 1.0 - g2 * n2;
}

// inverse of a 3x3 given row-major a[9]; same expression order as the oracle's inverse3
__device__ __forceinline__ void inverse3(const double a[9], double r[9]) {
  double c00 = a[4] * a[8] - a[5] * a[7];
  double c01 = a[5] * a[6] - a[3] * a[8];
  double c02 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double id = 1.0 / det;
  r[0] = c00 * id;
  r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// RING++ local-descriptor front end (LoopDetection/src/RING_ros/util.py:124-170, covariation_eigenvalue after
// build_neighbors_NN): sample covariance (divisor k-1) of the k neighbours of every point, its three eigenvalues in descending
// order, the two of the x-y block, and the eigen-entropy (util.py:114-121).  One thread per point (ORIGINAL order); double
// arithmetic (the reference: float32 numpy + LAPACK ssyevd, so agreement is to float32 rounding of the reference itself).
__global__ void __launch_bounds__(128) k_ringpp_eigens(const float4* __restrict__ porig, int n, int k, const int* __restrict__ knn_idx,
                                                       float* __restrict__ cov_out, float* __restrict__ entropy_out,
                                                       float* __restrict__ eigens_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* nb = knn_idx + (size_t)i * k;
  double m[3] = {0, 0, 0};
  int cnt = 0;
  for (int j = 0; j < k; j++) {
    const int id = nb[j];
    if (id < 0) continue;
    const float4 p = porig[id];
    m[0] += (double)p.x; m[1] += (double)p.y; m[2] += (double)p.z;
    cnt++;
  }
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (cnt > 1) {
    m[0] /= cnt; m[1] /= cnt; m[2] /= cnt;
    for (int j = 0; j < k; j++) {
      const int id = nb[j];
      if (id < 0) continue;
      const float4 p = porig[id];
      const double d[3] = {(double)p.x - m[0], (double)p.y - m[1], (double)p.z - m[2]};
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = a; b < 3; b++) C[3 * a + b] = fma(d[a], d[b], C[3 * a + b]);
    }
    const double inv = 1.0 / (double)(cnt - 1);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = a; b < 3; b++) {
        C[3 * a + b] *= inv;
        C[3 * b + a] = C[3 * a + b];
      }
  }
  if (cov_out)
    for (int a = 0; a < 9; a++) cov_out[(size_t)i * 9 + a] = (float)C[a];
  double w[3], V[9];
  sym_eig3(C, w, V);  // descending
  // 2x2 block [a b; b c]: closed form, descending
  const double a = C[0], b = C[1], c = C[4];
  const double tr = 0.5 * (a + c), df = 0.5 * (a - c);
  const double rad = sqrt(fma(df, df, b * b));
  const double e2a = tr + rad, e2b = tr - rad;
  float* e = eigens_out + (size_t)i * 5;
  e[0] = (float)w[0]; e[1] = (float)w[1]; e[2] = (float)w[2]; e[3] = (float)e2a; e[4] = (float)e2b;
  if (entropy_out) {
    const double L = (w[0] - w[1]) / w[0], P = (w[1] - w[2]) / w[0], S = w[2] / w[0];
    entropy_out[i] = (float)(-L * log(L) - P * log(P) - S * log(S));
  }
}

// ---------------------------------------------------------------------------------------------
// kNN search (k <= 32), then covariance + regularisation as a separate lane-per-point kernel.
// fast_gicp_impl.hpp:241-298 (calculate_covariances); GPU precedent brute_force_knn.cu:16-60 +
// covariance_estimation.cu:16-42 + covariance_regularization.cu:14-97.
// ---------------------------------------------------------------------------------------------
struct KnnParams {
  const float4* psort;
  const float4* porig;
  const float* tb;
  const float* sb;
  int n, ntiles, ntiles_pad, nsuper, nsuper_pad;
  int k, reg;
  int* ids;         // [npad*k] neighbour ORIGINAL indices per SORTED position (PAD_IDX = absent)
  double* cov;      // [npad*6] sorted order (may be null when only the neighbour lists are wanted)
  int* knn_idx;     // optional test hook, original order [n*k]
  float* knn_sqd;   // optional
  unsigned long long* stat_pairs;  // optional: number of (query,candidate) distance evaluations
  unsigned char* fb;  // [ntiles] per-tile flag written by k_knn_search_v2: 1 = this tile must be redone by k_knn_search
};

constexpr int SGROUP = MR_SGROUP;  // super-tile box tests kept in flight together (SGROUP x 32 boxes)

// (distance, index) packed into one 64-bit key: distances are >= +0, so their bit patterns order like
// unsigned integers and key order == lexicographic (distance, index) order.
typedef unsigned long long u64;
__device__ __forceinline__ u64 make_key(float d, int idx) { return ((u64)__float_as_uint(d) << 32) | (unsigned)idx; }
__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_idx(u64 k) { return (int)(unsigned)(k & 0xffffffffull); }
constexpr u64 KEY_NONE = ((u64)0x7f800000u << 32) | 0x7fffffffu;  // (+inf, PAD_IDX)
constexpr int PF = MR_PF;  // candidate tiles fetched ahead of their (serial) processing
constexpr int CAND = 64;  // candidate tiles a warp collects for its shared nearest-first walk (2 per lane)

// ascending bitonic sort of one key per lane
__device__ __forceinline__ u64 warp_sort_keys(u64 key, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const u64 other = __shfl_xor_sync(FULL, key, stride);
      const bool take_min = ((lane & size) == 0) == ((lane & stride) == 0);
      // keep `key` when it already is the wanted one of the pair (keys are distinct or identical: ties are harmless)
      key = ((key < other) == take_min) ? key : other;
    }
  }
  return key;
}

// The same sort as a network whose compare-exchanges all point the same way (lower lane keeps the smaller key): the first
// step of every merge pairs lane i with its mirror image inside the block (xor with size-1), the others with i ^ stride, so the
// direction of a step depends on ONE bit of the lane number instead of two.
__device__ __forceinline__ u64 warp_sort_keys_mirror(u64 key, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
    {
      const u64 other = __shfl_xor_sync(FULL, key, size - 1);
      const bool low = (lane & (size >> 1)) == 0;
      key = ((key < other) == low) ? key : other;
    }
#pragma unroll
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
      const u64 other = __shfl_xor_sync(FULL, key, stride);
      const bool low = (lane & stride) == 0;
      key = ((key < other) == low) ? key : other;
    }
  }
  return key;
}

// Warp-cooperative exact kNN of ONE query (all lanes hold the same query): the k-best list lives one key per lane.
// Returns this lane's entry of the sorted list (lane j = j-th neighbour).
__device__ __forceinline__ u64 warp_query_knn(const KnnParams& P, float qx, float qy, float qz, int qt, int k, int lane,
                                              unsigned long long& tiles_done, unsigned long long& inserts) {
  const int ntp = P.ntiles_pad, nsp = P.nsuper_pad;
  // super-tile box tests do not depend on the list: issue their loads first
  float b1[SGROUP];
#pragma unroll
  for (int u = 0; u < SGROUP; u++) {
    const int s1 = u * 32 + lane;
    b1[u] = (s1 < P.nsuper) ? pt_box_dist2(qx, qy, qz, P.sb, nsp, s1) : INFINITY;
  }
  u64 my_key;
  u64 wkey;
  {  // own tile: all 32 candidates enter the (empty) list -> one bitonic sort instead of 32 insertions
    const float4 c = P.psort[(size_t)qt * TILE + lane];
#if MR_KNN_MIRRORSORT
    my_key = warp_sort_keys_mirror(make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w)), lane);
#else
    my_key = warp_sort_keys(make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w)), lane);
#endif
    wkey = __shfl_sync(FULL, my_key, k - 1);
    tiles_done++;
  }
  auto merge_tile = [&](const float4 c) {
    const u64 ckey = make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
    unsigned m = __ballot_sync(FULL, ckey < wkey);
    if (m) {
      // A candidate that no longer beats the (tightened) k-th entry only touches lanes >= k, which are
      // never read -- so no re-check and no refresh of wkey inside the loop.
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const u64 ck = __shfl_sync(FULL, ckey, j);
        const u64 up = __shfl_up_sync(FULL, my_key, 1);
        if (ck < my_key) my_key = (lane > 0 && ck < up) ? up : ck;
        inserts++;
      }
      wkey = __shfl_sync(FULL, my_key, k - 1);
    }
    tiles_done++;
  };
#pragma unroll 1
  for (int o = 1; o <= HOME_RADIUS; o++) {
    if (qt - o >= 0) merge_tile(P.psort[(size_t)(qt - o) * TILE + lane]);
    if (qt + o < P.ntiles) merge_tile(P.psort[(size_t)(qt + o) * TILE + lane]);
  }
#pragma unroll 1
  for (int g0 = 0; g0 < nsp; g0 += 32 * SGROUP) {
    if (g0 > 0) {
#pragma unroll
      for (int u = 0; u < SGROUP; u++) {
        const int s1 = g0 + u * 32 + lane;
        b1[u] = (s1 < P.nsuper) ? pt_box_dist2(qx, qy, qz, P.sb, nsp, s1) : INFINITY;
      }
    }
#pragma unroll
    for (int u = 0; u < SGROUP; u++) {
      unsigned ms = __ballot_sync(FULL, b1[u] * CULL_SAFETY <= key_dist(wkey));
      while (ms) {
        const int js = __ffs(ms) - 1;
        ms &= ms - 1;
        if (__shfl_sync(FULL, b1[u], js) * CULL_SAFETY > key_dist(wkey)) continue;
        const int sidx = g0 + u * 32 + js;
        const int t = sidx * 32 + lane;
        const float bt = pt_box_dist2(qx, qy, qz, P.tb, ntp, t);
        const int dt = t - qt;
        unsigned mt = __ballot_sync(FULL, t < P.ntiles && (dt > HOME_RADIUS || dt < -HOME_RADIUS) && bt * CULL_SAFETY <= key_dist(wkey));
        while (mt) {
          float4 cc[PF];
          float bb[PF];
#pragma unroll
          for (int f = 0; f < PF; f++) {  // fetch up to PF passing tiles before touching the list
            bb[f] = INFINITY;
            if (mt) {
              const int jt = __ffs(mt) - 1;
              mt &= mt - 1;
              bb[f] = __shfl_sync(FULL, bt, jt);
              cc[f] = P.psort[(size_t)(sidx * 32 + jt) * TILE + lane];
            }
          }
#pragma unroll
          for (int f = 0; f < PF; f++)
            if (bb[f] * CULL_SAFETY <= key_dist(wkey)) merge_tile(cc[f]);
        }
      }
    }
  }
  return my_key;
}

// ---- two queries per warp (k <= 16): each 16-lane half keeps the k-best list of its own query -------------------------------
// ascending (DESC = false) or descending sort of one key per lane inside each 16-lane half (xor masks < 16 never cross halves)
template <bool DESC>
__device__ __forceinline__ u64 half_sort_keys(u64 key, int l16) {
#pragma unroll
  for (int size = 2; size <= 16; size <<= 1) {
    {
      const u64 other = __shfl_xor_sync(FULL, key, size - 1);
      const bool low = ((l16 & (size >> 1)) == 0) != DESC;
      key = ((key < other) == low) ? key : other;
    }
#pragma unroll
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
      const u64 other = __shfl_xor_sync(FULL, key, stride);
      const bool low = ((l16 & stride) == 0) != DESC;
      key = ((key < other) == low) ? key : other;
    }
  }
  return key;
}
// bitonic sequence of 16 keys per half -> ascending
__device__ __forceinline__ u64 half_merge_keys(u64 key, int l16) {
#pragma unroll
  for (int stride = 8; stride > 0; stride >>= 1) {
    const u64 other = __shfl_xor_sync(FULL, key, stride);
    const bool low = (l16 & stride) == 0;
    key = ((key < other) == low) ? key : other;
  }
  return key;
}
// position of the highest set bit (-1 if none): one FLO
__device__ __forceinline__ int top_bit(unsigned v) {
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
// 1 << s with PTX semantics (s >= 32, e.g. (unsigned)-1, gives 0)
__device__ __forceinline__ unsigned bit_clamp(int s) {
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(1u), "r"(s));
  return r;
}
__device__ __forceinline__ unsigned half_min_u32(unsigned v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

__device__ __forceinline__ void knn_write_warp_list(const KnnParams& P, u64 my_key, size_t spos, int qidx, int k, int lane) {
  if (lane < k) {
    const int my_i = key_idx(my_key);
    P.ids[spos * k + lane] = my_i;
    if (P.knn_idx) {
      const bool real = my_i != PAD_IDX;
      P.knn_idx[(size_t)qidx * k + lane] = real ? my_i : -1;
      if (P.knn_sqd) P.knn_sqd[(size_t)qidx * k + lane] = real ? key_dist(my_key) : INFINITY;
    }
  }
}

// One warp per query tile; its 32 queries are served one after the other, each by the whole warp (k-best list one key
// per lane).  The candidate tiles are found ONCE per query tile: every query's k-th neighbour lies within
// B = max pairwise distance of the tile's own 32 points (they are k <= 32 candidates already), so only tiles whose box is
// within B of the query tile's box can contribute to any of its queries.  Their ids and boxes stay in registers (2 per
// lane).  A query then (1) sorts its own tile, (2) computes its point-to-box distance to every candidate -- one per lane --
// and (3) visits candidates nearest-first (warp redux-min over (box distance | slot) keys) until the nearest unvisited box
// is farther than its current k-th neighbour.  Exact: a skipped tile t has box distance > d_k >= every later d_k.
// Tiles with padding points or more than KCAND candidates take the per-query two-level culling path (warp_query_knn).
constexpr int KCAND = 64;
#ifndef MR_KNN_SHARED
#define MR_KNN_SHARED 1
#endif

#ifndef MR_KNN_MINBLK
#define MR_KNN_MINBLK 8
#endif
#ifndef MR_KNN_PAIR
#define MR_KNN_PAIR 1
#endif
#ifndef MR_KNN_UNIFORM_ROUNDS
#define MR_KNN_UNIFORM_ROUNDS 1
#endif
#ifndef MR_KNN_INS_UNROLL
#define MR_KNN_INS_UNROLL 2
#endif
#define MR_PRAGMA_(x) _Pragma(#x)
#define MR_UNROLL(n) MR_PRAGMA_(unroll n)
#ifndef MR_KNN_MIRRORSORT
#define MR_KNN_MIRRORSORT 1  // 6 instead of 7 instructions per compare-exchange step (SASS), 15 steps per query
#endif
template <bool STATS>
__global__ void __launch_bounds__(128, MR_KNN_MINBLK) k_knn_search(const KnnParams* __restrict__ plist, int only_flagged) {
  __shared__ int s_list[4][KCAND];
#if MR_KNN_PAIR
  __shared__ float4 s_own[4][TILE];
#endif
  const KnnParams P = plist[blockIdx.y];
  // second pass behind k_knn_search_v2: only the tiles it flagged (one flag per tile; warp-uniform exit further down)
  if (only_flagged && blockIdx.x * (blockDim.x >> 5) >= P.ntiles) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int qt = blockIdx.x * (blockDim.x >> 5) + wib;  // 4 warps per block, or 1 (no block-level step in this kernel)
  if (qt >= P.ntiles) return;  // warp-uniform; no block-level barrier in this kernel
  if (only_flagged && !P.fb[qt]) return;
  const float4 qmine = P.psort[(size_t)qt * TILE + lane];
  const int k = P.k;
  unsigned long long tiles_done = 0, inserts = 0;
  bool shared_ok = MR_KNN_SHARED && (qt + 1) * TILE <= P.n;
  int ncand = 0;
  if (shared_ok) {
    const int ntp = P.ntiles_pad, nsp = P.nsuper_pad;
    // B: an upper bound of every query's k-th neighbour distance.  The k tile points closest to a query IN CURVE ORDER (a window of
    // k consecutive positions around it) are k candidates already, so the farthest of them bounds the query's k-th distance; the
    // Hilbert order keeps such a window much tighter than the whole tile (round 1 used the largest pairwise distance of all 32
    // points, which put 6 % of the tiles over the 64-candidate limit and onto the per-query path below).
    float B = 0.0f;
    {
      const int w0 = max(0, min(lane - k / 2, TILE - k));  // window [w0, w0 + k) contains this lane (k <= 32)
#pragma unroll 4
      for (int j = 0; j < TILE; j++) {
        const float px = __shfl_sync(FULL, qmine.x, j), py = __shfl_sync(FULL, qmine.y, j), pz = __shfl_sync(FULL, qmine.z, j);
        const float d = dist2(qmine.x, qmine.y, qmine.z, px, py, pz);
        if (j >= w0 && j < w0 + k) B = fmaxf(B, d);
      }
    }
    B = warp_max(B);
    float a[6];
#pragma unroll
    for (int c = 0; c < 6; c++) a[c] = P.tb[(size_t)c * ntp + qt];
#pragma unroll 1
    for (int s0 = 0; s0 < P.nsuper && shared_ok; s0 += 32) {
      const int s = s0 + lane;
      const float bs = (s < P.nsuper) ? box_box_dist2(a, P.sb, nsp, s) : INFINITY;
      unsigned ms = __ballot_sync(FULL, bs * CULL_SAFETY <= B);
      while (ms) {
        const int js = __ffs(ms) - 1;
        ms &= ms - 1;
        const int t = (s0 + js) * 32 + lane;
        const bool pass = t < P.ntiles && t != qt && box_box_dist2(a, P.tb, ntp, t) * CULL_SAFETY <= B;
        const unsigned mt = __ballot_sync(FULL, pass);
        const int pos = ncand + __popc(mt & ((1u << lane) - 1u));
        if (pass && pos < KCAND) s_list[wib][pos] = t;
        ncand += __popc(mt);
        if (ncand > KCAND) {
          shared_ok = false;
          break;
        }
      }
    }
    __syncwarp();
  }
#if MR_KNN_PAIR
  if (shared_ok && k <= 16 && ncand <= 32) {
    // TWO queries at a time, one per 16-lane half (k <= 16 list entries fit a half).  Every warp-wide instruction of the walk --
    // the sorting network, the box tests, the candidate visit, each insertion round -- now serves two queries; a half that has
    // finished its walk (or has no qualifier left in an insertion round) idles while the other one completes.  A candidate tile
    // is visited as two points per lane; qualifiers are inserted in (lane, then second element) order, which does not matter for
    // the final list (exact sorted insertion of distinct keys).
    const int ntp = P.ntiles_pad;
    const int l16 = lane & 15, h16 = lane & 16;
    s_own[wib][lane] = qmine;
    const int t0 = (l16 < ncand) ? s_list[wib][l16] : -1;
    const int t1 = (l16 + 16 < ncand) ? s_list[wib][l16 + 16] : -1;
    float b0[6], b1[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      b0[c] = (t0 >= 0) ? P.tb[(size_t)c * ntp + t0] : 0.0f;
      b1[c] = (t1 >= 0) ? P.tb[(size_t)c * ntp + t1] : 0.0f;
    }
    __syncwarp();
    constexpr unsigned NONE = 0xffffffffu;
#pragma unroll 1
    for (int qp = 0; qp < TILE / 2; qp++) {
      const int qi = 2 * qp + (lane >> 4);
      const float4 q = s_own[wib][qi];
      const float qx = q.x, qy = q.y, qz = q.z;
      u64 my_key;
      {  // own tile: 16 smallest of its 32 keys, ascending (half-cleaner of an ascending and a descending run, then a merge)
        const float4 pa = s_own[wib][l16], pb = s_own[wib][l16 + 16];
        const u64 ka = half_sort_keys<false>(make_key(dist2(qx, qy, qz, pa.x, pa.y, pa.z), __float_as_int(pa.w)), l16);
        const u64 kb = half_sort_keys<true>(make_key(dist2(qx, qy, qz, pb.x, pb.y, pb.z), __float_as_int(pb.w)), l16);
        my_key = half_merge_keys(ka < kb ? ka : kb, l16);
      }
      u64 wkey = __shfl_sync(FULL, my_key, h16 + k - 1);
      if (STATS) tiles_done++;
      unsigned c0 = NONE, c1 = NONE;
      if (t0 >= 0) {
        const float gx = fmaxf(0.0f, fmaxf(b0[0] - qx, qx - b0[3])), gy = fmaxf(0.0f, fmaxf(b0[1] - qy, qy - b0[4])),
                    gz = fmaxf(0.0f, fmaxf(b0[2] - qz, qz - b0[5]));
        c0 = (__float_as_uint(gx * gx + gy * gy + gz * gz) & ~63u) | (unsigned)l16;
      }
      if (t1 >= 0) {
        const float gx = fmaxf(0.0f, fmaxf(b1[0] - qx, qx - b1[3])), gy = fmaxf(0.0f, fmaxf(b1[1] - qy, qy - b1[4])),
                    gz = fmaxf(0.0f, fmaxf(b1[2] - qz, qz - b1[5]));
        c1 = (__float_as_uint(gx * gx + gy * gy + gz * gz) & ~63u) | (unsigned)(16 + l16);
      }
      bool active = true;  // uniform inside a half
      while (true) {
        const unsigned m = half_min_u32(min(c0, c1));
        active = active && m != NONE && !(__uint_as_float(m & ~63u) * CULL_SAFETY > key_dist(wkey));
        if (!__any_sync(FULL, active)) break;
        u64 x0 = KEY_NONE, x1 = KEY_NONE;
        if (active) {
          const int slot = (int)(m & 63u);
          const int t = s_list[wib][slot];
          if (slot == l16) c0 = NONE;
          if (slot == l16 + 16) c1 = NONE;
          const float4 p0 = P.psort[(size_t)t * TILE + l16], p1 = P.psort[(size_t)t * TILE + 16 + l16];
          x0 = make_key(dist2(qx, qy, qz, p0.x, p0.y, p0.z), __float_as_int(p0.w));
          x1 = make_key(dist2(qx, qy, qz, p1.x, p1.y, p1.z), __float_as_int(p1.w));
          if (STATS) tiles_done++;
        }
        const unsigned q0 = __ballot_sync(FULL, x0 < wkey), q1 = __ballot_sync(FULL, x1 < wkey);
        unsigned mm = ((q0 >> h16) & 0xffffu) | (((q1 >> h16) & 0xffffu) << 16);  // the qualifiers of MY half's tile
        // insertion rounds: as many as the busier half has qualifiers (a warp-uniform count from a warp reduction, so the loop
        // needs no convergence checks and unrolls)
        const int rounds = (int)__reduce_max_sync(FULL, (unsigned)__popc(mm));
        MR_UNROLL(MR_KNN_INS_UNROLL)
        for (int it = 0; it < rounds; it++) {
          const bool on = mm != 0u;  // an idle half inserts nothing
          const int j = top_bit(mm);  // highest qualifier first (-1 when idle)
          mm &= ~bit_clamp(j);
          const u64 ck = __shfl_sync(FULL, (j & 16) ? x1 : x0, h16 | (j & 15));
          const u64 up = __shfl_up_sync(FULL, my_key, 1, 16);
          if (on && ck < my_key) my_key = (l16 > 0 && ck < up) ? up : ck;
          if (STATS && on) inserts++;
        }
        wkey = __shfl_sync(FULL, my_key, h16 + k - 1);
      }
      if (l16 < k) {
        const size_t spos = (size_t)qt * TILE + qi;
        const int qidx = __float_as_int(q.w);
        const int my_i = key_idx(my_key);
        P.ids[spos * k + l16] = my_i;
        if (P.knn_idx) {
          const bool real = my_i != PAD_IDX;
          P.knn_idx[(size_t)qidx * k + l16] = real ? my_i : -1;
          if (P.knn_sqd) P.knn_sqd[(size_t)qidx * k + l16] = real ? key_dist(my_key) : INFINITY;
        }
      }
    }
    if (STATS) {  // fold the upper half's counters into lane 0
      tiles_done += __shfl_sync(FULL, tiles_done, 16);
      inserts += __shfl_sync(FULL, inserts, 16);
    }
  } else
#endif
  if (shared_ok) {
    const int ntp = P.ntiles_pad;
    const int t0 = (lane < ncand) ? s_list[wib][lane] : -1;
    const int t1 = (lane + 32 < ncand) ? s_list[wib][lane + 32] : -1;
    float b0[6], b1[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
      b0[c] = (t0 >= 0) ? P.tb[(size_t)c * ntp + t0] : 0.0f;
      b1[c] = (t1 >= 0) ? P.tb[(size_t)c * ntp + t1] : 0.0f;
    }
    constexpr unsigned NONE = 0xffffffffu;
#pragma unroll 1
    for (int qi = 0; qi < TILE; qi++) {
      const float qx = __shfl_sync(FULL, qmine.x, qi), qy = __shfl_sync(FULL, qmine.y, qi), qz = __shfl_sync(FULL, qmine.z, qi);
      const int qidx = __shfl_sync(FULL, __float_as_int(qmine.w), qi);
      // own tile: all 32 candidates enter the (empty) list -> one bitonic sort instead of 32 insertions
#if MR_KNN_MIRRORSORT
      u64 my_key = warp_sort_keys_mirror(make_key(dist2(qx, qy, qz, qmine.x, qmine.y, qmine.z), __float_as_int(qmine.w)), lane);
#else
      u64 my_key = warp_sort_keys(make_key(dist2(qx, qy, qz, qmine.x, qmine.y, qmine.z), __float_as_int(qmine.w)), lane);
#endif
      u64 wkey = __shfl_sync(FULL, my_key, k - 1);
      if (STATS) tiles_done++;
      // per-query distance to the candidate boxes; the low 6 bits of the (non-negative) float carry the slot, which rounds
      // the box distance DOWN by < 2^-17 relative: conservative for the cull test
      unsigned k0 = NONE, k1 = NONE;
      if (t0 >= 0) {
        const float gx = fmaxf(0.0f, fmaxf(b0[0] - qx, qx - b0[3])), gy = fmaxf(0.0f, fmaxf(b0[1] - qy, qy - b0[4])),
                    gz = fmaxf(0.0f, fmaxf(b0[2] - qz, qz - b0[5]));
        k0 = (__float_as_uint(gx * gx + gy * gy + gz * gz) & ~63u) | (unsigned)lane;
      }
      if (t1 >= 0) {
        const float gx = fmaxf(0.0f, fmaxf(b1[0] - qx, qx - b1[3])), gy = fmaxf(0.0f, fmaxf(b1[1] - qy, qy - b1[4])),
                    gz = fmaxf(0.0f, fmaxf(b1[2] - qz, qz - b1[5]));
        k1 = (__float_as_uint(gx * gx + gy * gy + gz * gz) & ~63u) | (unsigned)(32 + lane);
      }
      while (true) {
        const unsigned m = __reduce_min_sync(FULL, min(k0, k1));
        if (m == NONE || __uint_as_float(m & ~63u) * CULL_SAFETY > key_dist(wkey)) break;
        const int slot = (int)(m & 63u);
        const int t = s_list[wib][slot];
        if (slot == lane) k0 = NONE;
        if (slot == lane + 32) k1 = NONE;
        const float4 c = P.psort[(size_t)t * TILE + lane];
        const u64 ckey = make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
        unsigned mm = __ballot_sync(FULL, ckey < wkey);
        if (mm) {
#if MR_KNN_UNIFORM_ROUNDS
          // the round count comes out of a warp reduction, i.e. in a uniform register: no convergence check per round, unrollable
          const int rounds = (int)__reduce_max_sync(FULL, (unsigned)__popc(mm));
          MR_UNROLL(MR_KNN_INS_UNROLL)
          for (int it = 0; it < rounds; it++) {
            const int j = top_bit(mm);
            mm ^= 1u << j;
#else
          while (mm) {
            const int j = __ffs(mm) - 1;
            mm &= mm - 1;
#endif
            const u64 ck = __shfl_sync(FULL, ckey, j);
            const u64 up = __shfl_up_sync(FULL, my_key, 1);
            if (ck < my_key) my_key = (lane > 0 && ck < up) ? up : ck;
            if (STATS) inserts++;
          }
          wkey = __shfl_sync(FULL, my_key, k - 1);
        }
        if (STATS) tiles_done++;
      }
      knn_write_warp_list(P, my_key, (size_t)qt * TILE + qi, qidx, k, lane);
    }
  } else {
#pragma unroll 1
    for (int qi = 0; qi < TILE; qi++) {
      const float qx = __shfl_sync(FULL, qmine.x, qi), qy = __shfl_sync(FULL, qmine.y, qi), qz = __shfl_sync(FULL, qmine.z, qi);
      const int qidx = __shfl_sync(FULL, __float_as_int(qmine.w), qi);
      if (qidx == PAD_IDX) continue;
      const u64 my_key = warp_query_knn(P, qx, qy, qz, qt, k, lane, tiles_done, inserts);
      knn_write_warp_list(P, my_key, (size_t)qt * TILE + qi, qidx, k, lane);
    }
  }
  if (STATS && P.stat_pairs && lane == 0) {
    atomicAdd(P.stat_pairs, tiles_done * 32ull);
    atomicAdd(P.stat_pairs + 2, inserts);
    atomicAdd(P.stat_pairs + 4, shared_ok ? 0ull : 1ull);
    atomicAdd(P.stat_pairs + 5, (unsigned long long)ncand);
  }
}

// ---------------------------------------------------------------------------------------------
// kNN v2 (k <= 15): the same exact tiled brute force, re-organised around two observations from the round-1 profile
// (701 warp-instructions per query, 30 % of them serial shuffle-up insertions, 20 % the 64-bit bitonic sort of the own tile):
//   * list maintenance is ~6x cheaper LANE-PER-QUERY: each lane keeps the k-best list of ITS query as 16 sorted 64-bit
//     (distance, index) keys in registers and merges candidates sixteen at a time with fully unrolled sorting networks
//     (Batcher odd-even sort of the batch, 63 compare-exchanges; bitonic merge with the list, 16 + 32) -- no shuffles, 32
//     queries per instruction;
//   * distance evaluation is cheapest WARP-PER-QUERY (one candidate point per lane), but a visit does not have to touch the
//     list: qualifiers (key < the query's current k-th key) are ballot-compacted into a per-query buffer in shared memory as
//     16-bit (staged tile, lane) ids, at a cost that does not depend on how many there are.
// The candidate tiles of a query tile -- every tile whose box is within R = max over its queries of the own-tile k-th distance
// of the tile's box -- are found once per warp and staged in shared memory 32 tiles at a time by 1-D bulk-async copies
// (cp.async.bulk, completion on an mbarrier), while the lanes compute which staged tiles their own query needs
// (point-to-box test against the query's own k-th distance).
// Exactness: a tile is skipped for a query only if its box is provably farther than the query's current k-th key, and a
// candidate is dropped only if its key is not below that key, so the final list is the exact kNN under (distance, index).
// Tiles with more than V2_MAXC candidate tiles are flagged in P.fb and redone by k_knn_search.
// ---------------------------------------------------------------------------------------------
constexpr int KP = 16;
#ifndef MR_V2_CH
#define MR_V2_CH 24
#endif
#ifndef MR_V2_CAP
#define MR_V2_CAP 32
#endif
constexpr int V2_CH = MR_V2_CH;    // candidate tiles staged at a time (<= 32: one bit each in a lane's need mask); 24 tiles = 12 KB
constexpr int V2_MAXC = 128;       // candidate tiles a query tile may have; more -> k_knn_search redoes the tile
constexpr int V2_CAP = MR_V2_CAP;  // buffered qualifiers per query between merges (a visit adds at most 32)
constexpr int V2_BSTRIDE = V2_CAP + 2;  // u16 per row, an ODD number of 32-bit words: lane q reading row q is bank-conflict free
static_assert(V2_CH <= 32 && V2_CAP >= 32 && (V2_BSTRIDE / 2) % 2 == 1 && V2_BSTRIDE * 2 >= 15 * 4, "kNN v2 shared-memory layout");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP); bytes and both addresses are multiples of 16
__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void key_ce(u64& a, u64& b) {
  const bool sw = b < a;
  const u64 lo = sw ? b : a, hi = sw ? a : b;
  a = lo;
  b = hi;
}
// Batcher's odd-even merge sort of 16 keys held in registers: 63 compare-exchanges, every index a compile-time constant
__device__ __forceinline__ void key_sort16(u64 (&v)[KP]) {
  // (generated from the loop form p = 1,2,4,8; k = p..1; j = k % p, step 2k; i < k; same 2p-block -> compare v[i+j], v[i+j+k])
  key_ce(v[0], v[1]); key_ce(v[2], v[3]); key_ce(v[4], v[5]); key_ce(v[6], v[7]); key_ce(v[8], v[9]);
  key_ce(v[10], v[11]); key_ce(v[12], v[13]); key_ce(v[14], v[15]); key_ce(v[0], v[2]); key_ce(v[1], v[3]);
  key_ce(v[4], v[6]); key_ce(v[5], v[7]); key_ce(v[8], v[10]); key_ce(v[9], v[11]); key_ce(v[12], v[14]);
  key_ce(v[13], v[15]); key_ce(v[1], v[2]); key_ce(v[5], v[6]); key_ce(v[9], v[10]); key_ce(v[13], v[14]);
  key_ce(v[0], v[4]); key_ce(v[1], v[5]); key_ce(v[2], v[6]); key_ce(v[3], v[7]); key_ce(v[8], v[12]);
  key_ce(v[9], v[13]); key_ce(v[10], v[14]); key_ce(v[11], v[15]); key_ce(v[2], v[4]); key_ce(v[3], v[5]);
  key_ce(v[10], v[12]); key_ce(v[11], v[13]); key_ce(v[1], v[2]); key_ce(v[3], v[4]); key_ce(v[5], v[6]);
  key_ce(v[9], v[10]); key_ce(v[11], v[12]); key_ce(v[13], v[14]); key_ce(v[0], v[8]); key_ce(v[1], v[9]);
  key_ce(v[2], v[10]); key_ce(v[3], v[11]); key_ce(v[4], v[12]); key_ce(v[5], v[13]); key_ce(v[6], v[14]);
  key_ce(v[7], v[15]); key_ce(v[4], v[8]); key_ce(v[5], v[9]); key_ce(v[6], v[10]); key_ce(v[7], v[11]);
  key_ce(v[2], v[4]); key_ce(v[3], v[5]); key_ce(v[6], v[8]); key_ce(v[7], v[9]); key_ce(v[10], v[12]);
  key_ce(v[11], v[13]); key_ce(v[1], v[2]); key_ce(v[3], v[4]); key_ce(v[5], v[6]); key_ce(v[7], v[8]);
  key_ce(v[9], v[10]); key_ce(v[11], v[12]); key_ce(v[13], v[14]);
}
// L (sorted) <- the 16 smallest of L u B (both sorted ascending): element-wise min against the reversed batch leaves a bitonic
// sequence holding exactly those 16, which the bitonic merge network sorts (16 half + 32 full compare-exchanges)
__device__ __forceinline__ void key_merge16(u64 (&L)[KP], const u64 (&B)[KP]) {
#pragma unroll
  for (int i = 0; i < KP; i++) L[i] = B[KP - 1 - i] < L[i] ? B[KP - 1 - i] : L[i];
#pragma unroll
  for (int kk = KP / 2; kk >= 1; kk >>= 1)
#pragma unroll
    for (int i = 0; i < KP; i++)
      if ((i & kk) == 0) key_ce(L[i], L[i | kk]);
}

// One warp = one block = one query tile (32 queries); nothing is shared between warps, so there is no block-level barrier.
// Steps: (1) own tile -> shared; every lane sorts its own 32 distances into its register list (two batches of 16);
// (2) the warp collects the candidate tiles: boxes within R_w = largest own-tile k-th distance of its queries of the tile's box;
// (3) per chunk of <= V2_CH candidates: the tiles are staged by 1-D bulk-async copies while every lane computes which of
// them ITS query needs (point-to-box test against its own k-th distance) and which of those is nearest;
// (4) phase 1: warp-per-query scan of only the NEAREST needed tile of every query, then one lane-per-query merge -- the k-th
// distances are now close to final -- (5) phase 2: the need masks are recomputed with the tight distances, the remaining
// tiles are scanned and merged.  The first version of this kernel (one scan with the loose own-tile distance; round-2 profile
// r2c) buffered 19.6 qualifiers per query and, because a merge costs what the FULLEST buffer of the warp needs, spent 57 % of
// its 751 warp-instructions per query in merge rounds; the two-phase order cuts qualifiers and rounds.
template <bool STATS>
__global__ void __launch_bounds__(32) k_knn_search_v2(const KnnParams* __restrict__ plist) {
  __shared__ __align__(128) float4 s_stage[V2_CH * TILE];  // candidate tiles of the current chunk (bulk-copy destination)
  __shared__ __align__(16) float4 s_q[TILE];               // own tile
  __shared__ __align__(16) float4 s_box[V2_CH][2];         // boxes of the staged tiles (min xyz, max xyz)
  __shared__ int s_cand[V2_MAXC];                          // candidate tile ids
  __shared__ __align__(16) unsigned short s_buf[TILE][V2_BSTRIDE];  // per query: buffered qualifiers (staged slot * 32 + lane)
  __shared__ __align__(8) unsigned long long s_mbar;

  const KnnParams P = plist[blockIdx.y];
  const int lane = threadIdx.x;
  const int qt = blockIdx.x;
  if (qt >= P.ntiles) return;
  const int k = P.k, ntp = P.ntiles_pad, nsp = P.nsuper_pad;
  const unsigned lt = (1u << lane) - 1u;
  if (lane == 0) {
    mbar_init(&s_mbar, 1);
    fence_mbar_init();
  }
  // ---- (1) own tile -> shared; lane-per-query initial list from its 32 points ----
  const float4 qmine = P.psort[(size_t)qt * TILE + lane];
  s_q[lane] = qmine;
  __syncwarp();
  const float qx = qmine.x, qy = qmine.y, qz = qmine.z;
  const int qidx = __float_as_int(qmine.w);
  const bool qvalid = qidx != PAD_IDX;
  // The list always holds 16 keys: KP - k minimal sentinels (key 0) in front, so that the k-th best is L[KP-1] whatever k is
  // and every index into L stays a compile-time constant.  (A real key 0 -- point 0 at distance 0 -- is identical to a
  // sentinel, so it does not matter which of the two copies ends up in front.)
  u64 L[KP];
#pragma unroll
  for (int i = 0; i < KP; i++) L[i] = (i < KP - k) ? 0ull : KEY_NONE;
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
    u64 Bt[KP];
#pragma unroll
    for (int j = 0; j < KP; j++) {
      const float4 c = s_q[h * KP + j];
      Bt[j] = make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
    }
    key_sort16(Bt);
    key_merge16(L, Bt);
  }
  float tau_d = key_dist(L[KP - 1]);
  unsigned long long st_visits = 0, st_quals = 0, st_rounds = 0, st_oflow = 0;

  // ---- (2) candidate tiles of this warp ----
  int ncand = 0;
  {
    const float Rw = warp_max(qvalid ? tau_d : 0.0f);
    float a[6];
#pragma unroll
    for (int c = 0; c < 6; c++) a[c] = P.tb[(size_t)c * ntp + qt];
#pragma unroll 1
    for (int s0 = 0; s0 < P.nsuper; s0 += 32) {
      const int sidx = s0 + lane;
      const float bs = (sidx < P.nsuper) ? box_box_dist2(a, P.sb, nsp, sidx) : INFINITY;
      unsigned ms = __ballot_sync(FULL, bs * CULL_SAFETY <= Rw);
      while (ms) {
        const int js = __ffs(ms) - 1;
        ms &= ms - 1;
        const int t = (s0 + js) * 32 + lane;
        const bool pass = t < P.ntiles && t != qt && box_box_dist2(a, P.tb, ntp, t) * CULL_SAFETY <= Rw;
        const unsigned mt = __ballot_sync(FULL, pass);
        const int pos = ncand + __popc(mt & lt);
        if (pass && pos < V2_MAXC) s_cand[pos] = t;
        ncand += __popc(mt);
      }
    }
  }
  if (ncand > V2_MAXC) {  // warp-uniform: too many candidates for the staged scheme (huge R: sparse outliers, n < k)
    if (lane == 0) {
      P.fb[qt] = 1;
      if (STATS && P.stat_pairs) atomicAdd(P.stat_pairs + 4, 1ull << 32);  // high word: tiles handed to k_knn_search
    }
    return;
  }
  if (lane == 0) P.fb[qt] = 0;
  __syncwarp();

  // lane-per-query merge of the buffered qualifiers, 16 per round (the staged tiles are resident)
  int cnt = 0;  // buffered qualifiers of MY query
  auto merge_buffers = [&]() {
    const int rounds = (__reduce_max_sync(FULL, cnt) + KP - 1) / KP;
#pragma unroll 1
    for (int r = 0; r < rounds; r++) {
      u64 Bt[KP];
      const unsigned* rowq = reinterpret_cast<const unsigned*>(&s_buf[lane][r * KP]);
#pragma unroll
      for (int j2 = 0; j2 < KP / 2; j2++) {
        const unsigned two = rowq[j2];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int j = 2 * j2 + hh;
          const unsigned id = hh ? (two >> 16) : (two & 0xffffu);
          u64 key = KEY_NONE;
          if (r * KP + j < cnt) {
            const float4 c = s_stage[id];
            key = make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
          }
          Bt[j] = key;
        }
      }
      key_sort16(Bt);
      key_merge16(L, Bt);
      if (STATS) st_rounds++;
    }
    cnt = 0;
    tau_d = key_dist(L[KP - 1]);
    __syncwarp();
  };
  // warp-per-query scan of the tiles in `need` (bit i = staged slot i) of every query; qualifiers go to the buffers.  Returns
  // false when a buffer would overflow: the caller merges and calls again (the unvisited tiles stay in `need`).
  auto scan = [&](unsigned& need) -> bool {
#pragma unroll 1
    for (int qi = 0; qi < TILE; qi++) {
      unsigned m = __shfl_sync(FULL, need, qi);
      if (!m) continue;
      const float ox = __shfl_sync(FULL, qx, qi), oy = __shfl_sync(FULL, qy, qi), oz = __shfl_sync(FULL, qz, qi);
      const u64 otau = __shfl_sync(FULL, L[KP - 1], qi);
      int ocnt = __shfl_sync(FULL, cnt, qi);
      unsigned short* row = &s_buf[qi][0];
      bool overflow = false;
      while (m) {
        const int i = __ffs(m) - 1;
        const float4 c = s_stage[i * TILE + lane];
        const u64 ckey = make_key(dist2(ox, oy, oz, c.x, c.y, c.z), __float_as_int(c.w));
        const bool pass = ckey < otau;
        const unsigned bal = __ballot_sync(FULL, pass);
        const int np = __popc(bal);
        if (ocnt + np > V2_CAP) {  // warp-uniform
          overflow = true;
          break;
        }
        if (pass) row[ocnt + __popc(bal & lt)] = (unsigned short)(i * TILE + lane);
        ocnt += np;
        m &= m - 1;
        if (STATS) st_visits++, st_quals += np;
      }
      if (lane == qi) {
        cnt = ocnt;
        need = m;
      }
      if (overflow) return false;
    }
    __syncwarp();
    return true;
  };

  // ---- (3)-(5) chunks of V2_CH staged tiles ----
  unsigned phase = 0;
#pragma unroll 1
  for (int c0 = 0; c0 < ncand; c0 += V2_CH) {
    const int nch = min(V2_CH, ncand - c0);
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(&s_mbar, (unsigned)nch * TILE * 16u);
    __syncwarp();
    int myt = -1;
    if (lane < nch) {
      myt = s_cand[c0 + lane];
      bulk_load_1d(&s_stage[lane * TILE], P.psort + (size_t)myt * TILE, TILE * 16u, &s_mbar);
      s_box[lane][0] = make_float4(P.tb[myt], P.tb[(size_t)ntp + myt], P.tb[(size_t)2 * ntp + myt], 0.0f);
      s_box[lane][1] = make_float4(P.tb[(size_t)3 * ntp + myt], P.tb[(size_t)4 * ntp + myt], P.tb[(size_t)5 * ntp + myt], 0.0f);
    }
    __syncwarp();
    // which staged tiles does MY query need (point-to-box distance against my current k-th distance), and which is nearest?
    unsigned need = 0;
    int nearest = -1;
    float nearest_d = INFINITY;
#pragma unroll 4
    for (int i = 0; i < nch; i++) {
      const float4 bmn = s_box[i][0], bmx = s_box[i][1];
      const float gx = fmaxf(0.0f, fmaxf(bmn.x - qx, qx - bmx.x)), gy = fmaxf(0.0f, fmaxf(bmn.y - qy, qy - bmx.y)),
                  gz = fmaxf(0.0f, fmaxf(bmn.z - qz, qz - bmx.z));
      const float bd = gx * gx + gy * gy + gz * gz;
      if (bd * CULL_SAFETY <= tau_d) {
        need |= 1u << i;
        if (bd < nearest_d) {
          nearest_d = bd;
          nearest = i;
        }
      }
    }
    if (!qvalid) need = 0;
    mbar_wait(&s_mbar, phase);
    phase ^= 1u;
    // phase 1: only the nearest needed tile of every query (at most 32 qualifiers each: cannot overflow)
    unsigned first = (need && nearest >= 0) ? (1u << nearest) : 0u;
    need &= ~first;
    scan(first);
    merge_buffers();
    // phase 2: the other tiles, against the tightened k-th distances
    if (need) {
      unsigned keep = 0;
      unsigned m = need;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const float4 bmn = s_box[i][0], bmx = s_box[i][1];
        const float gx = fmaxf(0.0f, fmaxf(bmn.x - qx, qx - bmx.x)), gy = fmaxf(0.0f, fmaxf(bmn.y - qy, qy - bmx.y)),
                    gz = fmaxf(0.0f, fmaxf(bmn.z - qz, qz - bmx.z));
        if ((gx * gx + gy * gy + gz * gz) * CULL_SAFETY <= tau_d) keep |= 1u << i;
      }
      need = keep;
    }
#pragma unroll 1
    while (true) {
      const bool done = scan(need);
      merge_buffers();
      if (done) break;
      if (STATS) st_oflow++;
    }
  }

  // ---- results: neighbour j of my query is L[KP - k + j]; transpose through shared memory for coalesced stores ----
  __syncwarp();
  int* s_out = reinterpret_cast<int*>(&s_buf[0][0]);  // 32 x k ints <= 1920 B of the 4224 B buffer
#pragma unroll
  for (int i = 0; i < KP; i++) {
    const int j = i - (KP - k);
    if (j >= 0) s_out[lane * k + j] = key_idx(L[i]);
  }
  __syncwarp();
  {
    int* dst = P.ids + (size_t)qt * TILE * k;
    for (int e = lane; e < TILE * k; e += 32) dst[e] = s_out[e];
  }
  if (P.knn_idx && qvalid) {
#pragma unroll
    for (int i = 0; i < KP; i++) {
      const int j = i - (KP - k);
      if (j >= 0) {
        const int id = key_idx(L[i]);
        const bool real = id != PAD_IDX;
        P.knn_idx[(size_t)qidx * k + j] = real ? id : -1;
        if (P.knn_sqd) P.knn_sqd[(size_t)qidx * k + j] = real ? key_dist(L[i]) : INFINITY;
      }
    }
  }
  if (STATS && P.stat_pairs) {
    st_visits = __shfl_sync(FULL, st_visits, 0);
    st_quals = __shfl_sync(FULL, st_quals, 0);
    if (lane == 0) {
      atomicAdd(P.stat_pairs, st_visits * 32ull + 1024ull);  // (query, candidate) distance evaluations of this warp (32 x 32 own tile incl.)
      atomicAdd(P.stat_pairs + 2, st_quals);
      atomicAdd(P.stat_pairs + 5, (unsigned long long)ncand);
      atomicAdd(P.stat_pairs + 6, st_rounds);
      atomicAdd(P.stat_pairs + 7, st_oflow);
    }
  }
}

// (A lane-per-query variant with per-lane sorted lists in registers -- the scheme that pays off for NN-1 -- was measured
// 5x slower here: with 32 neighbouring queries almost every scanned candidate enters SOME lane's list, so the whole warp
// executes the O(k) register insertion for nearly every candidate.  The warp-shuffle k-min above stays the kNN kernel.)

// regularisation of a neighbourhood covariance (fast_gicp_impl.hpp:264-293); C symmetric 3x3 row-major, out: upper triangle valid
__device__ __forceinline__ void regularize_cov(int reg, const double C[9], double out[9]) {
  if (reg == MRSLAM_REG_NONE) {
#pragma unroll
    for (int a = 0; a < 9; a++) out[a] = C[a];
  } else if (reg == MRSLAM_REG_FROBENIUS) {
    double Cl[9], Ci[9], Cn[9];
#pragma unroll
    for (int a = 0; a < 9; a++) Cl[a] = C[a];
    Cl[0] += 1e-3; Cl[4] += 1e-3; Cl[8] += 1e-3;
    inverse3(Cl, Ci);
    double nrm = 0.0;
#pragma unroll
    for (int a = 0; a < 9; a++) nrm += Ci[a] * Ci[a];
    nrm = sqrt(nrm);
#pragma unroll
    for (int a = 0; a < 9; a++) Cn[a] = Ci[a] / nrm;
    inverse3(Cn, out);
  } else if (reg == MRSLAM_REG_PLANE) {
    const double c6[6] = {C[0], C[1], C[2], C[4], C[5], C[8]};
    double o6[6];
    plane_regularize(c6, o6);
    out[0] = o6[0]; out[1] = o6[1]; out[2] = o6[2]; out[4] = o6[3]; out[5] = o6[4]; out[8] = o6[5];
  } else {
    double w[3], V[9], vals[3];
    sym_eig3(C, w, V);
    if (reg == MRSLAM_REG_MIN_EIG) {
      vals[0] = fmax(w[0], 1e-3); vals[1] = fmax(w[1], 1e-3); vals[2] = fmax(w[2], 1e-3);
    } else {
      vals[0] = fmax(w[0] / w[0], 1e-3); vals[1] = fmax(w[1] / w[0], 1e-3); vals[2] = fmax(w[2] / w[0], 1e-3);
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = a; b < 3; b++)
        out[3 * a + b] = V[3 * a] * vals[0] * V[3 * b] + V[3 * a + 1] * vals[1] * V[3 * b + 1] + V[3 * a + 2] * vals[2] * V[3 * b + 2];
  }
}

// covariance of the k neighbours (double), one thread per SORTED point; grid = (ceil(npad/128), n_clouds)
#ifndef MR_COV_MINBLK
#define MR_COV_MINBLK 4  // 128 registers, four blocks per SM: 0.222 vs 0.257 ms per 16 clouds without the cap (gpurun_out/r2u)
#endif
__global__ void __launch_bounds__(128, MR_COV_MINBLK) k_cov(const KnnParams* __restrict__ plist) {
  const KnnParams P = plist[blockIdx.y];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.ntiles * TILE) return;
  if (__float_as_int(P.psort[s].w) == PAD_IDX) return;
  const int k = P.k;
  const int* ids = P.ids + (size_t)s * k;
  // fast_gicp_impl.hpp:256-261.  r2: the ncu profile of round 1 blamed the double-precision Jacobi, but replacing it (below) moved the
  // kernel by 10 % only -- the two passes over the k neighbours are what costs: 2 x k dependent 16-byte gathers per point, issued one
  // at a time.  For k <= 16 (the nodes use 15) the neighbours are fetched ONCE, all loads in flight together, and both passes run
  // from registers, in the same summation order as before (so the results do not change by a bit).
  double mean[3] = {0.0, 0.0, 0.0};
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const double dk = (double)k;
  if (k <= 16) {
    int id[16];
#pragma unroll
    for (int j = 0; j < 16; j++) id[j] = j < k ? ids[j] : PAD_IDX;
    float px[16], py[16], pz[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (id[j] != PAD_IDX) p = P.porig[id[j]];
      px[j] = p.x; py[j] = p.y; pz[j] = p.z;
    }
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j < k && id[j] != PAD_IDX) {
        mean[0] += (double)px[j];
        mean[1] += (double)py[j];
        mean[2] += (double)pz[j];
      }
    mean[0] /= dk; mean[1] /= dk; mean[2] /= dk;
#pragma unroll
    for (int j = 0; j < 16; j++)
      if (j < k) {
        double dd[3];
        if (id[j] != PAD_IDX) {
          dd[0] = (double)px[j] - mean[0]; dd[1] = (double)py[j] - mean[1]; dd[2] = (double)pz[j] - mean[2];
        } else {
          dd[0] = 0.0 - mean[0]; dd[1] = 0.0 - mean[1]; dd[2] = 0.0 - mean[2];
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int b = 0; b < 3; b++) C[3 * a + b] += dd[a] * dd[b];
      }
  } else {
#pragma unroll 1
    for (int j = 0; j < k; j++) {
      const int id = ids[j];
      if (id != PAD_IDX) {
        const float4 p = P.porig[id];
        mean[0] += (double)p.x;
        mean[1] += (double)p.y;
        mean[2] += (double)p.z;
      }
    }
    mean[0] /= dk; mean[1] /= dk; mean[2] /= dk;
#pragma unroll 1
    for (int j = 0; j < k; j++) {
      const int id = ids[j];
      double dd[3];
      if (id != PAD_IDX) {
        const float4 p = P.porig[id];
        dd[0] = (double)p.x - mean[0]; dd[1] = (double)p.y - mean[1]; dd[2] = (double)p.z - mean[2];
      } else {
        dd[0] = 0.0 - mean[0]; dd[1] = 0.0 - mean[1]; dd[2] = 0.0 - mean[2];
      }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) C[3 * a + b] += dd[a] * dd[b];
    }
  }
#pragma unroll
  for (int a = 0; a < 9; a++) C[a] /= dk;
  // (r2: splitting the regularisation into a second, high-occupancy kernel was measured and lost -- 0.285 vs 0.254 ms per 16
  // clouds: the extra 96 B/pt round trip costs more than the occupancy gains)
  double out[9];
  regularize_cov(P.reg, C, out);
  double* co = P.cov + (size_t)s * 6;
  co[0] = out[0]; co[1] = out[1]; co[2] = out[2]; co[3] = out[4]; co[4] = out[5]; co[5] = out[8];
}


// ---------------------------------------------------------------------------------------------
// NN-1 correspondence search (warp per query tile), then Mahalanobis + J^T M J / J^T M e / cost as a
// lane-per-point kernel.  fast_gicp_impl.hpp:115-152 (update_correspondences) + :155-211 (linearize);
// GPU precedent find_voxel_correspondences.cu + compute_mahalanobis.cu + compute_derivatives.cu.
// ---------------------------------------------------------------------------------------------
struct PairParams {
  const float4* src_sort;
  const double* src_cov;
  int src_ntiles;
  const float4* tgt_sort;
  const float* tgt_tb;
  const float* tgt_sb;
  const float* tgt_qp;  // quantisation frame of the target's curve keys (4 floats)
  const unsigned long long* tgt_keys;  // sorted curve keys of the target
  const double* tgt_cov;
  int tgt_n, tgt_ntiles, tgt_ntiles_pad, tgt_nsuper, tgt_nsuper_pad;
  int* corr;       // [src_npad] target SORTED position, -1 = none
  int* nnpos;      // [src_npad] last nearest neighbour (sorted position) -- seed of the next search
  float* sqd;      // [src_npad]
  double* mahal;   // [src_npad*6]
  double* partials;  // [nblocks * NRED]
  double thr2;     // max correspondence distance, squared
  int use_seed;
};

enum { MODE_LINEARIZE = 0, MODE_FITNESS = 1 };
// r2 experiment (per-lane need masks + one warp-wide visit per (query, tile) pair): exact, but 72.6 us per pair-pass against 38.5 us
// for the shared walk -- a query needs 4.1 tiles on average, not the 1-2 the idea counted on (132 visits per warp-pass).  Off.
#ifndef MR_NN1_WPQ
#define MR_NN1_WPQ 0
#endif
constexpr int NN1_STAGE = 16;  // candidate tiles staged in shared memory per warp (8 KB); later ones are read from L2
// Packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 work on two floats in a 64-bit register pair; every element is the same IEEE
// round-to-nearest operation as its scalar form, so a distance computed this way equals dist2() bit for bit).
#ifndef MR_NN1_X2
#define MR_NN1_X2 1
#endif
#ifndef MR_NN1_NEEDTEST
#define MR_NN1_NEEDTEST 1
#endif
#ifndef MR_NN1_KEY32
#define MR_NN1_KEY32 1
#endif
#ifndef MR_NN1_PREFETCH
#define MR_NN1_PREFETCH 0
#endif
#ifndef MR_NN1_SCAN_UNROLL
#define MR_NN1_SCAN_UNROLL 4
#endif
__device__ __forceinline__ u64 f2_pack(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 f2_sub(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// dist2() of one query against two points at once: fma(dz,dz, fma(dy,dy, dx*dx)) per element
__device__ __forceinline__ u64 f2_dist2(u64 qx2, u64 qy2, u64 qz2, u64 px2, u64 py2, u64 pz2) {
  const u64 dx = f2_sub(qx2, px2), dy = f2_sub(qy2, py2), dz = f2_sub(qz2, pz2);
  return f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
}

// Warp-cooperative exact NN-1 of ONE query (all lanes hold the same query); the 32 lanes evaluate the 32 points of a
// candidate tile in parallel.  (key0,pos0) is a known candidate (seed) or (KEY_NONE,-1).  Returns the best (key,pos).
__device__ __forceinline__ void warp_query_nn1(const PairParams& P, float qx, float qy, float qz, u64 key0, int pos0, int seed_tile,
                                               float cap, int lane, unsigned long long& tiles_done, u64& out_key, int& out_pos) {
  const int ntp = P.tgt_ntiles_pad, nsp = P.tgt_nsuper_pad;
  u64 bkey = lane == 0 ? key0 : KEY_NONE;
  int bp = lane == 0 ? pos0 : -1;
  float rad = fminf(key_dist(key0), cap);

  auto eval_tile = [&](const float4 c, int t) {
    const float d = dist2(qx, qy, qz, c.x, c.y, c.z);
    const u64 ckey = make_key(d, __float_as_int(c.w));
    if (ckey < bkey) {
      bkey = ckey;
      bp = t * TILE + lane;
    }
    rad = fminf(rad, __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(d))));
    tiles_done++;
  };

  if (seed_tile >= 0) {
    eval_tile(P.tgt_sort[(size_t)seed_tile * TILE + lane], seed_tile);  // e.g. the tile of the previous query's NN
    rad = fminf(rad, cap);
  } else if (pos0 < 0) {
    // no candidate at all: start with the tile whose box is closest to the query
    float mb = INFINITY;
    int ms = 0;
    for (int s0 = lane; s0 < P.tgt_nsuper; s0 += 32) {
      const float b = pt_box_dist2(qx, qy, qz, P.tgt_sb, nsp, s0);
      if (b < mb) {
        mb = b;
        ms = s0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, mb, o);
      const int os = __shfl_xor_sync(FULL, ms, o);
      if (ob < mb || (ob == mb && os < ms)) {
        mb = ob;
        ms = os;
      }
    }
    int mt = ms * 32 + lane;
    float tbd = (mt < P.tgt_ntiles) ? pt_box_dist2(qx, qy, qz, P.tgt_tb, ntp, mt) : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(FULL, tbd, o);
      const int ot = __shfl_xor_sync(FULL, mt, o);
      if (ob < tbd || (ob == tbd && ot < mt)) {
        tbd = ob;
        mt = ot;
      }
    }
    eval_tile(P.tgt_sort[(size_t)mt * TILE + lane], mt);
    rad = fminf(rad, cap);
  }
#pragma unroll 1
  for (int g0 = 0; g0 < nsp; g0 += 32 * SGROUP) {
    float b1[SGROUP];
#pragma unroll
    for (int u = 0; u < SGROUP; u++) {
      const int s1 = g0 + u * 32 + lane;
      b1[u] = (s1 < P.tgt_nsuper) ? pt_box_dist2(qx, qy, qz, P.tgt_sb, nsp, s1) : INFINITY;
    }
#pragma unroll
    for (int u = 0; u < SGROUP; u++) {
      unsigned ms = __ballot_sync(FULL, b1[u] * CULL_SAFETY <= rad);
      while (ms) {
        const int js = __ffs(ms) - 1;
        ms &= ms - 1;
        if (__shfl_sync(FULL, b1[u], js) * CULL_SAFETY > rad) continue;
        const int sidx = g0 + u * 32 + js;
        const int t = sidx * 32 + lane;
        const float bt = pt_box_dist2(qx, qy, qz, P.tgt_tb, ntp, t);
        unsigned mt = __ballot_sync(FULL, t < P.tgt_ntiles && bt * CULL_SAFETY <= rad);
        while (mt) {
          float4 cc[PF];
          float bb[PF];
          int tt[PF];
#pragma unroll
          for (int f = 0; f < PF; f++) {
            bb[f] = INFINITY;
            tt[f] = 0;
            if (mt) {
              const int jt = __ffs(mt) - 1;
              mt &= mt - 1;
              bb[f] = __shfl_sync(FULL, bt, jt);
              tt[f] = sidx * 32 + jt;
              cc[f] = P.tgt_sort[(size_t)tt[f] * TILE + lane];
            }
          }
#pragma unroll
          for (int f = 0; f < PF; f++)
            if (bb[f] * CULL_SAFETY <= rad) eval_tile(cc[f], tt[f]);
        }
      }
    }
  }
  // warp arg-min of the per-lane bests (key order == (distance, index) order)
  const unsigned hi = (unsigned)(bkey >> 32), lo = (unsigned)(bkey & 0xffffffffull);
  const unsigned mhi = __reduce_min_sync(FULL, hi);
  const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
  const int win = __ffs(__ballot_sync(FULL, hi == mhi && lo == mlo)) - 1;
  out_pos = __shfl_sync(FULL, bp, win);
  out_key = ((u64)mhi << 32) | mlo;
}

// grid = (ceil(src_ntiles/4), n_active), block = 128.  MODE_LINEARIZE: writes nnpos/sqd.
// MODE_FITNESS: per-block partial (sum of d <= max_range, count) into partials[.. + 0/1].
//
// Hybrid exact search per warp (= 32 spatially coherent source points):
//  phase A  (seeded passes) every lane takes its seed = NN of the previous iteration.  Lanes whose seed distance is within
//           4x (squared) of the warp's 24th smallest are "coherent": they share ONE candidate-tile walk.  Tiles whose box
//           is within R0 = max coherent best of the warp's bounding box are collected (<= 64), visited NEAREST FIRST,
//           staged in shared memory, and each lane scans the 32 points for its own query (lane-per-query brute force,
//           ~9 instructions per pair, no warp reductions inside).  The walk stops when the nearest remaining tile is
//           farther than every coherent best; a lane is final unless the list overflowed below its best.
//  phase B  all other lanes (isolated points, overflow leftovers, and every lane in the unseeded first pass) run the
//           warp-cooperative single-query search, so one far outlier cannot inflate the walk of the other 31 queries.
// (no minimum-blocks hint: `__launch_bounds__(128, 1)` lets ptxas take 93 registers and costs 14 %; measured in gpurun_out/r2u)
template <int MODE>
__global__ void __launch_bounds__(128) k_nn1_search(const PairParams* __restrict__ plist, const int* __restrict__ active,
                                                    const double* __restrict__ Tbuf, double max_range,
                                                    unsigned long long* stat_pairs) {
  const int pid = active[blockIdx.y];
  const PairParams P = plist[pid];
  __shared__ double s_red[4][2];
#if MR_NN1_X2
  __shared__ __align__(16) float s_tx[4][TILE], s_ty[4][TILE], s_tz[4][TILE];  // the staged target tile, one coordinate per array
  __shared__ __align__(16) int s_tw[4][TILE];                                  // ... and its original indices
#else
  __shared__ float4 s_tile[4][TILE];
#endif
  __shared__ u64 s_cand[4][CAND];
#if MR_NN1_WPQ
  __shared__ __align__(16) float4 s_pts[4][NN1_STAGE][TILE];  // the first NN1_STAGE candidate tiles of the warp, staged once
  __shared__ float s_cbox[4][CAND][6];                        // candidate boxes (for the per-lane need test)
#endif
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // MODE_LINEARIZE has no block-level step and may be launched with ONE warp per block (a block's slot is then freed when
  // its warp is done instead of when the slowest of four is); MODE_FITNESS reduces over the four warps of a block
  const int wpb = blockDim.x >> 5;
  const int qt = blockIdx.x * wpb + wib;
  if (blockIdx.x * wpb >= P.src_ntiles) return;  // block-uniform
  const bool wvalid = qt < P.src_ntiles;
  const double* Td = Tbuf + (size_t)pid * 12;
  float res_d = INFINITY;
  int res_pos = -1;
  bool qvalid = false;
  if (wvalid && P.tgt_ntiles > 0) {
    float Tf[12];
#pragma unroll
    for (int i = 0; i < 12; i++) Tf[i] = (float)Td[i];
    const size_t s = (size_t)qt * TILE + lane;
    const float4 p = P.src_sort[s];
    qvalid = __float_as_int(p.w) != PAD_IDX;
    const float qx = xform_row(Tf + 0, p.x, p.y, p.z);
    const float qy = xform_row(Tf + 4, p.x, p.y, p.z);
    const float qz = xform_row(Tf + 8, p.x, p.y, p.z);
    const int ntp = P.tgt_ntiles_pad, nsp = P.tgt_nsuper_pad;
    const float cap = (MODE == MODE_LINEARIZE) ? (float)fmin(P.thr2, (double)FLT_MAX) : (float)fmin(max_range, (double)FLT_MAX);
    unsigned long long tiles_done = 0;
    u64 bkey = KEY_NONE;
    int bp = -1;

    // lane-per-query scan of one target tile (staged in shared memory)
#if MR_NN1_X2
    // Four points per step on the packed-FP32 pipe: 12 two-wide operations give the four squared distances, their minimum decides
    // whether any of them can matter (m <= best distance, as `ckey < bkey` needs), and only then are indices loaded and keys
    // compared.  The winner is the smallest (distance, index) key among the four, i.e. exactly what four sequential updates leave.
    const u64 qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
    auto scan_tile = [&](const float4 c, int t) {
      __syncwarp();
      s_tx[wib][lane] = c.x;
      s_ty[wib][lane] = c.y;
      s_tz[wib][lane] = c.z;
      s_tw[wib][lane] = __float_as_int(c.w);
      __syncwarp();
      float bd = key_dist(bkey);
      MR_UNROLL(MR_NN1_SCAN_UNROLL)
      for (int g = 0; g < TILE / 4; g++) {
        const ulonglong2 X = reinterpret_cast<const ulonglong2*>(s_tx[wib])[g], Y = reinterpret_cast<const ulonglong2*>(s_ty[wib])[g],
                         Z = reinterpret_cast<const ulonglong2*>(s_tz[wib])[g];
        float d0, d1, d2, d3;
        f2_unpack(f2_dist2(qx2, qy2, qz2, X.x, Y.x, Z.x), d0, d1);
        f2_unpack(f2_dist2(qx2, qy2, qz2, X.y, Y.y, Z.y), d2, d3);
        const float m = fminf(fminf(d0, d1), fminf(d2, d3));
        if (m <= bd) {
          const int4 W = reinterpret_cast<const int4*>(s_tw[wib])[g];
          // unsigned order of the index bit patterns, as make_key() packs them
          const unsigned i0 = d0 == m ? (unsigned)W.x : 0xffffffffu, i1 = d1 == m ? (unsigned)W.y : 0xffffffffu,
                         i2 = d2 == m ? (unsigned)W.z : 0xffffffffu, i3 = d3 == m ? (unsigned)W.w : 0xffffffffu;
          const unsigned im = min(min(i0, i1), min(i2, i3));
          const u64 ckey = ((u64)__float_as_uint(m) << 32) | im;
          if (ckey < bkey) {
            bkey = ckey;
            bd = m;
            bp = t * TILE + g * 4 + (i0 == im ? 0 : i1 == im ? 1 : i2 == im ? 2 : 3);
          }
        }
      }
      tiles_done++;
    };
#else
    auto scan_tile = [&](const float4 c, int t) {
      __syncwarp();
      s_tile[wib][lane] = c;
      __syncwarp();
      float bd = key_dist(bkey);
#pragma unroll 8
      for (int j = 0; j < TILE; j++) {
        const float4 c4 = s_tile[wib][j];
        const float d = dist2(qx, qy, qz, c4.x, c4.y, c4.z);
        if (d <= bd) {
          const u64 ckey = make_key(d, __float_as_int(c4.w));
          if (ckey < bkey) {
            bkey = ckey;
            bp = t * TILE + j;
            bd = d;
          }
        }
      }
      tiles_done++;
    };
#endif

    // ---- phase A: seeds ----
    int sp = -1;
    if (P.use_seed) {
      sp = P.nnpos[s];
    } else if (qvalid && P.tgt_n > 0) {
      // first pass: the target is sorted by Morton key, so a binary search for the query's key lands next to points that
      // are (usually) spatially close -- a seed that costs ~17 probes per lane instead of a full per-query search
      const float mnx = P.tgt_qp[0], mny = P.tgt_qp[1], mnz = P.tgt_qp[2], sc = P.tgt_qp[3];
      const unsigned long long kq = morton48(qx, qy, qz, mnx, mny, mnz, sc) & KEY48_SORTED_MASK;
      int lo = 0, hi = P.tgt_n;  // first position whose key is >= kq
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P.tgt_keys[mid] < kq) lo = mid + 1;
        else hi = mid;
      }
      sp = min(lo, P.tgt_n - 1);
      if (sp > 0) {  // the neighbour on the other side of the key may be closer
        const float4 a = P.tgt_sort[sp - 1], b = P.tgt_sort[sp];
        if (make_key(dist2(qx, qy, qz, a.x, a.y, a.z), __float_as_int(a.w)) < make_key(dist2(qx, qy, qz, b.x, b.y, b.z), __float_as_int(b.w))) sp--;
      }
    }
    if (qvalid && sp >= 0) {
      const float4 c = P.tgt_sort[sp];
      bkey = make_key(dist2(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
      bp = sp;
    }
    const unsigned vmask = __ballot_sync(FULL, qvalid);
    unsigned cmask = 0;
    {
      // coherent set: best within 4x of the warp's 24th smallest (squared distances)
      const unsigned mybits = (qvalid && sp >= 0) ? __float_as_uint(key_dist(bkey)) : 0xffffffffu;
      const unsigned sorted = warp_sort_u32(mybits, lane);
      const float d24 = __uint_as_float(__shfl_sync(FULL, sorted, 23));
      const float thr = fmaxf(4.0f * d24, 1e-6f);
      bool coherent = qvalid && sp >= 0 && key_dist(bkey) <= thr;
      cmask = __ballot_sync(FULL, coherent);
      if (cmask) {
        float wb[6];  // bounding box of the coherent queries
        wb[0] = ord2f(__reduce_min_sync(FULL, coherent ? f2ord(qx) : 0xffffffffu));
        wb[1] = ord2f(__reduce_min_sync(FULL, coherent ? f2ord(qy) : 0xffffffffu));
        wb[2] = ord2f(__reduce_min_sync(FULL, coherent ? f2ord(qz) : 0xffffffffu));
        wb[3] = ord2f(__reduce_max_sync(FULL, coherent ? f2ord(qx) : 0u));
        wb[4] = ord2f(__reduce_max_sync(FULL, coherent ? f2ord(qy) : 0u));
        wb[5] = ord2f(__reduce_max_sync(FULL, coherent ? f2ord(qz) : 0u));
        const float R0 = fminf(__uint_as_float(__reduce_max_sync(FULL, coherent ? __float_as_uint(key_dist(bkey)) : 0u)), cap);
        // collect the candidate tiles (box within R0 of the warp box), at most CAND of them
        int cnt = 0;
        float dropped = INFINITY;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
        for (int g0 = 0; g0 < nsp; g0 += 32) {
          const int s1 = g0 + lane;
          const float b1 = (s1 < P.tgt_nsuper) ? box_box_dist2(wb, P.tgt_sb, nsp, s1) : INFINITY;
          unsigned ms = __ballot_sync(FULL, b1 * CULL_SAFETY <= R0);
          while (ms) {
            const int js = __ffs(ms) - 1;
            ms &= ms - 1;
            const int t = (g0 + js) * 32 + lane;
#if MR_NN1_WPQ
            float tbx[6];
#pragma unroll
            for (int c6 = 0; c6 < 6; c6++) tbx[c6] = P.tgt_tb[(size_t)c6 * ntp + t];  // t < ntiles_pad always
            const float gx_ = fmaxf(0.0f, fmaxf(tbx[0] - wb[3], wb[0] - tbx[3])), gy_ = fmaxf(0.0f, fmaxf(tbx[1] - wb[4], wb[1] - tbx[4])),
                        gz_ = fmaxf(0.0f, fmaxf(tbx[2] - wb[5], wb[2] - tbx[5]));
            const float bt = gx_ * gx_ + gy_ * gy_ + gz_ * gz_;
#else
            const float bt = box_box_dist2(wb, P.tgt_tb, ntp, t);
#endif
            const bool pass = t < P.tgt_ntiles && bt * CULL_SAFETY <= R0;
            const unsigned mt = __ballot_sync(FULL, pass);
            if (pass) {
              const int pos = cnt + __popc(mt & lt);
              if (pos < CAND) {
                s_cand[wib][pos] = ((u64)__float_as_uint(bt) << 32) | (unsigned)t;
#if MR_NN1_WPQ
#pragma unroll
                for (int c6 = 0; c6 < 6; c6++) s_cbox[wib][pos][c6] = tbx[c6];
#endif
              } else dropped = fminf(dropped, bt);
            }
            cnt += __popc(mt);
          }
        }
        const bool overflow = cnt > CAND;
        const float dropped_min = overflow ? warp_min(dropped) : INFINITY;
        const int n = min(cnt, CAND);
        __syncwarp();
#if MR_NN1_WPQ
        // r2: every coherent lane needs only the one to three candidate tiles whose box lies within ITS OWN seed distance, so the
        // shared lane-per-query walk of round 1 (every lane scans every tile of the warp's list: 11.7 tiles x 32 points, of which
        // one in six (lane, tile) pairs could matter) is replaced by: stage the tiles once, let every lane mark the tiles it
        // needs, then serve the (query, tile) pairs one at a time with the whole warp (one point per lane, warp arg-min).
        const int nst = min(n, NN1_STAGE);
        for (int c = 0; c < nst; c++) s_pts[wib][c][lane] = P.tgt_sort[(size_t)(unsigned)(s_cand[wib][c] & 0xffffffffull) * TILE + lane];
        unsigned need0 = 0, need1 = 0;
        {
          const float bestd = key_dist(bkey);
#pragma unroll 2
          for (int c = 0; c < n; c++) {
            const float* bx = s_cbox[wib][c];
            const float gx = fmaxf(0.0f, fmaxf(bx[0] - qx, qx - bx[3])), gy = fmaxf(0.0f, fmaxf(bx[1] - qy, qy - bx[4])),
                        gz = fmaxf(0.0f, fmaxf(bx[2] - qz, qz - bx[5]));
            const bool nd = coherent && (gx * gx + gy * gy + gz * gz) * CULL_SAFETY <= fminf(bestd, cap);
            if (c < 32) need0 |= nd ? (1u << c) : 0u;
            else need1 |= nd ? (1u << (c - 32)) : 0u;
          }
        }
        __syncwarp();
        unsigned todo = __ballot_sync(FULL, (need0 | need1) != 0u);
        while (todo) {
          const int qi = __ffs(todo) - 1;
          todo &= todo - 1;
          unsigned m0 = __shfl_sync(FULL, need0, qi), m1 = __shfl_sync(FULL, need1, qi);
          const float ox = __shfl_sync(FULL, qx, qi), oy = __shfl_sync(FULL, qy, qi), oz = __shfl_sync(FULL, qz, qi);
          u64 obest = __shfl_sync(FULL, bkey, qi);
          int obp = __shfl_sync(FULL, bp, qi);
          while (m0 | m1) {
            int c;
            if (m0) {
              c = __ffs(m0) - 1;
              m0 &= m0 - 1;
            } else {
              c = 32 + __ffs(m1) - 1;
              m1 &= m1 - 1;
            }
            const int t = (int)(unsigned)(s_cand[wib][c] & 0xffffffffull);
            const float4 c4 = c < NN1_STAGE ? s_pts[wib][c][lane] : P.tgt_sort[(size_t)t * TILE + lane];
            const float d = dist2(ox, oy, oz, c4.x, c4.y, c4.z);
            const unsigned hi = __float_as_uint(d), idx = (unsigned)__float_as_int(c4.w);
            const unsigned mhi = __reduce_min_sync(FULL, hi);
            const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? idx : 0xffffffffu);
            const u64 cmin = ((u64)mhi << 32) | mlo;
            if (cmin < obest) {
              obest = cmin;
              obp = t * TILE + (__ffs(__ballot_sync(FULL, hi == mhi && idx == mlo)) - 1);
            }
            tiles_done++;
          }
          if (lane == qi) {
            bkey = obest;
            bp = obp;
          }
        }
#else
        // Nearest-first walk over the candidate tiles: stop as soon as the nearest remaining tile is farther than every coherent best.
        // One 32-bit key per candidate, (box distance bits & ~63) | slot -- the distance rounded DOWN by < 2^-17 relative, which only
        // makes the stop test more conservative -- so one warp min-reduction picks the next tile.  The tile after the current one is
        // picked, and its row fetched, BEFORE the current one is scanned (the order does not depend on the scan), so the global load
        // overlaps ~180 instructions of scanning instead of stalling the warp at the head of every tile.
#if MR_NN1_KEY32
        constexpr unsigned NONE32 = 0xffffffffu;
        unsigned k0 = lane < n ? (((unsigned)(s_cand[wib][lane] >> 32) & ~63u) | (unsigned)lane) : NONE32;
        unsigned k1 = lane + 32 < n ? (((unsigned)(s_cand[wib][lane + 32] >> 32) & ~63u) | (unsigned)(lane + 32)) : NONE32;
        auto pick = [&](unsigned& dbits, int& t) -> bool {  // removes the nearest remaining candidate from the list
          const unsigned key = __reduce_min_sync(FULL, min(k0, k1));
          if (key == NONE32) return false;
          const int slot = (int)(key & 63u);
          if (slot == lane) k0 = NONE32;
          if (slot == lane + 32) k1 = NONE32;
          t = (int)(unsigned)(s_cand[wib][slot] & 0xffffffffull);
          dbits = key & ~63u;
          return true;
        };
#else
        u64 e0 = lane < n ? s_cand[wib][lane] : KEY_NONE;
        u64 e1 = lane + 32 < n ? s_cand[wib][lane + 32] : KEY_NONE;
        auto pick = [&](unsigned& dbits, int& t) -> bool {
          const u64 m = e0 < e1 ? e0 : e1;
          const unsigned hi = (unsigned)(m >> 32);
          const unsigned mhi = __reduce_min_sync(FULL, hi);
          if (mhi >= 0x7f800000u) return false;  // list exhausted
          const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? (unsigned)(m & 0xffffffffull) : 0xffffffffu);
          const u64 pk = ((u64)mhi << 32) | mlo;
          if (e0 == pk) e0 = KEY_NONE;
          else if (e1 == pk) e1 = KEY_NONE;
          dbits = mhi;
          t = (int)mlo;
          return true;
        };
#endif
        auto needed = [&](int t) -> bool {
#if MR_NN1_NEEDTEST
          // The list was built from the box of ALL coherent queries; a tile near one corner of that box is often out of reach of
          // every single query.  Scan it only if some coherent lane's own point-to-box distance is within its current best.
          const float pbd = pt_box_dist2(qx, qy, qz, P.tgt_tb, ntp, t);
          return __any_sync(FULL, coherent && pbd * CULL_SAFETY <= fminf(key_dist(bkey), cap));
#else
          return true;
#endif
        };
#if MR_NN1_PREFETCH
        unsigned nkey;
        int nt = 0;
        bool have = pick(nkey, nt);
        float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) cn = P.tgt_sort[(size_t)nt * TILE + lane];
        while (have) {
          const unsigned ckey = nkey;
          const int t = nt;
          const float4 c = cn;
          const float R = fminf(__uint_as_float(__reduce_max_sync(FULL, coherent ? __float_as_uint(key_dist(bkey)) : 0u)), cap);
          if (__uint_as_float(ckey) * CULL_SAFETY > R) break;
          have = pick(nkey, nt);
          if (have) cn = P.tgt_sort[(size_t)nt * TILE + lane];
          if (needed(t)) scan_tile(c, t);
        }
#else
        while (true) {
          unsigned dbits;
          int t;
          if (!pick(dbits, t)) break;
          const float R = fminf(__uint_as_float(__reduce_max_sync(FULL, coherent ? __float_as_uint(key_dist(bkey)) : 0u)), cap);
          if (__uint_as_float(dbits) * CULL_SAFETY > R) break;
          if (needed(t)) scan_tile(P.tgt_sort[(size_t)t * TILE + lane], t);
        }
#endif
#endif
        // with an overflowing candidate list some tiles at distance >= dropped_min were never looked at:
        // coherent lanes whose best is not provably closer go to phase B
        if (overflow) {
          coherent = coherent && key_dist(bkey) < dropped_min * CULL_SAFETY;
          cmask = __ballot_sync(FULL, coherent);
        }
      }
    }
    // ---- phase B: the remaining queries, one at a time (seeded with whatever is known for them) ----
    unsigned omask = vmask & ~cmask;
    int prev_tile = -1;
    while (omask) {
      const int qi = __ffs(omask) - 1;
      omask &= omask - 1;
      const float ox = __shfl_sync(FULL, qx, qi), oy = __shfl_sync(FULL, qy, qi), oz = __shfl_sync(FULL, qz, qi);
      const u64 k0 = __shfl_sync(FULL, bkey, qi);
      const int p0 = __shfl_sync(FULL, bp, qi);
      u64 ok;
      int op;
      // without any candidate (first pass) the tile of the previous query's NN is the seed
      warp_query_nn1(P, ox, oy, oz, k0, p0, p0 < 0 ? prev_tile : -1, cap, lane, tiles_done, ok, op);
      prev_tile = op >= 0 ? (op >> 5) : -1;
      if (lane == qi) {
        bkey = ok;
        bp = op;
      }
    }
    res_d = key_dist(bkey);
    res_pos = bp;
    if (stat_pairs && lane == 0) {
      atomicAdd(stat_pairs + 1, tiles_done * 32ull);
      atomicMax(stat_pairs + 3, tiles_done);
    }
    if (MODE == MODE_LINEARIZE && qvalid) {
      P.nnpos[s] = res_pos;
      P.sqd[s] = res_d;
    }
  } else if (wvalid && MODE == MODE_LINEARIZE) {
    const size_t s = (size_t)qt * TILE + lane;
    qvalid = __float_as_int(P.src_sort[s].w) != PAD_IDX;
    if (qvalid) {  // empty target
      P.nnpos[s] = -1;
      P.sqd[s] = INFINITY;
    }
  }
  if (MODE == MODE_FITNESS) {
    double a0 = 0.0, a1 = 0.0;
    if (wvalid && qvalid && res_pos >= 0 && (double)res_d <= max_range) {
      a0 = (double)res_d;
      a1 = 1.0;
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      s_red[wib][0] = a0;
      s_red[wib][1] = a1;
    }
    __syncthreads();
    if (threadIdx.x < 2)
      P.partials[(size_t)blockIdx.x * NRED + threadIdx.x] =
          ((s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + s_red[2][threadIdx.x]) + s_red[3][threadIdx.x];
  }
}

#ifndef MR_LIN_SMEMRED
#define MR_LIN_SMEMRED 1
#endif
// Mahalanobis + linearisation terms, one thread per SORTED source point.  grid = (ceil(src_ntiles/4), n_active)
__global__ void __launch_bounds__(128) k_linearize(const PairParams* __restrict__ plist, const int* __restrict__ active,
                                                   const double* __restrict__ Tbuf) {
  const int pid = active[blockIdx.y];
  const PairParams P = plist[pid];
  __shared__ double s_red[4][NRED];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (blockIdx.x * 4 >= P.src_ntiles) return;  // block-uniform
  const int qt = blockIdx.x * 4 + wib;
  const double* Td = Tbuf + (size_t)pid * 12;
  double acc[NRED];
#pragma unroll
  for (int i = 0; i < NRED; i++) acc[i] = 0.0;
  if (qt < P.src_ntiles) {
    const size_t s = (size_t)qt * TILE + lane;
    const float4 p = P.src_sort[s];
    if (__float_as_int(p.w) != PAD_IDX) {
      const int bpos = P.nnpos[s];
      const int corr = (bpos >= 0 && (double)P.sqd[s] < P.thr2) ? bpos : -1;
      P.corr[s] = corr;
      if (corr >= 0) {
        double R[9], tt[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          R[3 * r] = Td[4 * r]; R[3 * r + 1] = Td[4 * r + 1]; R[3 * r + 2] = Td[4 * r + 2];
          tt[r] = Td[4 * r + 3];
        }
        const double* ca = P.src_cov + s * 6;
        const double* cb = P.tgt_cov + (size_t)corr * 6;
        const double CA[9] = {ca[0], ca[1], ca[2], ca[1], ca[3], ca[4], ca[2], ca[4], ca[5]};
        const double CB[9] = {cb[0], cb[1], cb[2], cb[1], cb[3], cb[4], cb[2], cb[4], cb[5]};
        double RCA[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) RCA[3 * i + j] = R[3 * i] * CA[j] + R[3 * i + 1] * CA[3 + j] + R[3 * i + 2] * CA[6 + j];
        double RCR[9];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = i; j < 3; j++) {
            double sij = RCA[3 * i] * R[3 * j] + RCA[3 * i + 1] * R[3 * j + 1] + RCA[3 * i + 2] * R[3 * j + 2];
            RCR[3 * i + j] = sij + CB[3 * i + j];
          }
        RCR[3] = RCR[1]; RCR[6] = RCR[2]; RCR[7] = RCR[5];
        double M[9];
        inverse3(RCR, M);
        double* mo = P.mahal + s * 6;
        mo[0] = M[0]; mo[1] = M[1]; mo[2] = M[2]; mo[3] = M[4]; mo[4] = M[5]; mo[5] = M[8];

        const float4 pb = P.tgt_sort[corr];
        const double a0 = (double)p.x, a1 = (double)p.y, a2 = (double)p.z;
        double ta[3];
#pragma unroll
        for (int r = 0; r < 3; r++) ta[r] = R[3 * r] * a0 + R[3 * r + 1] * a1 + R[3 * r + 2] * a2 + tt[r];
        const double e[3] = {(double)pb.x - ta[0], (double)pb.y - ta[1], (double)pb.z - ta[2]};
        double Me[3];
#pragma unroll
        for (int r = 0; r < 3; r++) Me[r] = M[3 * r] * e[0] + M[3 * r + 1] * e[1] + M[3 * r + 2] * e[2];
        acc[27] = e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
        const double J[3][6] = {{0, -ta[2], ta[1], -1, 0, 0}, {ta[2], 0, -ta[0], 0, -1, 0}, {-ta[1], ta[0], 0, 0, 0, -1}};
        double MJ[3][6];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) MJ[r][c] = M[3 * r] * J[0][c] + M[3 * r + 1] * J[1][c] + M[3 * r + 2] * J[2][c];
        int o = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
          for (int c = r; c < 6; c++) acc[o++] = J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c] + J[2][r] * MJ[2][c];
        }
#pragma unroll
        for (int r = 0; r < 6; r++) acc[21 + r] = J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2];
      }
    }
  }
  // deterministic reduction: lanes -> warp -> block (fixed order) -> one partial row per block
#if MR_LIN_SMEMRED
  // through shared memory: 28 stores per thread, then thread (w, c) adds the 32 values of column c of warp w in a fixed order
  // (four interleaved partial sums) -- ~90 instructions per warp instead of the ~450 of 28 64-bit shuffle butterflies
  __shared__ double s_acc[4][NRED][33];
#pragma unroll
  for (int i = 0; i < NRED; i++) s_acc[wib][i][lane] = acc[i];
  __syncthreads();
  if (threadIdx.x < 4 * NRED) {
    const int w = threadIdx.x / NRED, c = threadIdx.x - w * NRED;
    const double* col = s_acc[w][c];
    double u0 = col[0], u1 = col[8], u2 = col[16], u3 = col[24];
#pragma unroll
    for (int l = 1; l < 8; l++) {
      u0 += col[l];
      u1 += col[8 + l];
      u2 += col[16 + l];
      u3 += col[24 + l];
    }
    s_red[w][c] = (u0 + u1) + (u2 + u3);
  }
  __syncthreads();
#else
#pragma unroll
  for (int i = 0; i < NRED; i++) {
    double v = warp_sum(acc[i]);
    if (lane == 0) s_red[wib][i] = v;
  }
  __syncthreads();
#endif
  if (threadIdx.x < NRED) {
    double v = ((s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + s_red[2][threadIdx.x]) + s_red[3][threadIdx.x];
    P.partials[(size_t)blockIdx.x * NRED + threadIdx.x] = v;
  }
}

// compute_error: fast_gicp_impl.hpp:214-236.  Streams src point, corr, target point, M. grid = (nblk, n_active)
__global__ void __launch_bounds__(128) k_error(const PairParams* __restrict__ plist, const int* __restrict__ active,
                                               const double* __restrict__ Tbuf) {
  const int pid = active[blockIdx.y];
  const PairParams P = plist[pid];
  __shared__ double s_red[4];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (blockIdx.x * 4 >= P.src_ntiles) return;
  const int qt = blockIdx.x * 4 + wib;
  const double* Td = Tbuf + (size_t)pid * 12;
  double err = 0.0;
  if (qt < P.src_ntiles) {
    const size_t s = (size_t)qt * TILE + lane;
    const float4 p = P.src_sort[s];
    const int corr = (__float_as_int(p.w) != PAD_IDX) ? P.corr[s] : -1;
    if (corr >= 0) {
      const float4 pb = P.tgt_sort[corr];
      const double* mo = P.mahal + s * 6;
      const double M[9] = {mo[0], mo[1], mo[2], mo[1], mo[3], mo[4], mo[2], mo[4], mo[5]};
      const double a0 = (double)p.x, a1 = (double)p.y, a2 = (double)p.z;
      double e[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double ta = Td[4 * r] * a0 + Td[4 * r + 1] * a1 + Td[4 * r + 2] * a2 + Td[4 * r + 3];
        e[r] = (double)(r == 0 ? pb.x : (r == 1 ? pb.y : pb.z)) - ta;
      }
      double Me[3];
#pragma unroll
      for (int r = 0; r < 3; r++) Me[r] = M[3 * r] * e[0] + M[3 * r + 1] * e[1] + M[3 * r + 2] * e[2];
      err = e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
    }
  }
  err = warp_sum(err);
  if (lane == 0) s_red[wib] = err;
  __syncthreads();
  if (threadIdx.x == 0) P.partials[(size_t)blockIdx.x * NRED + 27] = ((s_red[0] + s_red[1]) + s_red[2]) + s_red[3];
}

// sums partial rows in a fixed order: grid = n_active blocks, block = 32*ncols threads (warp c owns column col0+c)
__global__ void k_reduce_partials(const PairParams* __restrict__ plist, const int* __restrict__ active, int col0, int ncols,
                                  double* __restrict__ out /* [n_problems][NRED], pinned host or device */) {
  const int pid = active[blockIdx.x];
  const PairParams P = plist[pid];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w >= ncols) return;
  const int nblk = (P.src_ntiles + 3) / 4;
  double v = 0.0;
  for (int r = lane; r < nblk; r += 32) v += P.partials[(size_t)r * NRED + col0 + w];
  v = warp_sum(v);
  if (lane == 0) out[(size_t)pid * NRED + col0 + w] = v;
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_fill_int(int* p, size_t n, int v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// pcl::transformPointCloud(*input_, output, final_transformation_) -- lsq_registration_impl.hpp:78
__global__ void k_transform_out(const float4* __restrict__ porig, int n, const float* __restrict__ Tf12, float* __restrict__ out,
                                int stride) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = porig[i];
  float T[12];
#pragma unroll
  for (int k = 0; k < 12; k++) T[k] = Tf12[k];
  float* o = out + (size_t)i * stride;
  o[0] = xform_row(T + 0, p.x, p.y, p.z);
  o[1] = xform_row(T + 4, p.x, p.y, p.z);
  o[2] = xform_row(T + 8, p.x, p.y, p.z);
}

// user covariances (original order, row-major 4x4 doubles) <-> sorted 6-double storage
__global__ void k_cov_from_user(const double* __restrict__ user16, const float4* __restrict__ psort, int npad, double* __restrict__ cov) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= npad) return;
  int o = __float_as_int(psort[s].w);
  double* c = cov + (size_t)s * 6;
  if (o == PAD_IDX) {
    c[0] = c[3] = c[5] = 1.0;
    c[1] = c[2] = c[4] = 0.0;
    return;
  }
  const double* u = user16 + (size_t)o * 16;
  c[0] = u[0]; c[1] = u[1]; c[2] = u[2]; c[3] = u[5]; c[4] = u[6]; c[5] = u[10];
}
__global__ void k_cov_to_user(const double* __restrict__ cov, const float4* __restrict__ psort, int npad, double* __restrict__ user16) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= npad) return;
  int o = __float_as_int(psort[s].w);
  if (o == PAD_IDX) return;
  const double* c = cov + (size_t)s * 6;
  double* u = user16 + (size_t)o * 16;
  u[0] = c[0]; u[1] = c[1]; u[2] = c[2]; u[3] = 0;
  u[4] = c[1]; u[5] = c[3]; u[6] = c[4]; u[7] = 0;
  u[8] = c[2]; u[9] = c[4]; u[10] = c[5]; u[11] = 0;
  u[12] = u[13] = u[14] = u[15] = 0;
}

// debug: last linearisation state in source ORIGINAL order, corr as target ORIGINAL index
__global__ void k_debug_corr(const float4* __restrict__ src_sort, const float4* __restrict__ tgt_sort, int src_npad,
                             const int* __restrict__ corr, const float* __restrict__ sqd, const double* __restrict__ mahal,
                             int* __restrict__ corr_out, float* __restrict__ sqd_out, double* __restrict__ mahal9_out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= src_npad) return;
  int o = __float_as_int(src_sort[s].w);
  if (o == PAD_IDX) return;
  int c = corr[s];
  if (corr_out) corr_out[o] = c >= 0 ? __float_as_int(tgt_sort[c].w) : -1;
  if (sqd_out) sqd_out[o] = sqd[s];
  if (mahal9_out) {
    double* m = mahal9_out + (size_t)o * 9;
    if (c >= 0) {
      const double* mi = mahal + (size_t)s * 6;
      m[0] = mi[0]; m[1] = mi[1]; m[2] = mi[2]; m[3] = mi[1]; m[4] = mi[3]; m[5] = mi[4]; m[6] = mi[2]; m[7] = mi[4]; m[8] = mi[5];
    } else {
      for (int k = 0; k < 9; k++) m[k] = 0.0;
    }
  }
}

}  // namespace mrslam
