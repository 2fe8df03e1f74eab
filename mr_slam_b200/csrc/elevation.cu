// elevation.cu -- elevation-mapping point process + order-preserving fuse for sm_100a, and the C-ABI
// that replaces the free functions of the reference's libgpu.so.
//
// Reference (paths relative to Mapping/src/elevation_mapping_periodical/elevation_mapping/):
//   cuda/gpu_process.cu:384-455   G_pointsprocess      -> k_elev_process
//   cuda/gpu_process.cu:477-535   G_fuse (O(cells x points), per cell in ascending point index)
//                                 -> stable radix sort by cell + k_elev_fuse_segments (O(points))
//   cuda/gpu_process.cu:198-276   init / clear kernels -> k_elev_fill, k_elev_clear_*
//   cuda/gpu_process.cu:538-545   G_Mapvar_update      -> k_elev_mapvar
//   cuda/gpu_process.cu:938-1252  host wrappers (Init_GPU_elevationmap, Move, Process_points, Fuse,
//                                 Mapvar_update, Map_optmove, Map_closeloop)
//
// HBM layout: the map is SoA like the reference's (elevation, variance, intensity, lowest, traver :
// float[L*L]; colour R,G,B : int[L*L]) in ring-buffer storage order.  Per-frame point arrays are SoA
// device buffers that persist between calls (no cudaMalloc/cudaFree per frame as in the reference).
//
// Arithmetic: built with -fmad=false; every fused multiply-add below is explicit and is the
// contraction nvcc applies to the reference kernels (t=a*b; t=fma(c,d,t) ... left to right), so the
// cell indices are bit-identical to the reference's and to the CPU oracle's.

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "radix_sort.cuh"

namespace mrslam {

// csrc/elev_post.cu
void launch_mapfeature(const float* elevation, const float* variance, const float* intensity, const int* R, const int* G, const int* B,
                       float* traver_map, int L, float res, int sx, int sy, float* o_elev, float* o_var, int* o_R, int* o_G, int* o_B,
                       float* o_rough, float* o_slope, float* o_traver, float* o_int, cudaStream_t st);
void launch_raytracing(float* elevation, const float* variance, const float* traver, float* lowest, int L, int sx, int sy, float sensor_z,
                       float obstacle_thr, int* list, int* count, cudaStream_t st);

struct ElevGeom {
  int L;
  float res;
  float cx, cy;    // central_coordinate
  int sx, sy;      // start_indice
  int row_lo, row_hi;  // STORAGE rows this map instance owns (tile-sharded global grid); [0, L) for a whole map
};

struct ProcParams {
  float T[16];
  double lower, upper;
  float min_r, beam_a, beam_c;
  float sJ[3], rotVar[9], CSBT[9], PCBM[3], BrBS[9];
  int near_filter;
};

__device__ __forceinline__ bool cell_of(const ElevGeom& g, float px, float py, int& ix, int& iy) {
  // gpu_process.cu:309-358
  const float sx = px - g.cx, sy = py - g.cy;
  if (g.L % 2 == 0) {
    ix = (int)((float)(g.L / 2) - sx / g.res);
    iy = (int)((float)(g.L / 2) - sy / g.res);
  } else {
    ix = g.L / 2 - (int)((double)(sx / g.res) + 0.5 * (sx > 0 ? 1 : -1));
    iy = g.L / 2 - (int)((double)(sy / g.res) + 0.5 * (sy > 0 ? 1 : -1));
  }
  return ix >= 0 && ix < g.L && iy >= 0 && iy < g.L;
}

// atomic minimum of a float with ONE hardware atomic instead of a compare-and-swap loop (dozens of points of a scan line land in
// the same cell: the CAS loop of round 1 retried under that contention and made the process kernel latency-bound, 14 points/ns):
// non-negative floats order like signed integers, negative floats like unsigned integers in reverse.
__device__ __forceinline__ void atomic_min_float(float* addr, float val) {
#ifdef MR_ELEV_CAS_MIN  // the round-1 form, kept for A/B measurements
  int* ai = (int*)addr;
  int old = *ai, assumed;
  do {
    assumed = old;
    old = atomicCAS(ai, assumed, __float_as_int(fminf(val, __int_as_float(assumed))));
  } while (assumed != old);
#else
  if (!(val == val)) return;      // NaN never wins (the CAS form kept the old value)
  if (val == 0.0f) val = 0.0f;    // -0.0 would read as INT_MIN below
  if (val >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(val));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(val));
#endif
}

__global__ void k_elev_fill(float* elevation, float* variance, float* intensity, float* lowest, float* traver, int* R, int* G, int* B,
                            int n, int mode) {
  // mode 0: G_Init_map, 1: G_Clear_allmap
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  intensity[i] = 0;
  elevation[i] = -10;
  variance[i] = -10;
  if (mode == 0) lowest[i] = 100;
  traver[i] = -10;
  R[i] = 0;
  G[i] = 0;
  B[i] = 0;
}

// G_Clear_map(start, shift, row_flag): gpu_process.cu:254-276.  A Move clears up to four strips (two axes, each possibly
// wrapping around the ring buffer): they are handled by ONE launch, blockIdx.y = strip.
struct ClearRegions {
  int start[4], shift[4], row_flag[4];
  int count;
};
__global__ void k_elev_clear_region(float* elevation, float* variance, float* intensity, int* R, int* G, int* B, int L, ClearRegions cr) {
  const int r = blockIdx.y;
  const int start = cr.start[r], shift = cr.shift[r], row_flag = cr.row_flag[r];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * shift; i += gridDim.x * blockDim.x) {
    size_t c = row_flag ? (size_t)start * L + i : (size_t)(i / shift) * L + i % shift + start;
    intensity[c] = 0;
    elevation[c] = -10;
    variance[c] = -10;
    R[c] = 0;
    G[c] = 0;
    B[c] = 0;
  }
}

// G_pointsprocess for ONE point.  Outputs -1 everywhere for dropped points.
__device__ __forceinline__ void elev_process_point(float x, float y, float z, const ProcParams& P, const ElevGeom& g, float* __restrict__ lowest,
                                                   int& mi, float& hv, float& xt, float& yt, float& zt) {
  const float* T = P.T;
  const float height = fmaf(T[10], z, fmaf(T[9], y, T[8] * x)) + T[11];
  int flag = 0;
  if (P.near_filter && ((x > -1.5f && x < 1.5f && y > -1.5f && y < 1.5f) || (y > -1.0f && y < 1.0f) || y > 0.0f)) flag = 1;
  mi = -1;
  hv = -1.f; xt = -1.f; yt = -1.f; zt = -1.f;
  if (((double)height > P.lower && (double)height < P.upper) && flag == 0) {
    xt = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3];
    yt = fmaf(T[6], z, fmaf(T[5], y, T[4] * x)) + T[7];
    zt = height;
    const float dist = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
    const float varN = P.min_r * P.min_r;
    const float lat = fmaf(P.beam_a, dist, P.beam_c);
    const float varL = lat * lat;
    float v[3];
#pragma unroll
    for (int r = 0; r < 3; r++) v[r] = fmaf(P.CSBT[3 * r + 2], z, fmaf(P.CSBT[3 * r + 1], y, P.CSBT[3 * r] * x));
    float S[9] = {0.f, -v[2], v[1], v[2], 0.f, -v[0], -v[1], v[0], 0.f};
#pragma unroll
    for (int a = 0; a < 9; a++) S[a] = S[a] + P.BrBS[a];
    float rJ[3];
#pragma unroll
    for (int c = 0; c < 3; c++) rJ[c] = fmaf(P.PCBM[2], S[6 + c], fmaf(P.PCBM[1], S[3 + c], P.PCBM[0] * S[c]));
    float a1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) a1[c] = fmaf(rJ[2], P.rotVar[6 + c], fmaf(rJ[1], P.rotVar[3 + c], rJ[0] * P.rotVar[c]));
    hv = fmaf(a1[2], rJ[2], fmaf(a1[1], rJ[1], a1[0] * rJ[0]));
    const float sv[3] = {varL, varL, varN};
    float b1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) b1[c] = P.sJ[c] * sv[c];
    hv += fmaf(b1[2], P.sJ[2], fmaf(b1[1], P.sJ[1], b1[0] * P.sJ[0]));
    int ix, iy;
    const bool inside = cell_of(g, xt, yt, ix, iy);
    const int srow = inside ? (ix + g.sx) % g.L : -1;
    if (inside && srow >= g.row_lo && srow < g.row_hi) {  // cells of other tiles belong to another GPU's instance
      mi = srow * g.L + (iy + g.sy) % g.L;
      const int gi = ix * g.L + iy;
      // racy in the reference as well (gpu_process.cu:434-438); map_lowest only feeds the ray-tracing clean-up
      atomic_min_float(&lowest[gi], height);
      if (height == lowest[gi]) lowest[gi] = lowest[gi] + 3 * hv;
    }
  }
}

// keys[i] = map_index or `cells` (sorts last).
__global__ void k_elev_process(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz, int n,
                               ProcParams P, ElevGeom g, float* __restrict__ lowest, int* __restrict__ map_index,
                               float* __restrict__ var, float* __restrict__ x_ts, float* __restrict__ y_ts, float* __restrict__ z_ts,
                               int* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int mi;
  float hv, xt, yt, zt;
  elev_process_point(px[i], py[i], pz[i], P, g, lowest, mi, hv, xt, yt, zt);
  if (map_index) map_index[i] = mi;
  if (var) var[i] = hv;
  if (x_ts) {
    x_ts[i] = xt;
    y_ts[i] = yt;
  }
  if (z_ts) z_ts[i] = zt;
  if (keys) {
    keys[i] = (mi >= 0 && zt != -1.f) ? mi : g.L * g.L;
    vals[i] = i;
  }
}

// ---- whole-batch replay (mrslam_elev_process_fuse_device_batch): every frame of a replay in ONE pass ----------------------
// Frame f's geometry (after its Move) and sensor parameters live in a device table; all frames are processed by one launch
// (blockIdx.y = frame) into concatenated per-point arrays.  vals = frame << 24 | point index; the stable sort by cell then
// leaves every run in (frame, point) order -- the order in which the per-frame path applies its updates.
struct BatchFrame {
  const float *x, *y, *z;
  int n, off;
  ElevGeom g;
  ProcParams P;
};
constexpr int ELEV_GROUP = 64;  // frames per group: one bit per frame in the clear masks

__global__ void k_elev_process_batch(const BatchFrame* __restrict__ frames, int cells, float* __restrict__ lowest, float* __restrict__ ph,
                                     float* __restrict__ pv, int* __restrict__ keys, int* __restrict__ vals) {
  const int f = blockIdx.y;
  const BatchFrame& F = frames[f];  // (staging the table entry in shared memory was measured: slower)
  const int n = F.n, off = F.off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int mi;
    float hv, xt, yt, zt;
    elev_process_point(F.x[i], F.y[i], F.z[i], F.P, F.g, lowest, mi, hv, xt, yt, zt);
    ph[off + i] = zt;
    pv[off + i] = hv;
    keys[off + i] = (mi >= 0 && zt != -1.f) ? mi : cells;
    vals[off + i] = (f << 24) | i;
  }
}

struct BatchOffsets {
  int off[ELEV_GROUP];
};
__global__ void k_elev_gather_batch(const int* __restrict__ order, const int* __restrict__ d_n, BatchOffsets bo, const float* __restrict__ ph,
                                    const float* __restrict__ pv, float2* __restrict__ hv, unsigned char* __restrict__ fr) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *d_n) return;  // the sort dropped the invalid points: the count of the rest lives on the device
  const int v = order[j];
  const int f = v >> 24, pi = bo.off[f] + (v & 0xffffff);
  hv[j] = make_float2(ph[pi], pv[pi]);
  fr[j] = (unsigned char)f;
}

// the strips every Move of the group clears, applied up front (a cell that is cleared at frame f keeps nothing of the frames
// before f, so its earlier points are skipped by the run walk below instead)
struct ClearTab {
  int start, shift, row_flag;
};
__global__ void k_elev_clear_tab(float* elevation, float* variance, float* intensity, int* R, int* G, int* B, int L,
                                 const ClearTab* __restrict__ tab) {
  const ClearTab t = tab[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * t.shift; i += gridDim.x * blockDim.x) {
    size_t c = t.row_flag ? (size_t)t.start * L + i : (size_t)(i / t.shift) * L + i % t.shift + t.start;
    intensity[c] = 0;
    elevation[c] = -10;
    variance[c] = -10;
    R[c] = 0;
    G[c] = 0;
    B[c] = 0;
  }
}

// keys for Fuse() called with caller-provided indices
__global__ void k_elev_keys(const int* __restrict__ idx, const float* __restrict__ h, int n, int cells, int* keys, int* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = idx[i];
  keys[i] = (c >= 0 && c < cells && h[i] != -1.f) ? c : cells;
  vals[i] = i;
}

// gather the per-point inputs into sorted (cell, point-index) order so that the run walk below reads contiguous memory
__global__ void k_elev_gather(const int* __restrict__ order, const int* __restrict__ d_n, const float* __restrict__ ph, const float* __restrict__ pv,
                              const int* __restrict__ cR, const int* __restrict__ cG, const int* __restrict__ cB,
                              const float* __restrict__ inten, float2* __restrict__ hv, int4* __restrict__ col) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= *d_n) return;  // the sort dropped the invalid points; the count of the rest lives on the device
  const int pi = order[j];
  hv[j] = make_float2(ph[pi], pv[pi]);
  if (col) col[j] = make_int4(cR[pi], cG[pi], cB[pi], __float_as_int(inten[pi]));
}

// One thread per run of equal cell keys (points inside a run are in ascending point index because the radix sort is
// stable): the sequential recurrence of G_fuse, gpu_process.cu:477-535, over the gathered (height, variance) stream.
// The recurrence is a loop-carried dependency, so the kernel's duration is the longest run times the latency of one
// update.  Two things keep that latency short: (1) the end of the run is found up front (galloping + binary search on the
// sorted keys), so the loads of the next points do not sit behind the loop-exit test and are issued four at a time;
// (2) the Mahalanobis test |h-e|/sqrt(v) > 5 is decided from |h-e|*rsqrt(v) (relative error < 2^-20) unless that lands
// within 1e-4 of the threshold -- only then the IEEE sqrt and division of the reference expression are evaluated.  The
// decision, and with it every stored value, is identical to the reference's.
__device__ __forceinline__ bool elev_mahal_gt5(float h, float e, float v) {
  const float d = fabsf(h - e);
  const float approx = d * rsqrtf(v);
  if (fabsf(approx - 5.f) > 1e-4f) return approx > 5.f;  // (NaN / inf-inf fall through to the exact expression)
  return d / sqrtf(v) > 5.f;
}

__global__ void k_elev_fuse_segments(const int* __restrict__ keys, const int* __restrict__ d_n, int cells, const float2* __restrict__ hv,
                                     const int4* __restrict__ col, float* __restrict__ elevation, float* __restrict__ variance,
                                     float* __restrict__ intensity, int* __restrict__ R, int* __restrict__ G, int* __restrict__ B,
                                     unsigned long long* n_fused) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *d_n;
  if (i >= n) return;
  const int c = keys[i];
  if (c >= cells) return;
  if (i > 0 && keys[i - 1] == c) return;  // not the head of a run
  // end of the run: first position whose key differs
  int lo = i, step = 1;
  while (lo + step < n && keys[lo + step] == c) {
    lo += step;
    step <<= 1;
  }
  int hi = min(lo + step, n);  // keys[lo] == c, keys[hi] != c (or hi == n)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] == c) lo = mid;
    else hi = mid;
  }
  const int end = hi;
  float e = elevation[c], v = variance[c];
  float ci = 0.f;
  int cr = 0, cg = 0, cb = 0;
  bool colour_written = false;
  for (int j0 = i; j0 < end; j0 += 4) {
    float2 m4[4];
    int4 c4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = min(j0 + u, end - 1);
      m4[u] = hv[j];
      if (col) c4[u] = col[j];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (j0 + u >= end) break;
      const float h = m4[u].x, hvv = m4[u].y;
      bool colour = false;
      int4 cc = make_int4(0, 0, 0, 0);
      if (col) {
        cc = c4[u];
        colour = cc.x != 0 && cc.y != 0 && cc.z != 0 && __int_as_float(cc.w) != 0.f;
      }
      bool wc = false;
      if (e == -10.f) {
        e = h;
        v = hvv;
        wc = colour;
      } else if (elev_mahal_gt5(h, e, v)) {
        if (e < h) {
          e = h;
          v = hvv;
          wc = colour;
        }
      } else {
        const float ne = fmaf(v, h, hvv * e) / (v + hvv);
        const float nv = (hvv * v) / (hvv + v);
        e = ne;
        v = nv;
        wc = colour;
      }
      if (wc) {
        ci = __int_as_float(cc.w);
        cr = cc.x;
        cg = cc.y;
        cb = cc.z;
        colour_written = true;
      }
    }
  }
  elevation[c] = e;
  variance[c] = v;
  if (colour_written) {
    intensity[c] = ci;
    R[c] = cr;
    G[c] = cg;
    B[c] = cb;
  }
  if (n_fused) atomicAdd(n_fused, (unsigned long long)(end - i));
}

// Run walk of a whole group of frames: the per-frame walk above plus, inside a run, (1) the clamp that ends every frame
// (gpu_process.cu:531-532: variance < 1e-4 -> 1e-4) whenever the frame number changes and (2) the Moves: rowclr / colclr hold,
// per storage row / column, one bit per frame whose Move cleared it (k_elev_clear_tab has already applied the clears), so a
// run only applies the points of frames >= the last clear of its cell.  Result: exactly the state the per-frame sequence
// Move, Process_points, Fuse leaves, with one sort and one walk for the whole group.
__global__ void k_elev_fuse_runs_batch(const int* __restrict__ keys, const int* __restrict__ d_n, int cells, int L, const float2* __restrict__ hv,
                                       const unsigned char* __restrict__ fr, const unsigned long long* __restrict__ rowclr,
                                       const unsigned long long* __restrict__ colclr, float* __restrict__ elevation,
                                       float* __restrict__ variance, unsigned long long* n_fused) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *d_n;
  if (i >= n) return;
  const int c = keys[i];
  if (c >= cells) return;
  if (i > 0 && keys[i - 1] == c) return;  // not the head of a run
  int lo = i, step = 1;
  while (lo + step < n && keys[lo + step] == c) {
    lo += step;
    step <<= 1;
  }
  int hi = min(lo + step, n);
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] == c) lo = mid;
    else hi = mid;
  }
  const int end = hi;
  int first_frame = 0;
  if (rowclr) {
    const unsigned long long clr = rowclr[c / L] | colclr[c % L];
    first_frame = clr ? 63 - __clzll((long long)clr) : 0;  // frames before the last clear of this cell leave no trace
  }
  float e = elevation[c], v = variance[c];
  int prev_f = -1;
  for (int j0 = i; j0 < end; j0 += 4) {
    float2 m4[4];
    int f4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = min(j0 + u, end - 1);
      m4[u] = hv[j];
      f4[u] = fr[j];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (j0 + u >= end) break;
      const int f = f4[u];
      if (f < first_frame) continue;
      if (f != prev_f) {
        if (prev_f >= 0 && v < 0.0001f) v = 0.0001f;  // end of frame prev_f
        prev_f = f;
      }
      const float h = m4[u].x, hvv = m4[u].y;
      // (computing the fused estimate ahead of the Mahalanobis test, to overlap its divisions with the test, was measured: the walk
      // got 15 % slower -- 0.60 vs 0.52 ms for the cfg2 replay)
      if (e == -10.f) {
        e = h;
        v = hvv;
      } else if (elev_mahal_gt5(h, e, v)) {
        if (e < h) {
          e = h;
          v = hvv;
        }
      } else {
        const float ne = fmaf(v, h, hvv * e) / (v + hvv);
        const float nv = (hvv * v) / (hvv + v);
        e = ne;
        v = nv;
      }
    }
  }
  elevation[c] = e;
  variance[c] = v;  // the clamp of the last frame follows as k_elev_clamp_var over the whole map
  if (n_fused) atomicAdd(n_fused, (unsigned long long)(end - i));
}

// tail of G_fuse (gpu_process.cu:531-532): every cell, including empty ones
__global__ void k_elev_clamp_var(float* __restrict__ variance, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int i4 = i * 4;
  if (i4 + 3 < n) {
    float4 v = reinterpret_cast<float4*>(variance)[i];
    bool ch = v.x < 0.0001f || v.y < 0.0001f || v.z < 0.0001f || v.w < 0.0001f;
    if (ch) {
      v.x = v.x < 0.0001f ? 0.0001f : v.x;
      v.y = v.y < 0.0001f ? 0.0001f : v.y;
      v.z = v.z < 0.0001f ? 0.0001f : v.z;
      v.w = v.w < 0.0001f ? 0.0001f : v.w;
      reinterpret_cast<float4*>(variance)[i] = v;
    }
  } else {
    for (int j = i4; j < n; j++)
      if (variance[j] < 0.0001f) variance[j] = 0.0001f;
  }
}

__global__ void k_elev_mapvar(float* __restrict__ variance, int n, float upd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && variance[i] != -10.f) variance[i] += upd;
}
__global__ void k_elev_height(float* __restrict__ elevation, int n, float upd) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && elevation[i] != -10.f) elevation[i] += upd;
}

struct Elev {
  ElevGeom g{0, 0.f, 0.f, 0.f, 0, 0};
  float mahal_thr = 0, obstacle_thr = 0, sensor_z = 0;
  DevBuf<float> elevation, variance, intensity, lowest, traver;
  DevBuf<int> R, G, B;
  // per-frame buffers
  DevBuf<float> px, py, pz, pvar, pxt, pyt, pzt, pint, ph;
  DevBuf<int> pidx, pR, pG, pB, keys, keys_alt, vals, vals_alt;
  DevBuf<float2> ghv;
  DevBuf<int4> gcol;
  DevBuf<unsigned long long> counter;
  DevBuf<float> feat;   // Map_feature staging: 6 float planes
  DevBuf<int> feat_i;   // ... and 3 int planes
  // one-pass replay of a group of frames: frame table, clear table, per-row / per-column clear masks (pinned staging + device)
  DevBuf<BatchFrame> d_frames;
  PinBuf<BatchFrame> h_frames;
  DevBuf<ClearTab> d_ctab;
  PinBuf<ClearTab> h_ctab;
  DevBuf<unsigned long long> d_clr;
  PinBuf<unsigned long long> h_clr;
  DevBuf<unsigned char> gfr;
  RadixScratch rsort;
  StageFence batch_fence;
  cudaEvent_t ev_start = nullptr;
  // asynchronous replay: the map-writing half runs on `bstream` instead of the library stream, so that an unrelated batch
  // (a loop-closure storm) can use the library stream meanwhile; every other entry point joins first
  cudaStream_t bstream = nullptr;
  cudaEvent_t ev_done = nullptr;
  bool pending_join = false;
  int join() {
    if (pending_join) {
      MR_CUDA(cudaStreamWaitEvent(rt().stream, ev_done, 0));
      pending_join = false;
    }
    return MRSLAM_OK;
  }
  std::mutex mu;  // Fuse and Mapvar_update may be called from two threads (ElevationMapping.cpp:393-394)

  int cells() const { return g.L * g.L; }
  int reserve_points(int n) {
    size_t m = std::max(n, 1);
    MR_CHECK(px.reserve(m)); MR_CHECK(py.reserve(m)); MR_CHECK(pz.reserve(m));
    MR_CHECK(pvar.reserve(m)); MR_CHECK(pxt.reserve(m)); MR_CHECK(pyt.reserve(m)); MR_CHECK(pzt.reserve(m));
    MR_CHECK(pint.reserve(m)); MR_CHECK(ph.reserve(m));
    MR_CHECK(pidx.reserve(m)); MR_CHECK(pR.reserve(m)); MR_CHECK(pG.reserve(m)); MR_CHECK(pB.reserve(m));
    MR_CHECK(keys.reserve(m)); MR_CHECK(keys_alt.reserve(m)); MR_CHECK(vals.reserve(m)); MR_CHECK(vals_alt.reserve(m));
    MR_CHECK(counter.reserve(2));
    MR_CHECK(ghv.reserve(m));
    MR_CHECK(gcol.reserve(m));
    return MRSLAM_OK;
  }
  void release() {
    for (DevBuf<float>* b : {&elevation, &variance, &intensity, &lowest, &traver, &px, &py, &pz, &pvar, &pxt, &pyt, &pzt, &pint, &ph}) b->release();
    for (DevBuf<int>* b : {&R, &G, &B, &pidx, &pR, &pG, &pB, &keys, &keys_alt, &vals, &vals_alt}) b->release();
    counter.release();
    ghv.release();
    gcol.release();
    feat.release();
    feat_i.release();
    d_frames.release(); h_frames.release(); d_ctab.release(); h_ctab.release(); d_clr.release(); h_clr.release(); gfr.release();
    rsort.hist.release(); rsort.chunk.release(); rsort.count.release();
    if (bstream) cudaStreamSynchronize(bstream);
    if (bstream) cudaStreamDestroy(bstream);
    for (cudaEvent_t e : {ev_start, ev_done})
      if (e) cudaEventDestroy(e);
    bstream = nullptr;
    ev_start = ev_done = nullptr;
    pending_join = false;
  }
  // Returns (through *grew) whether any device buffer was (re)allocated: DevBuf allocates in LIBRARY-stream order, so a replay
  // that runs on the map's own stream must be ordered behind the library stream once after a growth (never in steady state).
  int reserve_batch(size_t total, int nframes, bool* grew) {
    const size_t caps0 = pzt.cap + pvar.cap + keys.cap + keys_alt.cap + vals.cap + vals_alt.cap + ghv.cap + gfr.cap + counter.cap + d_frames.cap +
                         d_ctab.cap + d_clr.cap + rsort.hist.cap + rsort.chunk.cap + rsort.count.cap;
    MR_CHECK(reserve_batch_(total, nframes));
    MR_CHECK(radix_reserve(rsort, (int)std::max<size_t>(total, 1)));
    const size_t caps1 = pzt.cap + pvar.cap + keys.cap + keys_alt.cap + vals.cap + vals_alt.cap + ghv.cap + gfr.cap + counter.cap + d_frames.cap +
                         d_ctab.cap + d_clr.cap + rsort.hist.cap + rsort.chunk.cap + rsort.count.cap;
    *grew = caps1 != caps0;
    return MRSLAM_OK;
  }
  int reserve_batch_(size_t total, int nframes) {
    size_t t = std::max<size_t>(total, 1);
    MR_CHECK(pzt.reserve(t)); MR_CHECK(pvar.reserve(t));
    MR_CHECK(keys.reserve(t)); MR_CHECK(keys_alt.reserve(t)); MR_CHECK(vals.reserve(t)); MR_CHECK(vals_alt.reserve(t));
    MR_CHECK(ghv.reserve(t)); MR_CHECK(gfr.reserve(t));
    MR_CHECK(counter.reserve(2));
    MR_CHECK(d_frames.reserve(nframes)); MR_CHECK(h_frames.reserve(nframes));
    MR_CHECK(d_ctab.reserve(4 * nframes)); MR_CHECK(h_ctab.reserve(4 * nframes));
    MR_CHECK(d_clr.reserve(2 * (size_t)g.L)); MR_CHECK(h_clr.reserve(2 * (size_t)g.L));
    return MRSLAM_OK;
  }
  int ensure_pipeline(int n) {
    (void)n;
    if (!bstream) {
      MR_CUDA(cudaStreamCreateWithFlags(&bstream, cudaStreamNonBlocking));
      for (cudaEvent_t* e : {&ev_start, &ev_done}) MR_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    return MRSLAM_OK;
  }
};

// sort (keys = cell, vals = point index), fuse runs, clamp.  Arrays are device pointers.
static int fuse_sorted(Elev* m, int n, const float* d_h, const float* d_v, const int* d_R, const int* d_G, const int* d_B,
                       const float* d_I, long long* n_fused) {
  cudaStream_t st = rt().stream;
  const int cells = m->cells();
  ProfScope ps(PROF_ELEV_FUSE);
  if (n > 0) {
    int bits = 1;
    while ((1ll << bits) <= cells) bits++;
    // stable sort by cell with the library's own radix sort (csrc/radix_sort.cuh); invalid points (key == cells) are dropped in
    // its first pass, the number of valid ones stays on the device
    unsigned *ks = nullptr, *vs = nullptr;
    const int* d_nv = nullptr;
    MR_CHECK(radix_sort_pairs(m->rsort, (unsigned*)m->keys.p, (unsigned*)m->vals.p, (unsigned*)m->keys_alt.p, (unsigned*)m->vals_alt.p, n, nullptr, 0,
                              ((bits + 7) / 8) * 8, (unsigned)cells, 1, st, &ks, &vs, &d_nv));
    if (n_fused) MR_CUDA(cudaMemsetAsync(m->counter.p, 0, sizeof(unsigned long long), st));
    k_elev_gather<<<(n + 255) / 256, 256, 0, st>>>((const int*)vs, d_nv, d_h, d_v, d_R, d_G, d_B, d_I, m->ghv.p, d_R ? m->gcol.p : nullptr);
    k_elev_fuse_segments<<<(n + 127) / 128, 128, 0, st>>>((const int*)ks, d_nv, cells, m->ghv.p, d_R ? m->gcol.p : nullptr, m->elevation.p,
                                                         m->variance.p, m->intensity.p, m->R.p, m->G.p, m->B.p,
                                                         n_fused ? m->counter.p : m->counter.p + 1);
    count_launch(2);
  }
  k_elev_clamp_var<<<((cells + 3) / 4 + 255) / 256, 256, 0, st>>>(m->variance.p, cells);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  if (n_fused) {
    unsigned long long c = 0;
    if (n > 0) MR_CUDA(cudaMemcpyAsync(&c, m->counter.p, sizeof(c), cudaMemcpyDeviceToHost, st));
    MR_CUDA(cudaStreamSynchronize(st));
    *n_fused = (long long)c;
  }
  return MRSLAM_OK;
}

static void fill_proc(ProcParams& P, const float* T16, double lower, double upper, float min_r, float beam_a, float beam_c,
                      const float* sJ, const float* rotVar, const float* CSBT, const float* PCBM, const float* BrBS, int near_filter) {
  memcpy(P.T, T16, sizeof(float) * 16);
  P.lower = lower; P.upper = upper;
  P.min_r = min_r; P.beam_a = beam_a; P.beam_c = beam_c;
  memcpy(P.sJ, sJ, 12); memcpy(P.rotVar, rotVar, 36); memcpy(P.CSBT, CSBT, 36); memcpy(P.PCBM, PCBM, 12); memcpy(P.BrBS, BrBS, 36);
  P.near_filter = near_filter;
}

// host helpers of gpu_process.cu:890-918, 994-1000
static int index_to_range(int index, int L) {
  if (index < 0) index += ((-index / L) + 1) * L;
  return index % L;
}
static float position_to_range(float p, float shift, float res) {
  int p_index = (int)std::round(p / res);
  int shift_index = (int)std::round(shift / res);
  return (p_index + shift_index) * res;
}
static void index_shift(int* ishift, const float* pshift, float res) {
  for (int i = 0; i < 2; i++) ishift[i] = static_cast<int>(pshift[i] / res + 0.5 * (pshift[i] > 0 ? 1 : -1));
}

}  // namespace mrslam

using namespace mrslam;
#define E(h) (reinterpret_cast<Elev*>(h))
#define REQ(h)                                   \
  do {                                           \
    if (!(h)) {                                  \
      set_error("%s: null handle", __func__);    \
      return MRSLAM_ERR_ARG;                     \
    }                                            \
  } while (0)
#define REQ_INIT(m)                                             \
  do {                                                          \
    if ((m)->g.L <= 0) {                                        \
      set_error("%s: Init_GPU_elevationmap not called", __func__); \
      return MRSLAM_ERR_STATE;                                  \
    }                                                           \
  } while (0)

extern "C" {

mrslam_elev_t* mrslam_elev_create(void) {
  if (ensure_init() != MRSLAM_OK) return nullptr;
  return reinterpret_cast<mrslam_elev_t*>(new Elev());
}
void mrslam_elev_destroy(mrslam_elev_t* h) {
  if (!h) return;
  std::lock_guard<std::mutex> lk(rt().mu);
  E(h)->release();
  delete E(h);
}

int mrslam_elev_init(mrslam_elev_t* h, int length, float resolution, float mahal, float obst) {
  REQ(h);
  MR_CHECK(ensure_init());
  if (length <= 0 || !(resolution > 0.f)) {
    set_error("Init_GPU_elevationmap: bad length/resolution");
    return MRSLAM_ERR_ARG;
  }
  Elev* m = E(h);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  m->g = ElevGeom{length, resolution, 0.f, 0.f, 0, 0, 0, length};
  m->mahal_thr = mahal;
  m->obstacle_thr = obst;
  size_t n = (size_t)length * length;
  for (DevBuf<float>* b : {&m->elevation, &m->variance, &m->intensity, &m->lowest, &m->traver}) MR_CHECK(b->reserve(n + 4));
  for (DevBuf<int>* b : {&m->R, &m->G, &m->B}) MR_CHECK(b->reserve(n));
  k_elev_fill<<<((int)n + 255) / 256, 256, 0, rt().stream>>>(m->elevation.p, m->variance.p, m->intensity.p, m->lowest.p, m->traver.p,
                                                               m->R.p, m->G.p, m->B.p, (int)n, 0);
  count_launch(1);
  MR_CHECK(m->counter.reserve(2));
  MR_CUDA(cudaMemsetAsync(m->counter.p, 0, 2 * sizeof(unsigned long long), rt().stream));
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

static int move_locked(Elev* m, const float* pos, float resolution, int length, float* central_out, int* start_out, float* shift_out,
                       cudaStream_t st);

int mrslam_elev_move(mrslam_elev_t* h, const float* pos, float resolution, int length, float* central_out, int* start_out,
                     float* shift_out) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  return move_locked(m, pos, resolution, length, central_out, start_out, shift_out, rt().stream);
}

// Move (gpu_process.cu:1002-1081) as a host-side plan: the new geometry plus the strips (or everything) to clear.
struct MovePlan {
  bool clear_all = false;
  int nstrips = 0;
  int start[4], shift[4], row_flag[4];
  float ashift[2];
};
static void plan_move(Elev* m, const float* pos, float resolution, int length, MovePlan& mp) {
  int ishift[2];
  float pshift[2];
  m->sensor_z = pos[2];
  float hc[2] = {m->g.cx, m->g.cy};
  int hs[2] = {m->g.sx, m->g.sy};
  pshift[0] = pos[0] - hc[0];
  pshift[1] = pos[1] - hc[1];
  index_shift(ishift, pshift, resolution);
  for (int i = 0; i < 2; i++) mp.ashift[i] = (float)ishift[i] * resolution;
  // strips of one axis never overlap each other; a row strip and a column strip may, but both write the same constants
  auto clear = [&](int start, int ncell, int row) {
    if (ncell * length <= 0) return;
    mp.start[mp.nstrips] = start;
    mp.shift[mp.nstrips] = ncell;
    mp.row_flag[mp.nstrips] = row;
    mp.nstrips++;
  };
  for (int i = 0; i < 2; i++) {
    if (ishift[i] != 0) {
      // |shift| >= length in EITHER direction leaves no cell of the old window inside the new one: clear everything.  (The
      // reference tests only `ishift >= length`; a large negative shift then clears |ishift| > length strips and writes past
      // the end of its six planes -- undefined behaviour there, a clear-all here and in the oracle.)
      if (ishift[i] >= length || ishift[i] <= -length) {
        mp.clear_all = true;
      } else {
        int sign = ishift[i] > 0 ? 1 : -1;
        int startIndex = hs[i] - (sign > 0 ? 1 : 0);
        int endIndex = startIndex + sign - ishift[i];
        int nCells = std::abs(ishift[i]);
        int index = sign < 0 ? startIndex : endIndex;
        index = index_to_range(index, length);
        if (index + nCells <= length) {
          clear(index, nCells, i == 0);
        } else {
          int firstn = length - index;
          clear(index, firstn, i == 0);
          clear(0, nCells - firstn, i == 0);
        }
      }
    }
    hs[i] -= ishift[i];
    hs[i] = index_to_range(hs[i], length);
    hc[i] = position_to_range(hc[i], mp.ashift[i], resolution);
  }
  m->g.cx = hc[0]; m->g.cy = hc[1];
  m->g.sx = hs[0]; m->g.sy = hs[1];
}
static int apply_move_clears(Elev* m, const MovePlan& mp, int length, cudaStream_t st) {
  if (mp.clear_all) {
    int n = length * length;
    k_elev_fill<<<(n + 127) / 128, 128, 0, st>>>(m->elevation.p, m->variance.p, m->intensity.p, m->lowest.p, m->traver.p, m->R.p,
                                                m->G.p, m->B.p, n, 1);
    count_launch(1);
  }
  if (mp.nstrips > 0) {
    ClearRegions cr;
    cr.count = mp.nstrips;
    int max_cnt = 0;
    for (int k = 0; k < mp.nstrips; k++) {
      cr.start[k] = mp.start[k];
      cr.shift[k] = mp.shift[k];
      cr.row_flag[k] = mp.row_flag[k];
      max_cnt = std::max(max_cnt, mp.shift[k] * length);
    }
    k_elev_clear_region<<<dim3((max_cnt + 255) / 256, cr.count), 256, 0, st>>>(m->elevation.p, m->variance.p, m->intensity.p, m->R.p,
                                                                                m->G.p, m->B.p, length, cr);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

static int move_locked(Elev* m, const float* pos, float resolution, int length, float* central_out, int* start_out, float* shift_out,
                       cudaStream_t st) {
  MovePlan mp;
  plan_move(m, pos, resolution, length, mp);
  MR_CHECK(apply_move_clears(m, mp, length, st));
  if (central_out) { central_out[0] = m->g.cx; central_out[1] = m->g.cy; }
  if (start_out) { start_out[0] = m->g.sx; start_out[1] = m->g.sy; }
  if (shift_out) { shift_out[0] = mp.ashift[0]; shift_out[1] = mp.ashift[1]; }
  return MRSLAM_OK;
}

int mrslam_elev_process_points(mrslam_elev_t* h, int* map_index, const float* x, const float* y, const float* z, float* var,
                               float* x_ts, float* y_ts, float* z_ts, const float* T16, int n, double lower, double upper, float min_r,
                               float beam_a, float beam_c, const float* sJ, const float* rotVar, const float* CSBT, const float* PCBM,
                               const float* BrBS) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (n < 0) {
    set_error("Process_points: n < 0");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  cudaStream_t st = rt().stream;
  MR_CHECK(m->reserve_points(n));
  if (n == 0) return MRSLAM_OK;
  size_t b = sizeof(float) * (size_t)n;
  MR_CUDA(cudaMemcpyAsync(m->px.p, x, b, cudaMemcpyHostToDevice, st));
  MR_CUDA(cudaMemcpyAsync(m->py.p, y, b, cudaMemcpyHostToDevice, st));
  MR_CUDA(cudaMemcpyAsync(m->pz.p, z, b, cudaMemcpyHostToDevice, st));
  ProcParams P;
  fill_proc(P, T16, lower, upper, min_r, beam_a, beam_c, sJ, rotVar, CSBT, PCBM, BrBS, 1);
  k_elev_process<<<(n + 255) / 256, 256, 0, st>>>(m->px.p, m->py.p, m->pz.p, n, P, m->g, m->lowest.p, m->pidx.p, m->pvar.p, m->pxt.p,
                                                 m->pyt.p, m->pzt.p, nullptr, nullptr);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaMemcpyAsync(var, m->pvar.p, b, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(z_ts, m->pzt.p, b, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(map_index, m->pidx.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(x_ts, m->pxt.p, b, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaMemcpyAsync(y_ts, m->pyt.p, b, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}

int mrslam_elev_fuse(mrslam_elev_t* h, int length, int n, const int* idx, const int* R, const int* G, const int* B,
                     const float* intensity, const float* height, const float* var, long long* n_fused) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (length != m->g.L || n < 0) {
    set_error("Fuse: length %d does not match the initialised map (%d)", length, m->g.L);
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  cudaStream_t st = rt().stream;
  MR_CHECK(m->reserve_points(n));
  if (n > 0) {
    size_t bi = sizeof(int) * (size_t)n, bf = sizeof(float) * (size_t)n;
    MR_CUDA(cudaMemcpyAsync(m->pidx.p, idx, bi, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->pR.p, R, bi, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->pG.p, G, bi, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->pB.p, B, bi, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->pint.p, intensity, bf, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->ph.p, height, bf, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(m->pvar.p, var, bf, cudaMemcpyHostToDevice, st));
    k_elev_keys<<<(n + 255) / 256, 256, 0, st>>>(m->pidx.p, m->ph.p, n, m->cells(), m->keys.p, m->vals.p);
    count_launch(1);
  }
  int s = fuse_sorted(m, n, m->ph.p, m->pvar.p, m->pR.p, m->pG.p, m->pB.p, m->pint.p, n_fused);
  if (s == MRSLAM_OK && !n_fused) MR_CUDA(cudaStreamSynchronize(st));  // Fuse is synchronous for the caller
  return s;
}

int mrslam_elev_process_fuse_device(mrslam_elev_t* h, const float* d_x, const float* d_y, const float* d_z, int n, const float* T16,
                                    double lower, double upper, float min_r, float beam_a, float beam_c, const float* sJ,
                                    const float* rotVar, const float* CSBT, const float* PCBM, const float* BrBS, int near_filter,
                                    long long* n_fused) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (n < 0) {
    set_error("process_fuse_device: n < 0");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  cudaStream_t st = rt().stream;
  MR_CHECK(m->reserve_points(n));
  if (n > 0) {
    ProcParams P;
    fill_proc(P, T16, lower, upper, min_r, beam_a, beam_c, sJ, rotVar, CSBT, PCBM, BrBS, near_filter);
    {
      ProfScope ps(PROF_ELEV_PROCESS);
      k_elev_process<<<(n + 255) / 256, 256, 0, st>>>(d_x, d_y, d_z, n, P, m->g, m->lowest.p, nullptr, m->pvar.p, nullptr, nullptr,
                                                     m->pzt.p, m->keys.p, m->vals.p);
    }
    count_launch(1);
  }
  return fuse_sorted(m, n, m->pzt.p, m->pvar.p, nullptr, nullptr, nullptr, nullptr, n_fused);
}

// The whole replay in one pass per group of <= 64 frames: host-side Move plans (geometry per frame, strips to clear, one bit per
// frame in per-row / per-column clear masks) -> one clear launch, one process launch over all frames, ONE stable sort by cell of
// all points (the concatenation is in (frame, point) order, so is every run afterwards), one gather, one run walk that applies the
// frame-end clamps and skips what a later Move wiped, one clamp.  ~10 launches for a group instead of ~8 per frame, and the sort /
// walk work on millions of keys instead of 120 k at a time.  Bit-identical to the per-frame sequence
// (tests/test_elevation_gpu.py::test_pipelined_batch_equals_per_frame_sequence).
static int elev_group_onepass(Elev* m, const mrslam_elev_frame_t* frames, int nframes, float resolution, int length, cudaStream_t st) {
  const int L = m->g.L, cells = m->cells();
  size_t total = 0;
  int max_n = 0;
  for (int f = 0; f < nframes; f++) {
    total += (size_t)frames[f].n;
    max_n = std::max(max_n, frames[f].n);
  }
  if (total >= (1ull << 31) || max_n >= (1 << 24)) {
    set_error("process_fuse_device_batch: too many points in one group");
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(m->batch_fence.wait());
  bool grew = false;
  MR_CHECK(m->reserve_batch(total, nframes, &grew));
  if (grew && st != rt().stream) {  // fresh allocations exist in library-stream order only
    MR_CUDA(cudaEventRecord(m->ev_start, rt().stream));
    MR_CUDA(cudaStreamWaitEvent(st, m->ev_start, 0));
  }
  unsigned long long* rowclr = m->h_clr.p;
  unsigned long long* colclr = m->h_clr.p + L;
  memset(m->h_clr.p, 0, sizeof(unsigned long long) * 2 * (size_t)L);
  int nclear = 0, max_cnt = 0;
  bool any_move = false;
  BatchOffsets bo;
  size_t off = 0;
  for (int f = 0; f < nframes; f++) {
    const mrslam_elev_frame_t& F = frames[f];
    if (F.do_move) {
      MovePlan mp;
      plan_move(m, F.position, resolution, length, mp);
      if (mp.clear_all) {  // cannot happen here: the caller splits groups at clear-all moves
        set_error("internal: clear-all move inside a group");
        return MRSLAM_ERR_STATE;
      }
      for (int k = 0; k < mp.nstrips; k++) {
        m->h_ctab.p[nclear++] = ClearTab{mp.start[k], mp.shift[k], mp.row_flag[k]};
        max_cnt = std::max(max_cnt, mp.shift[k] * L);
        unsigned long long* msk = mp.row_flag[k] ? rowclr : colclr;
        for (int r = 0; r < mp.shift[k]; r++) msk[mp.start[k] + r] |= 1ull << f;
        any_move = true;
      }
    }
    BatchFrame& B = m->h_frames.p[f];
    B.x = F.d_x; B.y = F.d_y; B.z = F.d_z;
    B.n = F.n;
    B.off = (int)off;
    B.g = m->g;
    fill_proc(B.P, F.T16, F.lower, F.upper, F.min_r, F.beam_a, F.beam_c, F.sJ, F.rotVar, F.CSBT, F.PCBM, F.BrBS, F.near_filter);
    bo.off[f] = (int)off;
    off += (size_t)F.n;
  }
  for (int f = nframes; f < ELEV_GROUP; f++) bo.off[f] = 0;
  MR_CUDA(upload_params(m->d_frames.p, m->h_frames.p, sizeof(BatchFrame) * nframes, st));
  if (nclear) MR_CUDA(upload_params(m->d_ctab.p, m->h_ctab.p, sizeof(ClearTab) * nclear, st));
  if (any_move) MR_CUDA(upload_params(m->d_clr.p, m->h_clr.p, sizeof(unsigned long long) * 2 * (size_t)L, st));
  MR_CHECK(m->batch_fence.mark(st));
  if (nclear) {
    k_elev_clear_tab<<<dim3((max_cnt + 255) / 256, nclear), 256, 0, st>>>(m->elevation.p, m->variance.p, m->intensity.p, m->R.p, m->G.p,
                                                                           m->B.p, L, m->d_ctab.p);
    count_launch(1);
  }
  const int n = (int)total;
  if (n > 0) {
    {
      ProfScope ps(PROF_ELEV_PROCESS);
      // grid-stride over the points of a frame: enough blocks to fill the GPU whatever the number of frames
      const int gx = std::max(1, std::min((max_n + 255) / 256, std::max(64, 4096 / nframes)));
      k_elev_process_batch<<<dim3(gx, nframes), 256, 0, st>>>(m->d_frames.p, cells, m->lowest.p, m->pzt.p, m->pvar.p, m->keys.p, m->vals.p);
      count_launch(1);
    }
    ProfScope ps(PROF_ELEV_FUSE);
    int bits = 1;
    while ((1ll << bits) <= cells) bits++;
    // own stable radix sort (csrc/radix_sort.cuh): the first pass drops the invalid points (key == cells: filtered, outside the
    // map, or in another GPU's rows of a tile-sharded grid), the count of the rest stays on the device
    unsigned *ks = nullptr, *vs = nullptr;
    const int* d_nv = nullptr;
    MR_CHECK(radix_sort_pairs(m->rsort, (unsigned*)m->keys.p, (unsigned*)m->vals.p, (unsigned*)m->keys_alt.p, (unsigned*)m->vals_alt.p, n, nullptr, 0,
                              ((bits + 7) / 8) * 8, (unsigned)cells, 1, st, &ks, &vs, &d_nv));
    k_elev_gather_batch<<<(n + 255) / 256, 256, 0, st>>>((const int*)vs, d_nv, bo, m->pzt.p, m->pvar.p, m->ghv.p, m->gfr.p);
    k_elev_fuse_runs_batch<<<(n + 127) / 128, 128, 0, st>>>((const int*)ks, d_nv, cells, L, m->ghv.p, m->gfr.p, any_move ? m->d_clr.p : nullptr,
                                                           any_move ? m->d_clr.p + L : nullptr, m->elevation.p, m->variance.p,
                                                           m->counter.p + 1);
    count_launch(2);
  }
  {
    ProfScope ps(PROF_ELEV_FUSE);
    k_elev_clamp_var<<<((cells + 3) / 4 + 255) / 256, 256, 0, st>>>(m->variance.p, cells);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

static int elev_batch(mrslam_elev_t* h, const mrslam_elev_frame_t* frames, int nframes, float resolution, int length, bool async) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (nframes < 0 || (nframes > 0 && !frames)) {
    set_error("process_fuse_device_batch: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  if (nframes == 0) return MRSLAM_OK;
  for (int f = 0; f < nframes; f++)
    if (frames[f].n < 0 || (frames[f].n > 0 && (!frames[f].d_x || !frames[f].d_y || !frames[f].d_z))) {
      set_error("process_fuse_device_batch: bad frame %d", f);
      return MRSLAM_ERR_ARG;
    }
  MR_CHECK(m->ensure_pipeline(1));
  MR_CHECK(m->counter.reserve(2));
  cudaStream_t st = async ? m->bstream : rt().stream;
  if (async) {
    MR_CUDA(cudaEventRecord(m->ev_start, rt().stream));  // inputs and the map state are ordered on the library stream
    MR_CUDA(cudaStreamWaitEvent(st, m->ev_start, 0));
  }
  // groups of <= ELEV_GROUP frames; a Move that jumps by a whole map length (clear-all) starts a new group after its fill
  int f0 = 0;
  while (f0 < nframes) {
    int f1 = f0;
    // dry-run the Moves on the geometry (restored afterwards) to find the end of the group: a clear-all move ends it
    const ElevGeom g0 = m->g;
    const float sz0 = m->sensor_z;
    bool first_is_clear_all = false;
    while (f1 < nframes && f1 - f0 < ELEV_GROUP) {
      if (frames[f1].do_move) {
        MovePlan mp;
        plan_move(m, frames[f1].position, resolution, length, mp);
        if (mp.clear_all) {
          first_is_clear_all = f1 == f0;
          break;
        }
      }
      f1++;
    }
    m->g = g0;
    m->sensor_z = sz0;
    if (first_is_clear_all) {
      MovePlan real;
      plan_move(m, frames[f0].position, resolution, length, real);
      MR_CHECK(apply_move_clears(m, real, length, st));
      mrslam_elev_frame_t F = frames[f0];  // the frame itself is replayed without its (already applied) move
      F.do_move = 0;
      MR_CHECK(elev_group_onepass(m, &F, 1, resolution, length, st));
      f0++;
      continue;
    }
    MR_CHECK(elev_group_onepass(m, frames + f0, f1 - f0, resolution, length, st));
    f0 = f1;
  }
  if (async) {
    MR_CUDA(cudaEventRecord(m->ev_done, st));
    m->pending_join = true;
  }
  return MRSLAM_OK;
}

int mrslam_elev_process_fuse_device_batch(mrslam_elev_t* h, const mrslam_elev_frame_t* frames, int nframes, float resolution, int length) {
  return elev_batch(h, frames, nframes, resolution, length, false);
}
int mrslam_elev_process_fuse_device_batch_async(mrslam_elev_t* h, const mrslam_elev_frame_t* frames, int nframes, float resolution,
                                                int length) {
  return elev_batch(h, frames, nframes, resolution, length, true);
}

// cell fuses accumulated by calls that passed n_fused == NULL (no host sync per frame); optionally reset
int mrslam_elev_fused_total(mrslam_elev_t* h, int reset, long long* total) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  unsigned long long c = 0;
  if (m->counter.p) {
    MR_CUDA(cudaMemcpyAsync(&c, m->counter.p + 1, sizeof(c), cudaMemcpyDeviceToHost, rt().stream));
    if (reset) MR_CUDA(cudaMemsetAsync(m->counter.p + 1, 0, sizeof(c), rt().stream));
    MR_CUDA(cudaStreamSynchronize(rt().stream));
  }
  if (total) *total = (long long)c;
  return MRSLAM_OK;
}

int mrslam_elev_mapvar_update(mrslam_elev_t* h, int length, float upd) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  int n = length * length;
  k_elev_mapvar<<<(n + 255) / 256, 256, 0, rt().stream>>>(m->variance.p, std::min(n, m->cells()), upd);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

int mrslam_elev_map_optmove(mrslam_elev_t* h, const float* opt_p, float height_update, float resolution, int length, float* out2) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  float last[2] = {m->g.cx, m->g.cy};
  for (int i = 0; i < 2; i++) {
    float ps = opt_p[i] - last[i];
    int is = static_cast<int>(ps / resolution + 0.5 * (ps > 0 ? 1 : -1));
    out2[i] = last[i] + resolution * is;
  }
  m->g.cx = out2[0];
  m->g.cy = out2[1];
  int n = std::min(length * length, m->cells());
  k_elev_height<<<(n + 255) / 256, 256, 0, rt().stream>>>(m->elevation.p, n, height_update);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

int mrslam_elev_map_closeloop(mrslam_elev_t* h, const float* upos, float height_update, int length, float resolution) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  int ishift[2];
  float pshift[2] = {upos[0] - m->g.cx, upos[1] - m->g.cy}, ashift[2];
  index_shift(ishift, pshift, resolution);
  for (int i = 0; i < 2; i++) ashift[i] = (float)ishift[i] * resolution;
  m->g.cx = position_to_range(m->g.cx, ashift[0], resolution);
  m->g.cy = position_to_range(m->g.cy, ashift[1], resolution);
  int n = std::min(length * length, m->cells());
  k_elev_height<<<(n + 255) / 256, 256, 0, rt().stream>>>(m->elevation.p, n, height_update);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

int mrslam_elev_get_map(mrslam_elev_t* h, float* elevation, float* variance, float* intensity, int* R, int* G, int* B, float* lowest) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  cudaStream_t st = rt().stream;
  size_t n = (size_t)m->cells();
  if (elevation) MR_CUDA(cudaMemcpyAsync(elevation, m->elevation.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (variance) MR_CUDA(cudaMemcpyAsync(variance, m->variance.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (intensity) MR_CUDA(cudaMemcpyAsync(intensity, m->intensity.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (R) MR_CUDA(cudaMemcpyAsync(R, m->R.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (G) MR_CUDA(cudaMemcpyAsync(G, m->G.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (B) MR_CUDA(cudaMemcpyAsync(B, m->B.p, n * 4, cudaMemcpyDeviceToHost, st));
  if (lowest) MR_CUDA(cudaMemcpyAsync(lowest, m->lowest.p, n * 4, cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}
// Map_feature (gpu_process.cu:1254-1298): traversability features of every cell + the read-out of the map.  Host arrays of
// length*length; any of them may be NULL (not copied).  Updates the resident traversability plane like the reference.
int mrslam_elev_map_feature(mrslam_elev_t* h, int length, float* elevation, float* var, int* colorR, int* colorG, int* colorB, float* rough,
                            float* slope, float* traver, float* intensity) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (length != m->g.L) {
    set_error("Map_feature: length %d does not match the map (%d)", length, m->g.L);
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  cudaStream_t st = rt().stream;
  const size_t n = (size_t)m->cells();
  MR_CHECK(m->feat.reserve(6 * n));
  MR_CHECK(m->feat_i.reserve(3 * n));
  float* f = m->feat.p;
  int* fi = m->feat_i.p;
  launch_mapfeature(m->elevation.p, m->variance.p, m->intensity.p, m->R.p, m->G.p, m->B.p, m->traver.p, m->g.L, m->g.res, m->g.sx, m->g.sy,
                    f, f + n, fi, fi + n, fi + 2 * n, f + 2 * n, f + 3 * n, f + 4 * n, f + 5 * n, st);
  MR_CUDA(cudaGetLastError());
  auto back = [&](void* dst, const void* src) -> int {
    if (dst) MR_CUDA(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToHost, st));
    return MRSLAM_OK;
  };
  MR_CHECK(back(elevation, f)); MR_CHECK(back(var, f + n)); MR_CHECK(back(rough, f + 2 * n)); MR_CHECK(back(slope, f + 3 * n));
  MR_CHECK(back(traver, f + 4 * n)); MR_CHECK(back(intensity, f + 5 * n));
  MR_CHECK(back(colorR, fi)); MR_CHECK(back(colorG, fi + n)); MR_CHECK(back(colorB, fi + 2 * n));
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}

// Raytracing (gpu_process.cu:1301-1316): obstacle cells (traversability below the obstacle threshold) that lie above the
// sensor-to-lowest-point rays are removed; map_lowest is reset to 10 afterwards.
int mrslam_elev_raytracing(mrslam_elev_t* h, int length) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (length != m->g.L) {
    set_error("Raytracing: length %d does not match the map (%d)", length, m->g.L);
    return MRSLAM_ERR_ARG;
  }
  if (m->g.row_lo != 0 || m->g.row_hi != m->g.L) {
    // a tile of a sharded grid only holds the map_lowest / traversability values of its own rows; rays cross tiles
    set_error("Raytracing: this map instance owns storage rows [%d, %d) of %d only (tile-sharded grid); gather the planes first",
              m->g.row_lo, m->g.row_hi, m->g.L);
    return MRSLAM_ERR_STATE;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  MR_CHECK(m->feat_i.reserve((size_t)m->cells() + 4));  // obstacle list + its counter (the Map_feature int staging is free here)
  launch_raytracing(m->elevation.p, m->variance.p, m->traver.p, m->lowest.p, m->g.L, m->g.sx, m->g.sy, m->sensor_z, m->obstacle_thr,
                    m->feat_i.p, m->feat_i.p + m->cells(), rt().stream);
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaStreamSynchronize(rt().stream));  // the reference synchronises here as well
  return MRSLAM_OK;
}

// Tile-sharded global grid (BASELINE config 5): this instance fuses only the points that fall into STORAGE rows
// [row_lo, row_hi) -- cells are disjoint across tiles, so the union of the tiles equals the single-GPU map bit for bit.
int mrslam_elev_set_owned_rows(mrslam_elev_t* h, int row_lo, int row_hi) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (row_lo < 0 || row_hi > m->g.L || row_lo > row_hi) {
    set_error("set_owned_rows: [%d, %d) outside [0, %d]", row_lo, row_hi, m->g.L);
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  m->g.row_lo = row_lo;
  m->g.row_hi = row_hi;
  return MRSLAM_OK;
}

// n_rows consecutive STORAGE rows of one plane <-> a DEVICE buffer of n_rows * length 4-byte elements (halo rows of a neighbouring
// tile before Map_feature, tile blocks for assembling the whole map).  plane: 0 elevation, 1 variance, 2 intensity, 3 lowest,
// 4 traversability, 5/6/7 colour R/G/B (int).  to_map != 0 writes the buffer into the map.  Stream-ordered on mrslam_stream().
int mrslam_elev_rows_copy(mrslam_elev_t* h, int plane, int storage_row, int n_rows, void* d_buf, int to_map) {
  REQ(h);
  Elev* m = E(h);
  REQ_INIT(m);
  if (plane < 0 || plane > 7 || storage_row < 0 || n_rows < 0 || storage_row + n_rows > m->g.L || (n_rows > 0 && !d_buf)) {
    set_error("rows_copy: bad arguments (plane %d rows %d+%d)", plane, storage_row, n_rows);
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk2(m->mu);
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(m->join());
  if (n_rows == 0) return MRSLAM_OK;
  void* planes[8] = {m->elevation.p, m->variance.p, m->intensity.p, m->lowest.p, m->traver.p, m->R.p, m->G.p, m->B.p};
  char* mp = (char*)planes[plane] + (size_t)storage_row * m->g.L * 4;
  const size_t bytes = (size_t)n_rows * m->g.L * 4;
  if (to_map) MR_CUDA(cudaMemcpyAsync(mp, d_buf, bytes, cudaMemcpyDeviceToDevice, rt().stream));
  else MR_CUDA(cudaMemcpyAsync(d_buf, mp, bytes, cudaMemcpyDeviceToDevice, rt().stream));
  return MRSLAM_OK;
}

int mrslam_elev_get_state(mrslam_elev_t* h, float* central2, int* start2) {
  REQ(h);
  Elev* m = E(h);
  central2[0] = m->g.cx;
  central2[1] = m->g.cy;
  start2[0] = m->g.sx;
  start2[1] = m->g.sy;
  return MRSLAM_OK;
}

// Stable radix sort of (key, value) pairs by key bits [0, bits) with the hot path's own kernels (csrc/radix_sort.cuh): host arrays
// in, host arrays out; keys equal to drop_key are removed when use_drop != 0.  A test / utility entry point -- the elevation
// replay calls the same code on device-resident keys.
int mrslam_sort_pairs_u32(const unsigned* keys, const unsigned* vals, int n, int bits, unsigned drop_key, int use_drop, unsigned* keys_out,
                          unsigned* vals_out, int* n_out) {
  MR_CHECK(ensure_init());
  if (n < 0 || bits < 1 || bits > 32 || (n > 0 && (!keys || !vals || !keys_out || !vals_out)) || !n_out) {
    set_error("sort_pairs_u32: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  cudaStream_t st = rt().stream;
  static RadixScratch S;
  static DevBuf<unsigned> ka, va, kb, vb;
  const size_t cap = std::max(n, 1);
  MR_CHECK(ka.reserve(cap)); MR_CHECK(va.reserve(cap)); MR_CHECK(kb.reserve(cap)); MR_CHECK(vb.reserve(cap));
  if (n > 0) {
    MR_CUDA(cudaMemcpyAsync(ka.p, keys, sizeof(unsigned) * (size_t)n, cudaMemcpyHostToDevice, st));
    MR_CUDA(cudaMemcpyAsync(va.p, vals, sizeof(unsigned) * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  unsigned *ks = nullptr, *vs = nullptr;
  const int* d_n = nullptr;
  MR_CHECK(radix_sort_pairs(S, ka.p, va.p, kb.p, vb.p, n, nullptr, 0, ((bits + 7) / 8) * 8, drop_key, use_drop, st, &ks, &vs, &d_n));
  int m = 0;
  MR_CUDA(cudaMemcpyAsync(&m, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  MR_CUDA(cudaStreamSynchronize(st));
  if (m > 0) {
    MR_CUDA(cudaMemcpyAsync(keys_out, ks, sizeof(unsigned) * (size_t)m, cudaMemcpyDeviceToHost, st));
    MR_CUDA(cudaMemcpyAsync(vals_out, vs, sizeof(unsigned) * (size_t)m, cudaMemcpyDeviceToHost, st));
    MR_CUDA(cudaStreamSynchronize(st));
  }
  *n_out = m;
  return MRSLAM_OK;
}
}
