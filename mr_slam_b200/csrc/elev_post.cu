// elev_post.cu -- elevation post-passes on the resident map: traversability features and the
// ray-tracing obstacle clean-up that ElevationMapping::Callback runs after every Fuse (ElevationMapping.cpp:410, :425).
//
// What is replaced (Mapping/src/elevation_mapping_periodical/elevation_mapping/cuda/gpu_process.cu):
//   G_Mapfeature :547-668 (+ computerEigenvalue :66-185)   one thread per cell gathers 25 neighbours from global memory into
//                                                          three 25-float local arrays, twice
//   G_Raytracing :706-889 (+ helpers :670-704)             one thread per CELL; the few obstacle cells each walk a ray of up
//                                                          to L steps sequentially while the other threads of their warp idle
//   G_Clear_maplowest :232-239
// B200 design:
//   k_elev_mapfeature_tiled   a block owns an 8 x 32 tile of storage cells; the tile plus its 2-cell halo (wrapping, as the ring
//                             buffer does) sits in shared memory once, both passes over the 5 x 5 window read it from there,
//                             no per-thread arrays; the symmetric moment matrix is six registers and the Jacobi iteration
//                             works on the three off-diagonals directly; all nine read-out planes are written coalesced.
//   k_elev_ray_collect        obstacle cells (traversability below the threshold, off the robot's row / column) are compacted
//                             into a list: only they do work afterwards.
//   k_elev_ray_events         ONE WARP PER RAY.  The reference's DDA is a merge of the two increasing sequences of x- and
//                             y-crossing times, t_x(k) = (k + 1/2) step_x / dir_x and t_y(m) likewise -- both exact in closed
//                             form because the numerators are half-integers -- so every crossing event can be evaluated
//                             independently: its lane finds how many crossings of the other axis precede it (estimate + exact
//                             float fix-up), which gives the cell it leaves, the previous event time, and the reference's
//                             acceptance test; the lowest bound is a warp min-reduction.  Same floats, same decisions.
//
// Parity: the Jacobi iteration stops at an absolute off-diagonal of 0.01 on UN-normalised moments (gpu_process.cu:80-100), so its
// result is defined by the iteration itself, not by the eigenvector it approximates; rotation arithmetic and pivot rule are
// therefore kept operation by operation (formula-level restatement), and this TU is compiled like the reference -- WITHOUT
// -fmad=false -- so nvcc contracts the same expressions.  Gate: tests/test_elevation_gpu.py runs the reference kernels side by
// side on the B200 (> 99 % of cells bit-identical, all within 2e-4, the same cells cleaned up).
//
// Reference quirks kept on purpose (a drop-in must evolve the map identically):
//   * the 5x5 neighbourhood is addressed in STORAGE order (wrapping), its validity is tested in geographic order, and the
//     x / y coordinates fed to the plane fit are the wrapped storage indices times the resolution;
//   * cells without elevation leave rough/slope/traver unwritten in the reference (uninitialised cudaMalloc memory); here
//     they are written as 0 / 0 / -10, the values of the "fewer than 8 neighbours" branch;
//   * the ray walk uses geographic indices as storage indices when it reads map_lowest, truncates the robot index to int,
//     returns early (no clean-up) for cells on the robot's row/column, and compares map_lowest with 10 although the map
//     is initialised with 100 (so every cell is "valid" until the first Raytracing reset it to 10).
#include "common.cuh"

namespace mrslam {

// ---------------------------------------------------------------------------------------------------------------------------
// traversability features
// ---------------------------------------------------------------------------------------------------------------------------
// Axis of the smallest diagonal entry after the reference's Jacobi iteration on the symmetric matrix
// [d0 o01 o02; o01 d1 o12; o02 o12 d2] (gpu_process.cu:66-185: pivot = largest |off-diagonal|, ties to the earlier of
// (0,1), (0,2), (1,2); stop below eps or after max_rot rotations).  The matrix lives in six registers; V is the accumulated
// rotation.  Every rotation evaluates the reference's expressions in the reference's order.
__device__ __forceinline__ void jacobi_min_axis(float d0, float d1, float d2, float o01, float o02, float o12, float eps, int max_rot,
                                                float vec[3]) {
  float V[9] = {1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f};
  int rotations = 0;
  while (true) {
    // the reference starts its scan from the SIGNED (0,1) entry and replaces it on a strictly larger magnitude
    float big = o01;
    int pair = 0;
    if (fabs(o01) > big) big = fabs(o01);
    if (fabs(o02) > big) { big = fabs(o02); pair = 1; }
    if (fabs(o12) > big) { big = fabs(o12); pair = 2; }
    if (big < eps) break;
    if (rotations > max_rot) break;
    rotations++;
    // (p, q) and the third axis r; app / aqq / apq and the two entries that couple r with p and q
    const int p = pair == 2 ? 1 : 0, q = pair == 0 ? 1 : 2;
    const float app = p == 0 ? d0 : d1, aqq = q == 1 ? d1 : d2, apq = pair == 0 ? o01 : (pair == 1 ? o02 : o12);
    float arp = pair == 0 ? o02 : (pair == 1 ? o01 : o01);  // a[r][p]
    float arq = pair == 0 ? o12 : (pair == 1 ? o12 : o02);  // a[r][q]
    float ang = 0.5 * atan2(-2 * apq, aqq - app);
    float s = sin(ang);
    float c = cos(ang);
    float s2 = sin(2 * ang);
    float c2 = cos(2 * ang);
    const float npp = app * c * c + aqq * s * s + 2 * apq * c * s;
    const float nqq = app * s * s + aqq * c * c - 2 * apq * c * s;
    const float npq = 0.5 * (aqq - app) * s2 + apq * c2;
    const float nrp = arq * s + arp * c;
    const float nrq = arq * c - arp * s;
    if (pair == 0) { d0 = npp; d1 = nqq; o01 = npq; o02 = nrp; o12 = nrq; }
    else if (pair == 1) { d0 = npp; d2 = nqq; o02 = npq; o01 = nrp; o12 = nrq; }
    else { d1 = npp; d2 = nqq; o12 = npq; o01 = nrp; o02 = nrq; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float t = V[i * 3 + p];
      V[i * 3 + p] = V[i * 3 + q] * s + t * c;
      V[i * 3 + q] = V[i * 3 + q] * c - t * s;
    }
  }
  int best = 0;
  float least = d0;
  if (least > d1) { least = d1; best = 1; }
  if (least > d2) { least = d2; best = 2; }
  for (int i = 0; i < 3; i++) vec[i] = V[best + 3 * i];
}

constexpr int MF_TR = 8, MF_TC = 32, MF_H = 2;  // tile rows x cols, halo

__global__ void __launch_bounds__(MF_TR* MF_TC) k_elev_mapfeature_tiled(
    const float* __restrict__ elevation, const float* __restrict__ variance, const float* __restrict__ intensity, const int* __restrict__ R,
    const int* __restrict__ G, const int* __restrict__ B, float* __restrict__ traver_map, int L, float res, int sx, int sy,
    float* __restrict__ o_elev, float* __restrict__ o_var, int* __restrict__ o_R, int* __restrict__ o_G, int* __restrict__ o_B,
    float* __restrict__ o_rough, float* __restrict__ o_slope, float* __restrict__ o_traver, float* __restrict__ o_int) {
  __shared__ float s_e[MF_TR + 2 * MF_H][MF_TC + 2 * MF_H + 1];
  const int r0 = blockIdx.y * MF_TR, c0 = blockIdx.x * MF_TC;
  const int tc = threadIdx.x & (MF_TC - 1), tr = threadIdx.x / MF_TC;
  // tile + halo of the elevation plane, storage order with wrap-around (the ring buffer's neighbourhood, gpu_process.cu:583-584)
  for (int e = threadIdx.x; e < (MF_TR + 2 * MF_H) * (MF_TC + 2 * MF_H); e += MF_TR * MF_TC) {
    const int rr = e / (MF_TC + 2 * MF_H), cc = e % (MF_TC + 2 * MF_H);
    const int pr = (r0 + rr - MF_H + 2 * L) % L, pc = (c0 + cc - MF_H + 2 * L) % L;
    s_e[rr][cc] = elevation[(size_t)pr * L + pc];
  }
  __syncthreads();
  const int cx = r0 + tr, cy = c0 + tc;
  if (cx >= L || cy >= L) return;
  const int idx = cx * L + cy;
  const float here = s_e[tr + MF_H][tc + MF_H];
  o_elev[idx] = here;
  o_R[idx] = R[idx];
  o_G[idx] = G[idx];
  o_B[idx] = B[idx];
  o_int[idx] = intensity[idx];
  o_var[idx] = variance[idx];
  float slope = 0, rough = 0, tr_out = -10;
  bool touch_map = false;
  if (here != -10) {
    touch_map = true;
    const int gx0 = (cx + L - sx) % L, gy0 = (cy + L - sy) % L;  // geographic position: validity of a neighbour is tested there
    // pass 1: count and means, in the reference's (i, j) order
    float mx = 0, my = 0, mz = 0;
    int n = 0;
    unsigned used = 0;  // bit (i+2)*5 + (j+2): neighbour contributes
#pragma unroll
    for (int i = -2; i < 3; i++)
#pragma unroll
      for (int j = -2; j < 3; j++) {
        const int gx = gx0 + i, gy = gy0 + j;
        if (gx >= 0 && gx < L && gy >= 0 && gy < L) {
          const float z = s_e[tr + MF_H + i][tc + MF_H + j];
          if (z != -10) {
            const int px = (cx + i + L) % L, py = (cy + j + L) % L;
            mx = mx + px * res;
            my = my + py * res;
            mz = mz + z;
            n++;
            used |= 1u << ((i + 2) * 5 + (j + 2));
          }
        }
      }
    if (n > 7) {
      mx = mx / n;
      my = my / n;
      mz = mz / n;
      // pass 2: centred second moments over the same neighbours, same order
      float m00 = 0, m11 = 0, m22 = 0, m01 = 0, m02 = 0, m12 = 0;
#pragma unroll
      for (int i = -2; i < 3; i++)
#pragma unroll
        for (int j = -2; j < 3; j++)
          if (used & (1u << ((i + 2) * 5 + (j + 2)))) {
            const int px = (cx + i + L) % L, py = (cy + j + L) % L;
            // the reference reads X, Y back from its local arrays, i.e. as ROUNDED products: keep them out of any contraction
            const float X = __fmul_rn((float)px, res), Y = __fmul_rn((float)py, res), Z = s_e[tr + MF_H + i][tc + MF_H + j];
            m00 = m00 + (X - mx) * (X - mx);
            m11 = m11 + (Y - my) * (Y - my);
            m22 = m22 + (Z - mz) * (Z - mz);
            m01 = m01 + (X - mx) * (Y - my);
            m02 = m02 + (X - mx) * (Z - mz);
            m12 = m12 + (Y - my) * (Z - mz);
          }
      float nv[3];
      jacobi_min_axis(m00, m11, m22, m01, m02, m12, 0.01f, 30, nv);
      if (nv[2] > 0)
        slope = acos(nv[2]);
      else
        slope = acos(-nv[2]);
      rough = fabs(here - mz);
      tr_out = 0.5 * (1.0 - slope / 0.6) + 0.5 * (1.0 - (rough / 0.2));
    }
  }
  o_slope[idx] = slope;
  o_rough[idx] = rough;
  o_traver[idx] = tr_out;
  if (touch_map) traver_map[idx] = tr_out;
}

// ---------------------------------------------------------------------------------------------------------------------------
// ray-tracing clean-up
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ray_robot_index(int L) { return (L % 2 == 0) ? (int)(float)(L / 2 - 0.5) : (int)(float)(L / 2); }

// obstacle cells that will walk a ray: traversability below the threshold, a height, and off the robot's row and column (for
// those the reference walks the ray and then returns without touching the map, gpu_process.cu:760-775)
__global__ void k_elev_ray_collect(const float* __restrict__ elevation, const float* __restrict__ traver, int L, int sx, int sy,
                                   float obstacle_thr, int* __restrict__ list, int* __restrict__ count) {
  const int i = blockDim.x * blockIdx.x + threadIdx.x;
  bool take = false;
  if (i < L * L && traver[i] < obstacle_thr && elevation[i] != -10) {
    const int robot = ray_robot_index(L);
    const int ox = (i / L + L - sx) % L, oy = (i % L + L - sy) % L;
    take = ox != robot && oy != robot;
  }
  const unsigned m = __ballot_sync(FULL, take);
  if (m) {
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (take) list[base + __popc(m & ((1u << lane) - 1u))] = i;
  }
}

// bound that the ray through (cellx, celly) leaves for the obstacle's height; gpu_process.cu:691-704 (geographic indices used
// as storage indices, as there)
__device__ __forceinline__ float ray_bound(const float* __restrict__ lowest, int L, float sensor_z, int cellx, int celly, int obstacle_x,
                                           float robot_x) {
  float x1 = (float)(cellx - obstacle_x);
  float x2 = (float)cellx - robot_x;
  int si = cellx * L + celly;
  float h2 = sensor_z - lowest[si];
  return lowest[si] + h2 / x2 * x1;
}

// crossing time of event number e on one axis: the reference's b += step; t = b / dir with b = step / 2 + e * step (exact)
__device__ __forceinline__ float ray_time(int e, int step, float dir) { return ((float)step / 2 + (float)(e * step)) / dir; }
// number of crossings e' >= 0 of an axis with time strictly below t (times increase with e')
__device__ __forceinline__ int ray_count_before(float t, int step, float dir, int emax) {
  int e = (int)floorf(t * fabsf(dir) - 0.5f) + 1;  // real-arithmetic estimate of #{e' : (e' + 1/2) / |dir| < t}
  e = max(0, min(e, emax + 1));
  while (e > 0 && !(ray_time(e - 1, step, dir) < t)) e--;
  while (e <= emax && ray_time(e, step, dir) < t) e++;
  return e;
}

// One warp per collected obstacle cell (warps stride over the list).
__global__ void __launch_bounds__(128) k_elev_ray_events(float* __restrict__ elevation, const float* __restrict__ variance,
                                                         const float* __restrict__ lowest, int L, int sx, int sy, float sensor_z,
                                                         const int* __restrict__ list, const int* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int n = *count;
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nwarps) {
    const int i = list[w];
    const int ob0 = (i / L + L - sx) % L, ob1 = (i % L + L - sy) % L;
    const float ob_h = elevation[i];
    const int robot = ray_robot_index(L);
    float inc[2];
    inc[0] = ob0 - robot;
    inc[1] = ob1 - robot;
    const int stepx = inc[0] > 0 ? 1 : -1, stepy = inc[1] > 0 ? 1 : -1;  // neither is 0 (k_elev_ray_collect)
    float dis = sqrt(inc[0] * inc[0] + inc[1] * inc[1]);
    float dir[2];
    dir[0] = inc[0] / dis;
    dir[1] = inc[1] / dis;
    float thr;
    if (fabs(inc[0]) > fabs(inc[1]))
      thr = sqrt(0.5 * 0.5 + pow(0.5 / inc[0] * inc[1], 2));
    else
      thr = sqrt(0.5 * 0.5 + pow(0.5 / inc[1] * inc[0], 2));
    // last in-grid state along each axis: the walk is at (ob0 + k stepx, ob1 + m stepy) with 0 <= k <= kmax, 0 <= m <= mmax
    const int kmax = stepx > 0 ? L - 1 - ob0 : ob0, mmax = stepy > 0 ? L - 1 - ob1 : ob1;
    float limit = ob_h;
    // x-crossing events (they also own the ties t_x == t_y): event k happens at state (k, m) with m = #{t_y < t_x(k)}
    for (int k = lane; k <= kmax; k += 32) {
      const float t = ray_time(k, stepx, dir[0]);
      const int m = ray_count_before(t, stepy, dir[1], mmax);
      if (m > mmax) continue;  // the walk has left the grid before this event
      float tprev = 0;
      if (k > 0) tprev = ray_time(k - 1, stepx, dir[0]);
      if (m > 0) tprev = fmaxf(tprev, ray_time(m - 1, stepy, dir[1]));
      if (t - tprev > thr && k != 0 && m != 0) {
        const int c0 = ob0 + k * stepx, c1 = ob1 + m * stepy;
        if (!(lowest[c0 * L + c1] == 10)) {
          const float bound = ray_bound(lowest, L, sensor_z, c0, c1, ob0, robot);
          if (bound < limit) limit = bound;
        }
      }
    }
    // y-crossing events that are not ties: event m happens at state (k, m) with k = #{t_x < t_y(m)}, unless t_x(k) == t_y(m)
    for (int m = lane; m <= mmax; m += 32) {
      const float t = ray_time(m, stepy, dir[1]);
      const int k = ray_count_before(t, stepx, dir[0], kmax);
      if (k > kmax) continue;
      if (ray_time(k, stepx, dir[0]) == t) continue;  // tie: evaluated as the x event above
      float tprev = 0;
      if (m > 0) tprev = ray_time(m - 1, stepy, dir[1]);
      if (k > 0) tprev = fmaxf(tprev, ray_time(k - 1, stepx, dir[0]));
      if (t - tprev > thr && k != 0 && m != 0) {
        const int c0 = ob0 + k * stepx, c1 = ob1 + m * stepy;
        if (!(lowest[c0 * L + c1] == 10)) {
          const float bound = ray_bound(lowest, L, sensor_z, c0, c1, ob0, robot);
          if (bound < limit) limit = bound;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(FULL, limit, o);
      if (other < limit) limit = other;
    }
    if (lane == 0 && ob_h - 3 * sqrt(variance[i]) > limit) elevation[i] = -10;
  }
}

// G_Clear_maplowest (gpu_process.cu:232-239)
__global__ void k_elev_reset_lowest(float* __restrict__ lowest, int n) {
  int i = blockDim.x * blockIdx.x + threadIdx.x;
  if (i < n) lowest[i] = 10;
}

void launch_mapfeature(const float* elevation, const float* variance, const float* intensity, const int* R, const int* G, const int* B,
                       float* traver_map, int L, float res, int sx, int sy, float* o_elev, float* o_var, int* o_R, int* o_G, int* o_B,
                       float* o_rough, float* o_slope, float* o_traver, float* o_int, cudaStream_t st) {
  dim3 grid((L + MF_TC - 1) / MF_TC, (L + MF_TR - 1) / MF_TR);
  k_elev_mapfeature_tiled<<<grid, MF_TR * MF_TC, 0, st>>>(elevation, variance, intensity, R, G, B, traver_map, L, res, sx, sy, o_elev, o_var, o_R,
                                                           o_G, o_B, o_rough, o_slope, o_traver, o_int);
  count_launch(1);
}

// scratch: `list` holds up to L*L ints, `count` one int (both device memory owned by the map object)
void launch_raytracing(float* elevation, const float* variance, const float* traver, float* lowest, int L, int sx, int sy, float sensor_z,
                       float obstacle_thr, int* list, int* count, cudaStream_t st) {
  const int n = L * L;
  cudaMemsetAsync(count, 0, sizeof(int), st);
  k_elev_ray_collect<<<(n + 255) / 256, 256, 0, st>>>(elevation, traver, L, sx, sy, obstacle_thr, list, count);
  k_elev_ray_events<<<148 * 4, 128, 0, st>>>(elevation, variance, lowest, L, sx, sy, sensor_z, list, count);
  k_elev_reset_lowest<<<(n + 255) / 256, 256, 0, st>>>(lowest, n);
  count_launch(3);
}

}  // namespace mrslam
