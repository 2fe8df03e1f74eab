// elev_post.cu -- elevation post-passes on the resident map: traversability features and the
// ray-tracing obstacle clean-up that ElevationMapping::Callback runs after every Fuse (ElevationMapping.cpp:410, :425).
//
// Reference (Mapping/src/elevation_mapping_periodical/elevation_mapping/cuda/gpu_process.cu):
//   :66-185    computerEigenvalue   (float Jacobi rotations, largest off-diagonal pivot)  -> min_eigvec3
//   :547-668   G_Mapfeature         -> k_elev_mapfeature
//   :670-704   StorageP2geoP / Storageindex / P_isVaild / d_min_elevation               -> ray_bound
//   :706-889   G_Raytracing         -> k_elev_raytracing
//   :232-239   G_Clear_maplowest    -> folded into k_elev_raytracing's epilogue kernel
//
// Arithmetic: unlike the other translation units this one is compiled WITHOUT -fmad=false: the expressions below keep the
// operand order and float/double promotion of the reference's, and nvcc's own contraction then produces the same FMAs as in
// the reference build (checked on the B200 against the reference kernels compiled unmodified: tests/test_elevation_gpu.py).
//
// Reference quirks kept on purpose (a drop-in must evolve the map identically):
//   * the 5x5 neighbourhood is addressed in STORAGE order (wrapping), its validity is tested in geographic order;
//   * cells without elevation leave rough/slope/traver unwritten in the reference (uninitialised cudaMalloc memory); here
//     they are written as 0 / 0 / -10, the values of the "fewer than 8 neighbours" branch;
//   * the ray walk uses geographic indices as storage indices when it reads map_lowest, truncates the robot index to int,
//     returns early (no clean-up) for cells on the robot's row/column, and compares map_lowest with 10 although the map
//     is initialised with 100 (so every cell is "valid" until the first Raytracing reset it to 10).
#include "common.cuh"

namespace mrslam {

// eigenvector of the smallest eigenvalue of a symmetric 3x3 (row-major a[9], destroyed); gpu_process.cu:66-185
__device__ void min_eigvec3(float* a, float* vec, float eps, int max_rot) {
  float V[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i * 3 + j] = (i == j) ? 1.0f : 0.0f;
  int rotations = 0;
  while (true) {
    // largest off-diagonal magnitude; the scan starts from a[1] UNSIGNED, like the reference
    float big = a[1];
    int p = 0, q = 1;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        float d = fabs(a[i * 3 + j]);
        if ((i != j) && (d > big)) {
          big = d;
          p = i;
          q = j;
        }
      }
    if (big < eps) break;
    if (rotations > max_rot) break;
    rotations++;
    float app = a[p * 3 + p];
    float apq = a[p * 3 + q];
    float aqq = a[q * 3 + q];
    float ang = 0.5 * atan2(-2 * apq, aqq - app);
    float s = sin(ang);
    float c = cos(ang);
    float s2 = sin(2 * ang);
    float c2 = cos(2 * ang);
    a[p * 3 + p] = app * c * c + aqq * s * s + 2 * apq * c * s;
    a[q * 3 + q] = app * s * s + aqq * c * c - 2 * apq * c * s;
    a[p * 3 + q] = 0.5 * (aqq - app) * s2 + apq * c2;
    a[q * 3 + p] = a[p * 3 + q];
    for (int i = 0; i < 3; i++)
      if ((i != q) && (i != p)) {
        int u = i * 3 + p, w = i * 3 + q;
        float t = a[u];
        a[u] = a[w] * s + t * c;
        a[w] = a[w] * c - t * s;
      }
    for (int j = 0; j < 3; j++)
      if ((j != q) && (j != p)) {
        int u = p * 3 + j, w = q * 3 + j;
        float t = a[u];
        a[u] = a[w] * s + t * c;
        a[w] = a[w] * c - t * s;
      }
    for (int i = 0; i < 3; i++) {
      int u = i * 3 + p, w = i * 3 + q;
      float t = V[u];
      V[u] = V[w] * s + t * c;
      V[w] = V[w] * c - t * s;
    }
  }
  int best = 0;
  float least = a[0];
  for (int i = 1; i < 3; i++)
    if (least > a[i * 3 + i]) {
      least = a[i * 3 + i];
      best = i;
    }
  for (int i = 0; i < 3; i++) vec[i] = V[best + 3 * i];
}

// gpu_process.cu:547-668.  One thread per storage cell; also the read-out of the map (the reference's only one).
__global__ void k_elev_mapfeature(const float* __restrict__ elevation, const float* __restrict__ variance,
                                  const float* __restrict__ intensity, const int* __restrict__ R, const int* __restrict__ G,
                                  const int* __restrict__ B, float* __restrict__ traver_map, int L, float res, int sx, int sy,
                                  float* __restrict__ o_elev, float* __restrict__ o_var, int* __restrict__ o_R, int* __restrict__ o_G,
                                  int* __restrict__ o_B, float* __restrict__ o_rough, float* __restrict__ o_slope,
                                  float* __restrict__ o_traver, float* __restrict__ o_int) {
  int idx = threadIdx.x + blockDim.x * blockIdx.x;
  if (idx >= L * L) return;
  const float here = elevation[idx];
  o_elev[idx] = here;
  o_R[idx] = R[idx];
  o_G[idx] = G[idx];
  o_B[idx] = B[idx];
  o_int[idx] = intensity[idx];
  o_var[idx] = variance[idx];
  if (here == -10) {  // the reference returns here and leaves the three outputs unwritten
    o_slope[idx] = 0;
    o_rough[idx] = 0;
    o_traver[idx] = -10;
    return;
  }
  float X[25], Y[25], Z[25];
  float mx = 0, my = 0, mz = 0;
  const int cx = idx / L, cy = idx % L;
  int n = 0;
  for (int i = -2; i < 3; i++)
    for (int j = -2; j < 3; j++) {
      int gx = (cx + L - sx) % L;
      int gy = (cy + L - sy) % L;
      gx = gx + i;
      gy = gy + j;
      if (gx >= 0 && gx < L && gy >= 0 && gy < L) {
        int px = (cx + i + L) % L;
        int py = (cy + j + L) % L;
        float z = elevation[px * L + py];
        if (z != -10) {
          X[n] = px * res;
          Y[n] = py * res;
          Z[n] = z;
          mx = mx + X[n];
          my = my + Y[n];
          mz = mz + Z[n];
          n++;
        }
      }
    }
  if (n > 7) {
    mx = mx / n;
    my = my / n;
    mz = mz / n;
    float M[9] = {0};
    for (int i = 0; i < n; i++) {
      M[0] = M[0] + (X[i] - mx) * (X[i] - mx);
      M[4] = M[4] + (Y[i] - my) * (Y[i] - my);
      M[8] = M[8] + (Z[i] - mz) * (Z[i] - mz);
      M[1] = M[1] + (X[i] - mx) * (Y[i] - my);
      M[2] = M[2] + (X[i] - mx) * (Z[i] - mz);
      M[5] = M[5] + (Y[i] - my) * (Z[i] - mz);
      M[3] = M[1];
      M[6] = M[2];
      M[7] = M[5];
    }
    float nv[3];
    min_eigvec3(M, nv, 0.01f, 30);
    float slope;
    if (nv[2] > 0)
      slope = acos(nv[2]);
    else
      slope = acos(-nv[2]);
    float rough = fabs(here - mz);
    float tr = 0.5 * (1.0 - slope / 0.6) + 0.5 * (1.0 - (rough / 0.2));
    o_slope[idx] = slope;
    o_rough[idx] = rough;
    o_traver[idx] = tr;
    traver_map[idx] = tr;
  } else {
    o_slope[idx] = 0;
    o_rough[idx] = 0;
    o_traver[idx] = -10;
    traver_map[idx] = -10;
  }
}

// upper bound that the ray through (cell) leaves for the obstacle's height; gpu_process.cu:691-704 (geographic indices used
// as storage indices, as there)
__device__ __forceinline__ float ray_bound(const float* __restrict__ lowest, int L, float sensor_z, int cellx, int celly, int obstacle_x,
                                           float robot_x) {
  float x1 = (float)(cellx - obstacle_x);
  float x2 = (float)cellx - robot_x;
  int si = cellx * L + celly;
  float h2 = sensor_z - lowest[si];
  return lowest[si] + h2 / x2 * x1;
}
__device__ __forceinline__ bool ray_valid(const float* __restrict__ lowest, int L, int cellx, int celly) { return !(lowest[cellx * L + celly] == 10); }

// gpu_process.cu:706-889
__global__ void k_elev_raytracing(float* __restrict__ elevation, const float* __restrict__ variance, const float* __restrict__ traver,
                                  const float* __restrict__ lowest, int L, int sx, int sy, float sensor_z, float obstacle_thr) {
  int i = blockDim.x * blockIdx.x + threadIdx.x;
  if (i >= L * L) return;
  if (!(traver[i] < obstacle_thr && elevation[i] != -10)) return;
  const int cx = i / L, cy = i % L;
  int ob[2] = {(cx + L - sx) % L, (cy + L - sy) % L};
  const float ob_h = elevation[i];
  int cur[2] = {ob[0], ob[1]};
  int robot;
  float inc[2];
  if (L % 2 == 0) {
    robot = (float)(L / 2 - 0.5);
    inc[0] = ob[0] - robot;
    inc[1] = ob[1] - robot;
  } else {
    robot = (float)(L / 2);
    inc[0] = ob[0] - robot;
    inc[1] = ob[1] - robot;
  }
  const int stepx = inc[0] > 0 ? 1 : (inc[0] == 0 ? 0 : -1);
  const int stepy = inc[1] > 0 ? 1 : (inc[1] == 0 ? 0 : -1);
  // cells on the robot's row or column: the reference walks the ray and then returns without touching the map
  if (stepx == 0 || stepy == 0) return;
  float bound;
  float limit = ob_h;
  float dis = sqrt(inc[0] * inc[0] + inc[1] * inc[1]);
  float dir[2];
  dir[0] = inc[0] / dis;
  dir[1] = inc[1] / dis;
  float thr;
  if (fabs(inc[0]) > fabs(inc[1]))
    thr = sqrt(0.5 * 0.5 + pow(0.5 / inc[0] * inc[1], 2));
  else
    thr = sqrt(0.5 * 0.5 + pow(0.5 / inc[1] * inc[0], 2));
  float bx = (float)stepx / 2;
  float by = (float)stepy / 2;
  float tx = bx / dir[0];
  float ty = by / dir[1];
  float tprev = 0;
  while (cur[0] >= 0 && cur[0] < L && cur[1] >= 0 && cur[1] < L) {
    if (tx > ty) {
      if (ty - tprev > thr && cur[0] != ob[0] && cur[1] != ob[1]) {
        if (ray_valid(lowest, L, cur[0], cur[1])) {
          bound = ray_bound(lowest, L, sensor_z, cur[0], cur[1], ob[0], robot);
          if (bound < limit) limit = bound;
        }
      }
      cur[1] += stepy;
      by += (float)stepy;
      tprev = ty;
      ty = by / dir[1];
    } else if (tx < ty) {
      if (tx - tprev > thr && cur[0] != ob[0] && cur[1] != ob[1]) {
        if (ray_valid(lowest, L, cur[0], cur[1])) {
          bound = ray_bound(lowest, L, sensor_z, cur[0], cur[1], ob[0], robot);
          if (bound < limit) limit = bound;
        }
      }
      cur[0] += stepx;
      bx += (float)stepx;
      tprev = tx;
      tx = bx / dir[0];
    } else {
      if (tx - tprev > thr && cur[0] != ob[0] && cur[1] != ob[1]) {
        if (ray_valid(lowest, L, cur[0], cur[1])) {
          bound = ray_bound(lowest, L, sensor_z, cur[0], cur[1], ob[0], robot);
          if (bound < limit) limit = bound;
        }
      }
      cur[0] += stepx;
      cur[1] += stepy;
      bx += (float)stepx;
      by += (float)stepy;
      tprev = tx;
      tx = bx / dir[0];
      ty = by / dir[1];
    }
  }
  if (ob_h - 3 * sqrt(variance[i]) > limit) elevation[i] = -10;
}

// G_Clear_maplowest (gpu_process.cu:232-239)
__global__ void k_elev_reset_lowest(float* __restrict__ lowest, int n) {
  int i = blockDim.x * blockIdx.x + threadIdx.x;
  if (i < n) lowest[i] = 10;
}

void launch_mapfeature(const float* elevation, const float* variance, const float* intensity, const int* R, const int* G, const int* B,
                       float* traver_map, int L, float res, int sx, int sy, float* o_elev, float* o_var, int* o_R, int* o_G, int* o_B,
                       float* o_rough, float* o_slope, float* o_traver, float* o_int, cudaStream_t st) {
  const int n = L * L;
  k_elev_mapfeature<<<(n + 255) / 256, 256, 0, st>>>(elevation, variance, intensity, R, G, B, traver_map, L, res, sx, sy, o_elev, o_var, o_R,
                                                    o_G, o_B, o_rough, o_slope, o_traver, o_int);
  count_launch(1);
}

void launch_raytracing(float* elevation, const float* variance, const float* traver, float* lowest, int L, int sx, int sy, float sensor_z,
                       float obstacle_thr, cudaStream_t st) {
  const int n = L * L;
  k_elev_raytracing<<<(n + 255) / 256, 256, 0, st>>>(elevation, variance, traver, lowest, L, sx, sy, sensor_z, obstacle_thr);
  k_elev_reset_lowest<<<(n + 255) / 256, 256, 0, st>>>(lowest, n);
  count_launch(2);
}

}  // namespace mrslam
