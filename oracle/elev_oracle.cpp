// oracle/elev_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Sequential CPU restatement of the reference elevation-mapping point process + fuse
// (paths relative to /root/reference/Mapping/src/elevation_mapping_periodical/elevation_mapping/):
//   cuda/gpu_process.cu:198-276    G_Init_map / G_Clear_allmap / G_Clear_map (sentinels)
//   cuda/gpu_process.cu:309-358    PointsToIndex / PointsToMapIndex
//   cuda/gpu_process.cu:384-455    G_pointsprocess
//   cuda/gpu_process.cu:477-535    G_fuse  (per cell, ascending point index)
//   cuda/gpu_process.cu:538-545    G_Mapvar_update
//   cuda/gpu_process.cu:890-918    getIndexShiftFromPositionShift / IndexToRange
//   cuda/gpu_process.cu:938-1081   Init_GPU_elevationmap / PositionToRange / Move
//   cuda/gpu_process.cu:1193-1252  G_update_mapheight / Map_optmove / Map_closeloop
//
// The reference is CUDA compiled by nvcc with its default -fmad=true, i.e. a*b+c chains
// contract to FMAs.  This restatement is compiled with -ffp-contract=off and writes the
// contraction nvcc performs explicitly (left-to-right: t=a*b; t=fma(c,d,t); ...), so that
// cell indices can be compared bit-exactly.  The reference .cu itself needs Eigen (absent)
// -- it is additionally compiled against a minimal Eigen shim into oracle/_ref/ for the GPU
// box (see oracle/build_ref.sh); reference tests/fixtures for this path: none exist
// ("parity unpinned" by the reference; pinned by running oracle/_ref on the GPU box).
//
// Deliberate restatement notes:
//   * G_fuse is O(cells x points) in the reference; per cell it visits points in ascending
//     index -- here: one ascending pass over points, which yields the same per-cell order.
//   * map_lowest update in G_pointsprocess is racy in the reference (gpu_process.cu:434-438);
//     the oracle keeps the plain minimum (+3*var when the point is the minimum, in index order).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct ElevMap {
  int L = 0;
  float res = 0;
  float mahal_thr = 0, obstacle_thr = 0;
  float central[2] = {0, 0};
  int start[2] = {0, 0};
  float sensor_z = 0;
  int row_lo = 0, row_hi = 0;  // owned STORAGE rows (tile-sharded grid); [0, L) after init
  std::vector<float> elevation, variance, intensity, lowest, traver;
  std::vector<int> R, G, B;

  void init(int length, float resolution, float mahal, float obst) {
    L = length;
    res = resolution;
    mahal_thr = mahal;
    obstacle_thr = obst;
    central[0] = central[1] = 0;
    start[0] = start[1] = 0;
    row_lo = 0;
    row_hi = L;
    size_t n = (size_t)L * L;
    elevation.assign(n, -10.f);
    variance.assign(n, -10.f);
    intensity.assign(n, 0.f);
    lowest.assign(n, 100.f);
    traver.assign(n, -10.f);
    R.assign(n, 0);
    G.assign(n, 0);
    B.assign(n, 0);
  }

  // gpu_process.cu:309-358 (both helpers share the index math)
  bool cell_of(float px, float py, int& ix, int& iy) const {
    float sx = px - central[0], sy = py - central[1];
    if (L % 2 == 0) {
      ix = (int)((float)(L / 2) - sx / res);
      iy = (int)((float)(L / 2) - sy / res);
    } else {
      ix = L / 2 - static_cast<int>((double)(sx / res) + 0.5 * (sx > 0 ? 1 : -1));
      iy = L / 2 - static_cast<int>((double)(sy / res) + 0.5 * (sy > 0 ? 1 : -1));
    }
    return ix >= 0 && ix < L && iy >= 0 && iy < L;
  }
  int points_to_index(float px, float py) const {
    int ix, iy;
    return cell_of(px, py, ix, iy) ? ix * L + iy : -1;
  }
  int points_to_map_index(float px, float py) const {
    int ix, iy;
    if (!cell_of(px, py, ix, iy)) return -1;
    return ((ix + start[0]) % L) * L + (iy + start[1]) % L;
  }

  void clear_cell(size_t i, bool all) {
    intensity[i] = 0;
    elevation[i] = -10;
    variance[i] = -10;
    if (all) traver[i] = -10;
    R[i] = G[i] = B[i] = 0;
  }
  void clear_rows(int startrow, int nrows) {  // G_Clear_map(start, shift, true)
    for (int i = 0; i < L * nrows; i++) clear_cell((size_t)startrow * L + i, false);
  }
  void clear_cols(int startcol, int ncols) {  // G_Clear_map(start, shift, false)
    for (int i = 0; i < L * ncols; i++) clear_cell((size_t)(i / ncols) * L + i % ncols + startcol, false);
  }
};

inline int index_to_range(int index, int L) {
  if (index < 0) index += ((-index / L) + 1) * L;
  return index % L;
}
inline float position_to_range(float p, float shift, float res) {
  int p_index = (int)std::round(p / res);
  int shift_index = (int)std::round(shift / res);
  return (p_index + shift_index) * res;
}
inline void index_shift(int* ishift, const float* pshift, float res) {
  for (int i = 0; i < 2; i++) ishift[i] = static_cast<int>((double)(pshift[i] / res) + 0.5 * (pshift[i] > 0 ? 1 : -1));
}

}  // namespace

extern "C" {

void* orc_elev_create() { return new ElevMap(); }
void orc_elev_destroy(void* h) { delete (ElevMap*)h; }
void orc_elev_init(void* h, int length, float resolution, float mahal, float obst) { ((ElevMap*)h)->init(length, resolution, mahal, obst); }

// Move (gpu_process.cu:1002-1081)
void orc_elev_move(void* h, const float* current_position3, float resolution, int length, float* central_out2, int* start_out2,
                   float* aligned_shift_out2) {
  ElevMap* m = (ElevMap*)h;
  int ishift[2];
  float pshift[2];
  m->sensor_z = current_position3[2];
  float hc[2] = {m->central[0], m->central[1]};
  int hs[2] = {m->start[0], m->start[1]};
  pshift[0] = current_position3[0] - hc[0];
  pshift[1] = current_position3[1] - hc[1];
  index_shift(ishift, pshift, resolution);
  for (int i = 0; i < 2; i++) aligned_shift_out2[i] = (float)ishift[i] * resolution;
  for (int i = 0; i < 2; i++) {
    if (ishift[i] != 0) {
      // The reference tests only `ishift >= length`; a negative shift of >= length cells then clears more strips than the map has
      // and writes out of bounds (undefined behaviour).  Oracle and CUDA path both treat |shift| >= length as clear-all.
      if (ishift[i] >= length || ishift[i] <= -length) {
        for (size_t c = 0; c < (size_t)length * length; c++) m->clear_cell(c, true);
      } else {
        int sign = ishift[i] > 0 ? 1 : -1;
        int startIndex = hs[i] - (sign > 0 ? 1 : 0);
        int endIndex = startIndex + sign - ishift[i];
        int nCells = std::abs(ishift[i]);
        int index = sign < 0 ? startIndex : endIndex;
        index = index_to_range(index, length);
        if (index + nCells <= length) {
          if (i == 0) m->clear_rows(index, nCells);
          else m->clear_cols(index, nCells);
        } else {
          int first = index, firstn = length - first;
          if (i == 0) m->clear_rows(first, firstn);
          else m->clear_cols(first, firstn);
          int secondn = nCells - firstn;
          if (i == 0) m->clear_rows(0, secondn);
          else m->clear_cols(0, secondn);
        }
      }
    }
    hs[i] -= ishift[i];
    hs[i] = index_to_range(hs[i], length);
    hc[i] = position_to_range(hc[i], aligned_shift_out2[i], resolution);
  }
  m->start[0] = hs[0];
  m->start[1] = hs[1];
  m->central[0] = hc[0];
  m->central[1] = hc[1];
  central_out2[0] = hc[0];
  central_out2[1] = hc[1];
  start_out2[0] = hs[0];
  start_out2[1] = hs[1];
}

// Process_points (gpu_process.cu:1083-1142 -> G_pointsprocess :384-455).
// T: row-major 4x4; sensorJacobian (1x3), rotationVariance (3x3 row-major), C_SB_transpose (3x3),
// P_mul_C_BM_transpose (1x3), B_r_BS_skew (3x3).  Host x/y/z are NOT modified (reference copies).
int orc_elev_process_points(void* h, int* map_index, const float* px, const float* py, const float* pz, float* var, float* x_ts,
                            float* y_ts, float* z_ts, const float* T, int n, double lower, double upper, float min_r, float beam_a,
                            float beam_c, const float* sJ, const float* rotVar, const float* C_SB_T, const float* P_mul_C_BM_T,
                            const float* B_r_BS_skew) {
  ElevMap* m = (ElevMap*)h;
  for (int i = 0; i < n; i++) {
    float x = px[i], y = py[i], z = pz[i];
    // transform(2,0)*x + transform(2,1)*y + transform(2,2)*z + transform(2,3)
    float height = fmaf(T[10], z, fmaf(T[9], y, T[8] * x)) + T[11];
    int flag = 0;
    if ((x > -1.5 && x < 1.5 && y > -1.5 && y < 1.5) || (y > -1 && y < 1) || y > 0) flag = 1;
    if (((double)height > lower && (double)height < upper) && flag == 0) {
      float xt = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3];
      float yt = fmaf(T[6], z, fmaf(T[5], y, T[4] * x)) + T[7];
      x_ts[i] = xt;
      y_ts[i] = yt;
      z_ts[i] = height;
      float dist = sqrtf(fmaf(z, z, fmaf(y, y, x * x)));
      float varN = min_r * min_r;                    // pow(C_min_r, 2)
      float lat = fmaf(beam_a, dist, beam_c);        // C_beam_c + C_beam_a * measurementDistance
      float varL = lat * lat;
      // v = C_SB_transpose * p ; skew(v) + B_r_BS_skew ; rotationJacobian = P_mul_C_BM_T * (...)
      float v[3];
      for (int r = 0; r < 3; r++) v[r] = fmaf(C_SB_T[3 * r + 2], z, fmaf(C_SB_T[3 * r + 1], y, C_SB_T[3 * r] * x));
      float S[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
      for (int a = 0; a < 9; a++) S[a] = S[a] + B_r_BS_skew[a];
      float rJ[3];
      for (int c = 0; c < 3; c++) rJ[c] = fmaf(P_mul_C_BM_T[2], S[6 + c], fmaf(P_mul_C_BM_T[1], S[3 + c], P_mul_C_BM_T[0] * S[c]));
      // cuda_computer(A,B,C) = (A*B) . C
      float a1[3];
      for (int c = 0; c < 3; c++) a1[c] = fmaf(rJ[2], rotVar[6 + c], fmaf(rJ[1], rotVar[3 + c], rJ[0] * rotVar[c]));
      float hv = fmaf(a1[2], rJ[2], fmaf(a1[1], rJ[1], a1[0] * rJ[0]));
      float sv[3] = {varL, varL, varN};
      float b1[3];
      for (int c = 0; c < 3; c++) b1[c] = sJ[c] * sv[c];  // diagonal sensorVariance: off-diagonal products are exact zeros
      hv += fmaf(b1[2], sJ[2], fmaf(b1[1], sJ[1], b1[0] * sJ[0]));
      var[i] = hv;
      int gi = m->points_to_index(xt, yt);
      int mi = m->points_to_map_index(xt, yt);
      if (mi != -1 && (mi / m->L < m->row_lo || mi / m->L >= m->row_hi)) {  // a cell of another tile (tile-sharded grid)
        mi = -1;
        gi = -1;
      }
      map_index[i] = mi;
      if (gi != -1) {
        if (height < m->lowest[gi]) m->lowest[gi] = height + 3 * hv;
      }
    } else {
      map_index[i] = -1;
      x_ts[i] = y_ts[i] = z_ts[i] = -1;
      var[i] = -1;
    }
  }
  return 0;
}

// Fuse (gpu_process.cu:1152-1191 -> G_fuse :477-535); returns the number of "cell fuses"
// (valid point landing in a cell) -- the unit of the cells-fused/s metric.
long orc_elev_fuse(void* h, int length, int n, const int* idx, const int* cR, const int* cG, const int* cB, const float* inten,
                   const float* ph, const float* pv) {
  ElevMap* m = (ElevMap*)h;
  long fused = 0;
  const int cells = length * length;
  for (int i = 0; i < n; i++) {
    int c = idx[i];
    if (c < 0 || c >= cells || ph[i] == -1) continue;
    fused++;
    bool colour = cR[i] != 0 && cG[i] != 0 && cB[i] != 0 && inten[i] != 0;
    float& e = m->elevation[c];
    float& v = m->variance[c];
    bool write_colour = false;
    if (e == -10) {
      e = ph[i];
      v = pv[i];
      write_colour = colour;
    } else {
      float md = fabsf(ph[i] - e) / sqrtf(v);
      if (md > 5) {
        if (e < ph[i]) {
          e = ph[i];
          v = pv[i];
          write_colour = colour;
        }
      } else {
        // (var*h + pv*e) / (var + pv): nvcc contracts the numerator to fma(var, h, pv*e)
        float ne = fmaf(v, ph[i], pv[i] * e) / (v + pv[i]);
        float nv = (pv[i] * v) / (pv[i] + v);
        e = ne;
        v = nv;
        write_colour = colour;
      }
    }
    if (write_colour) {
      m->intensity[c] = inten[i];
      m->R[c] = cR[i];
      m->G[c] = cG[i];
      m->B[c] = cB[i];
    }
  }
  for (int c = 0; c < cells; c++)
    if (m->variance[c] < 0.0001f) m->variance[c] = 0.0001f;  // includes empty (-10) cells, as the reference
  return fused;
}

void orc_elev_mapvar_update(void* h, int length, float var_update) {
  ElevMap* m = (ElevMap*)h;
  for (int i = 0; i < length * length; i++)
    if (m->variance[i] != -10) m->variance[i] += var_update;
}

void orc_elev_map_optmove(void* h, const float* opt_p, float height_update, float resolution, int length, float* opt_aligned2) {
  ElevMap* m = (ElevMap*)h;
  float last[2] = {m->central[0], m->central[1]};
  for (int i = 0; i < 2; i++) {
    float ps = opt_p[i] - last[i];
    int is = static_cast<int>((double)(ps / resolution) + 0.5 * (ps > 0 ? 1 : -1));
    opt_aligned2[i] = last[i] + resolution * (float)is;
    m->central[i] = opt_aligned2[i];
  }
  for (int i = 0; i < length * length; i++)
    if (m->elevation[i] != -10) m->elevation[i] += height_update;
}

void orc_elev_map_closeloop(void* h, const float* update_position, float height_update, int length, float resolution) {
  ElevMap* m = (ElevMap*)h;
  int ishift[2];
  float pshift[2], ashift[2];
  pshift[0] = update_position[0] - m->central[0];
  pshift[1] = update_position[1] - m->central[1];
  index_shift(ishift, pshift, resolution);
  for (int i = 0; i < 2; i++) ashift[i] = (float)ishift[i] * resolution;
  for (int i = 0; i < 2; i++) m->central[i] = position_to_range(m->central[i], ashift[i], resolution);
  for (int i = 0; i < length * length; i++)
    if (m->elevation[i] != -10) m->elevation[i] += height_update;
}

// ---- post-passes.  The CPU libm (sinf/cosf/atan2f/acosf) is not CUDA's, so these twins agree with the
// GPU to ~1e-6, not bit for bit; the bit-level pin for them is the reference kernel itself run on the GPU box (oracle/_ref).
// computerEigenvalue, gpu_process.cu:66-185: Jacobi rotations on the largest off-diagonal entry, eigenvector of the smallest
// eigenvalue.  The pivot scan starts from the SIGNED a[1] like the reference.
static void orc_min_eigvec3(float* a, float* vec, float eps, int max_rot) {
  float V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int rot = 0;
  for (;;) {
    float big = a[1];
    int p = 0, q = 1;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        float d = std::fabs(a[i * 3 + j]);
        if (i != j && d > big) { big = d; p = i; q = j; }
      }
    if (big < eps || rot > max_rot) break;
    rot++;
    float app = a[p * 3 + p], apq = a[p * 3 + q], aqq = a[q * 3 + q];
    float ang = (float)(0.5 * (double)atan2f(-2 * apq, aqq - app));
    float sn = sinf(ang), cs = cosf(ang), s2 = sinf(2 * ang), c2 = cosf(2 * ang);
    a[p * 3 + p] = app * cs * cs + aqq * sn * sn + 2 * apq * cs * sn;
    a[q * 3 + q] = app * sn * sn + aqq * cs * cs - 2 * apq * cs * sn;
    a[p * 3 + q] = (float)(0.5 * (double)(aqq - app) * (double)s2 + (double)(apq * c2));
    a[q * 3 + p] = a[p * 3 + q];
    for (int i = 0; i < 3; i++)
      if (i != q && i != p) {
        float t = a[i * 3 + p];
        a[i * 3 + p] = a[i * 3 + q] * sn + t * cs;
        a[i * 3 + q] = a[i * 3 + q] * cs - t * sn;
      }
    for (int j = 0; j < 3; j++)
      if (j != q && j != p) {
        float t = a[p * 3 + j];
        a[p * 3 + j] = a[q * 3 + j] * sn + t * cs;
        a[q * 3 + j] = a[q * 3 + j] * cs - t * sn;
      }
    for (int i = 0; i < 3; i++) {
      float t = V[i * 3 + p];
      V[i * 3 + p] = V[i * 3 + q] * sn + t * cs;
      V[i * 3 + q] = V[i * 3 + q] * cs - t * sn;
    }
  }
  int best = 0;
  float least = a[0];
  for (int i = 1; i < 3; i++)
    if (least > a[i * 3 + i]) { least = a[i * 3 + i]; best = i; }
  for (int i = 0; i < 3; i++) vec[i] = V[best + 3 * i];
}

// G_Mapfeature, gpu_process.cu:547-668 (+ host wrapper :1254-1298).  Cells without elevation: 0 / 0 / -10 (the reference leaves
// them unwritten).  Outputs may be NULL.
void orc_elev_map_feature(void* h, float* elevation, float* var, int* R, int* G, int* B, float* rough, float* slope, float* traver,
                          float* intensity) {
  ElevMap* m = (ElevMap*)h;
  const int L = m->L;
  for (int idx = 0; idx < L * L; idx++) {
    if (elevation) elevation[idx] = m->elevation[idx];
    if (var) var[idx] = m->variance[idx];
    if (R) R[idx] = m->R[idx];
    if (G) G[idx] = m->G[idx];
    if (B) B[idx] = m->B[idx];
    if (intensity) intensity[idx] = m->intensity[idx];
    float sl = 0, ro = 0, tr = -10;
    if (m->elevation[idx] != -10) {
      float X[25], Y[25], Z[25], mx = 0, my = 0, mz = 0;
      int n = 0;
      const int cx = idx / L, cy = idx % L;
      for (int i = -2; i < 3; i++)
        for (int j = -2; j < 3; j++) {
          int gx = (cx + L - m->start[0]) % L + i, gy = (cy + L - m->start[1]) % L + j;
          if (gx >= 0 && gx < L && gy >= 0 && gy < L) {
            int px = (cx + i + L) % L, py = (cy + j + L) % L;
            float z = m->elevation[(size_t)px * L + py];
            if (z != -10) {
              X[n] = px * m->res; Y[n] = py * m->res; Z[n] = z;
              mx += X[n]; my += Y[n]; mz += Z[n];
              n++;
            }
          }
        }
      if (n > 7) {
        mx /= n; my /= n; mz /= n;
        float M[9] = {0};
        for (int i = 0; i < n; i++) {
          M[0] = fmaf(X[i] - mx, X[i] - mx, M[0]);
          M[4] = fmaf(Y[i] - my, Y[i] - my, M[4]);
          M[8] = fmaf(Z[i] - mz, Z[i] - mz, M[8]);
          M[1] = fmaf(X[i] - mx, Y[i] - my, M[1]);
          M[2] = fmaf(X[i] - mx, Z[i] - mz, M[2]);
          M[5] = fmaf(Y[i] - my, Z[i] - mz, M[5]);
        }
        M[3] = M[1]; M[6] = M[2]; M[7] = M[5];
        float nv[3];
        orc_min_eigvec3(M, nv, 0.01f, 30);
        sl = nv[2] > 0 ? acosf(nv[2]) : acosf(-nv[2]);
        ro = std::fabs(m->elevation[idx] - mz);
        tr = (float)(0.5 * (1.0 - (double)sl / 0.6) + 0.5 * (1.0 - ((double)ro / 0.2)));
      }
      m->traver[idx] = tr;
    }
    if (slope) slope[idx] = sl;
    if (rough) rough[idx] = ro;
    if (traver) traver[idx] = tr;
  }
}

// G_Raytracing + G_Clear_maplowest, gpu_process.cu:706-889, :232-239, with the reference's quirks (see csrc/elev_post.cu header)
void orc_elev_raytracing(void* h) {
  ElevMap* m = (ElevMap*)h;
  const int L = m->L;
  std::vector<float> out(m->elevation);
  auto bound = [&](int cxx, int cyy, int obx, float robot) {
    float x1 = (float)(cxx - obx), x2 = (float)cxx - robot;
    float lo = m->lowest[(size_t)cxx * L + cyy];
    float h2 = m->sensor_z - lo;
    return fmaf(h2 / x2, x1, lo);
  };
  for (int i = 0; i < L * L; i++) {
    if (!(m->traver[i] < m->obstacle_thr && m->elevation[i] != -10)) continue;
    const int cx = i / L, cy = i % L;
    const int ob[2] = {(cx + L - m->start[0]) % L, (cy + L - m->start[1]) % L};
    int cur[2] = {ob[0], ob[1]};
    int robot = (L % 2 == 0) ? (int)(float)(L / 2 - 0.5) : (int)(float)(L / 2);
    float inc[2] = {(float)(ob[0] - robot), (float)(ob[1] - robot)};
    int stepx = inc[0] > 0 ? 1 : (inc[0] == 0 ? 0 : -1), stepy = inc[1] > 0 ? 1 : (inc[1] == 0 ? 0 : -1);
    if (stepx == 0 || stepy == 0) continue;
    float limit = m->elevation[i];
    float dis = sqrtf(fmaf(inc[0], inc[0], inc[1] * inc[1]));
    float dir[2] = {inc[0] / dis, inc[1] / dis};
    float thr = std::fabs(inc[0]) > std::fabs(inc[1]) ? (float)std::sqrt(0.25 + std::pow(0.5 / inc[0] * inc[1], 2))
                                                       : (float)std::sqrt(0.25 + std::pow(0.5 / inc[1] * inc[0], 2));
    float bx = (float)stepx / 2, by = (float)stepy / 2, tx = bx / dir[0], ty = by / dir[1], tprev = 0;
    while (cur[0] >= 0 && cur[0] < L && cur[1] >= 0 && cur[1] < L) {
      bool probe = cur[0] != ob[0] && cur[1] != ob[1] && !(m->lowest[(size_t)cur[0] * L + cur[1]] == 10);
      if (tx > ty) {
        if (ty - tprev > thr && probe) limit = std::min(limit, bound(cur[0], cur[1], ob[0], (float)robot));
        cur[1] += stepy; by += (float)stepy; tprev = ty; ty = by / dir[1];
      } else if (tx < ty) {
        if (tx - tprev > thr && probe) limit = std::min(limit, bound(cur[0], cur[1], ob[0], (float)robot));
        cur[0] += stepx; bx += (float)stepx; tprev = tx; tx = bx / dir[0];
      } else {
        if (tx - tprev > thr && probe) limit = std::min(limit, bound(cur[0], cur[1], ob[0], (float)robot));
        cur[0] += stepx; cur[1] += stepy; bx += (float)stepx; by += (float)stepy; tprev = tx; tx = bx / dir[0]; ty = by / dir[1];
      }
    }
    if (m->elevation[i] - 3 * sqrtf(m->variance[i]) > limit) out[i] = -10;
  }
  m->elevation.swap(out);
  std::fill(m->lowest.begin(), m->lowest.end(), 10.f);
}

void orc_elev_set_owned_rows(void* h, int row_lo, int row_hi) {
  ElevMap* m = (ElevMap*)h;
  m->row_lo = row_lo;
  m->row_hi = row_hi;
}
// plane: 0 elevation, 1 variance, 2 intensity, 3 lowest, 4 traver, 5/6/7 R/G/B; buf = n_rows * L 4-byte elements
void orc_elev_rows_copy(void* h, int plane, int storage_row, int n_rows, void* buf, int to_map) {
  ElevMap* m = (ElevMap*)h;
  void* planes[8] = {m->elevation.data(), m->variance.data(), m->intensity.data(), m->lowest.data(), m->traver.data(),
                     m->R.data(), m->G.data(), m->B.data()};
  char* mp = (char*)planes[plane] + (size_t)storage_row * m->L * 4;
  size_t bytes = (size_t)n_rows * m->L * 4;
  if (to_map) memcpy(mp, buf, bytes);
  else memcpy(buf, mp, bytes);
}

void orc_elev_get_map(void* h, float* elevation, float* variance, float* intensity, int* R, int* G, int* B, float* lowest) {
  ElevMap* m = (ElevMap*)h;
  size_t n = (size_t)m->L * m->L;
  if (elevation) memcpy(elevation, m->elevation.data(), n * 4);
  if (variance) memcpy(variance, m->variance.data(), n * 4);
  if (intensity) memcpy(intensity, m->intensity.data(), n * 4);
  if (R) memcpy(R, m->R.data(), n * 4);
  if (G) memcpy(G, m->G.data(), n * 4);
  if (B) memcpy(B, m->B.data(), n * 4);
  if (lowest) memcpy(lowest, m->lowest.data(), n * 4);
}
void orc_elev_get_state(void* h, float* central2, int* start2) {
  ElevMap* m = (ElevMap*)h;
  central2[0] = m->central[0];
  central2[1] = m->central[1];
  start2[0] = m->start[0];
  start2[1] = m->start[1];
}
}
