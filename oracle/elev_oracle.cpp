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
  std::vector<float> elevation, variance, intensity, lowest, traver;
  std::vector<int> R, G, B;

  void init(int length, float resolution, float mahal, float obst) {
    L = length;
    res = resolution;
    mahal_thr = mahal;
    obstacle_thr = obst;
    central[0] = central[1] = 0;
    start[0] = start[1] = 0;
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
      if (ishift[i] >= length) {  // (sic) only the positive overflow clears everything in the reference
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
      map_index[i] = m->points_to_map_index(xt, yt);
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
