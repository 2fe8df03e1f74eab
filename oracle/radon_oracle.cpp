// oracle/radon_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of (paths relative to /root/reference/LoopDetection/):
//   torch-radon/src/forward.cu:12-124          radon_forward_kernel<parallel_beam=true, channels, float>
//   torch-radon/src/texture.cu:135-140         texture: border addressing, linear filter, unnormalised coords
//   torch-radon/torch_radon/radon.py:139-167   ParallelBeam defaults (det_spacing 1, centred volume, voxel 1)
//   src/RING_ros/util.py:192-198               RING: angles linspace(0,2pi,120), normalise, fft2(dim=-2,"ortho")
//   src/RING_ros/util.py:362-374               fast_corr
// The reference samples the image with the GPU texture unit (cudaFilterModeLinear).  The CUDA
// programming guide (appendix "Texture Fetching / Linear Filtering") defines it as
//   tex(x,y) = (1-a)(1-b)T[i,j] + a(1-b)T[i+1,j] + (1-a)b T[i,j+1] + ab T[i+1,j+1],
//   xB = x-0.5, i = floor(xB), a = frac(xB), with a,b held in 9-bit fixed point (8 fractional bits).
// weight_bits = 0 evaluates exact fp32 bilinear weights; weight_bits = 8 quantises a,b to 1/256
// (round to nearest) like the hardware.  fast_trig=1 is reserved for the GPU path's __sinf/__cosf
// (not reproducible on a CPU); the oracle always uses sinf/cosf.
//
// PARITY PIN: torch-radon's own test compares against analytic line integrals with a tolerance of
// 2e-3*(512/A)*(512/D) (tests/test_parallel_beam.py:70) -- loose; tests/ reuses the reference's
// analytic oracle (src/symbolic.cpp, compiled unmodified into oracle/_ref) with that tolerance, and
// on the GPU box the reference kernel itself (forward.cu + texture.cu compiled unmodified into
// oracle/_ref) is run side by side.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

namespace {

inline float texel(const float* img, int W, int H, int i, int j) {  // border addressing -> 0
  return (i >= 0 && i < W && j >= 0 && j < H) ? img[(size_t)j * W + i] : 0.0f;
}

inline float tex2d(const float* img, int W, int H, float x, float y, int weight_bits) {
  float xb = x - 0.5f, yb = y - 0.5f;
  float fi = floorf(xb), fj = floorf(yb);
  float a = xb - fi, b = yb - fj;
  if (weight_bits > 0) {
    float s = (float)(1 << weight_bits);
    a = rintf(a * s) / s;
    b = rintf(b * s) / s;
  }
  int i = (int)fi, j = (int)fj;
  float t00 = texel(img, W, H, i, j), t10 = texel(img, W, H, i + 1, j);
  float t01 = texel(img, W, H, i, j + 1), t11 = texel(img, W, H, i + 1, j + 1);
  float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  return fmaf(w11, t11, fmaf(w01, t01, fmaf(w10, t10, w00 * t00)));
}

}  // namespace

extern "C" {

// img: [batch][H][W]; angles[n_angles]; out: [batch][n_angles][det_count]
void orc_radon_forward(const float* img, int batch, int H, int W, const float* angles, int n_angles, int det_count,
                       float det_spacing, int weight_bits, float* out) {
  // VolumeCfg for Volume2D(): dx = dy = 0, sx = sy = 1 (torch_radon/volumes.py)
  const float vdx = 0.f, vdy = 0.f, vsx = 1.f, vsy = 1.f, inv_sx = 1.f, inv_sy = 1.f;
#pragma omp parallel for collapse(2) schedule(static)
  for (int bz = 0; bz < batch; bz++)
    for (int angle_id = 0; angle_id < n_angles; angle_id++) {
      const float* im = img + (size_t)bz * H * W;
      for (int ray_id = 0; ray_id < det_count; ray_id++) {
        float* o = out + ((size_t)bz * n_angles + angle_id) * det_count + ray_id;
        float sx = (ray_id - det_count * 0.5f + 0.5f) * det_spacing;
        float sy = hypotf(fabsf(vdx) + W * 0.5f, fabsf(vdy) + H * 0.5f);
        float ex = sx, ey = -sy;
        float angle = angles[angle_id];
        // accurate-trig mode of the CUDA path: the double-precision value rounded to float (bit-identical on CPU/GPU
        // up to double rounding); the reference kernel itself uses the approximate __cosf/__sinf intrinsics.
        float cs = (float)cos((double)angle), sn = (float)sin((double)angle);
        float rsx = fmaf(sy, sn, sx * cs);          // sx*cs + sy*sn
        float rsy = fmaf(sy, cs, -sx * sn);         // -sx*sn + sy*cs
        float rdx = fmaf(ey, sn, ex * cs) - rsx;    // ex*cs + ey*sn - rsx
        float rdy = fmaf(ey, cs, -ex * sn) - rsy;   // -ex*sn + ey*cs - rsy
        float vol_orig_x = vdx - 0.5f * W * vsx;
        float vol_orig_y = vdy - 0.5f * H * vsy;
        rsx = (rsx - vol_orig_x) * inv_sx;
        rsy = (rsy - vol_orig_y) * inv_sy;
        rdx = rdx * inv_sx;
        rdy = rdy * inv_sy;
        // max(rdx, 1e-6): double literal -> double max -> float
        float dx = rdx >= 0 ? (float)std::max((double)rdx, 1e-6) : (float)std::min((double)rdx, -1e-6);
        float dy = rdy >= 0 ? (float)std::max((double)rdy, 1e-6) : (float)std::min((double)rdy, -1e-6);
        float alpha_x_m = (-rsx) / dx, alpha_x_p = (W - rsx) / dx;
        float alpha_y_m = (-rsy) / dy, alpha_y_p = (H - rsy) / dy;
        float alpha_s = std::max(std::min(alpha_x_p, alpha_x_m), std::min(alpha_y_p, alpha_y_m));
        float alpha_e = std::min(std::max(alpha_x_p, alpha_x_m), std::max(alpha_y_p, alpha_y_m));
        if ((double)alpha_s > (double)alpha_e - 1e-6) {
          *o = 0.f;
          continue;
        }
        rsx = fmaf(rdx, alpha_s, rsx);
        rsy = fmaf(rdy, alpha_s, rsy);
        rdx *= (alpha_e - alpha_s);
        rdy *= (alpha_e - alpha_s);
        float mx = std::max(fabsf(rdx), fabsf(rdy));
        int n_steps = (int)rintf(mx);
        float vx = rdx / mx, vy = rdy / mx;
        float n = hypotf(vx * vsx, vy * vsy);
        float step;
        if (fabsf(rdy) >= fabsf(rdx)) {
          float y_inc = 0.5f - rsy + rintf(rsy);
          step = y_inc / vy;
          step += vy < 0;
        } else {
          float x_inc = 0.5f - rsx + rintf(rsx);
          step = x_inc / vx;
          step += vx < 0;
        }
        rsx = fmaf(step, vx, rsx);
        rsy = fmaf(step, vy, rsy);
        float acc = 0.f;
        for (int j = 0; j < n_steps; j++) {
          acc += tex2d(im, W, H, rsx, rsy, weight_bits);
          rsx += vx;
          rsy += vy;
        }
        *o = acc * n;
      }
    }
}

// RING tail (util.py:197-198): x = (sino - mean) / std_unbiased over the whole [C,A,D] tensor;
// TIRING[c,k,d] = 1/sqrt(A) * sum_a x[c,a,d] exp(-2 pi i k a / A).  out: interleaved re,im float.
void orc_ring_tiring(const float* sino, int C, int A, int D, float* out_norm, float* out_tiring) {
  size_t n = (size_t)C * A * D;
  double s = 0;
  for (size_t i = 0; i < n; i++) s += sino[i];
  double mean = s / (double)n;
  double ss = 0;
  for (size_t i = 0; i < n; i++) ss += ((double)sino[i] - mean) * ((double)sino[i] - mean);
  double sd = std::sqrt(ss / (double)(n - 1));
  float fmean = (float)mean, fsd = (float)sd;
  std::vector<float> x(n);
  for (size_t i = 0; i < n; i++) x[i] = (sino[i] - fmean) / fsd;
  if (out_norm) memcpy(out_norm, x.data(), n * sizeof(float));
  const double scale = 1.0 / std::sqrt((double)A);
#pragma omp parallel for collapse(2) schedule(static)
  for (int c = 0; c < C; c++)
    for (int k = 0; k < A; k++)
      for (int d = 0; d < D; d++) {
        double re = 0, im = 0;
        for (int a = 0; a < A; a++) {
          double ph = -2.0 * M_PI * (double)((long)k * a % A) / (double)A;
          double v = x[((size_t)c * A + a) * D + d];
          re += v * std::cos(ph);
          im += v * std::sin(ph);
        }
        size_t o = (((size_t)c * A + k) * D + d) * 2;
        out_tiring[o] = (float)(re * scale);
        out_tiring[o + 1] = (float)(im * scale);
      }
}

// fast_corr (util.py:362-374): a,b complex [C,A,D] (interleaved). corr_out[A] (after fftshift), returns via pointers
void orc_ring_fast_corr(const float* a, const float* b, int C, int A, int D, float* corr_out, float* dist, int* angle) {
  std::vector<double> corr(A, 0.0);
  const double scale = 1.0 / std::sqrt((double)A);
  for (int c = 0; c < C; c++)
    for (int d = 0; d < D; d++)
      for (int t = 0; t < A; t++) {
        double re = 0, im = 0;
        for (int k = 0; k < A; k++) {
          size_t o = (((size_t)c * A + k) * D + d) * 2;
          double ar = a[o], ai = a[o + 1], br = b[o], bi = -b[o + 1];  // conj(b)
          double pr = ar * br - ai * bi, pi = ar * bi + ai * br;
          double ph = 2.0 * M_PI * (double)((long)k * t % A) / (double)A;
          double cs = std::cos(ph), sn = std::sin(ph);
          re += pr * cs - pi * sn;
          im += pr * sn + pi * cs;
        }
        corr[t] += std::sqrt(re * re + im * im) * scale;
      }
  // fftshift1d: roll by A/2 (A even): out[i] = corr[(i + A - A/2) % A]  (util.py:330-337 cat(back, front) with n = ceil(A/2))
  int nsh = (A + 1) / 2;
  std::vector<double> sh(A);
  for (int i = 0; i < A; i++) sh[i] = corr[(i + nsh) % A];
  int am = 0;
  for (int i = 1; i < A; i++)
    if ((float)sh[i] > (float)sh[am]) am = i;
  for (int i = 0; i < A; i++) corr_out[i] = (float)sh[i];
  *dist = 1.0f - (float)sh[am] / (float)(0.15 * A * D);
  *angle = -am + A / 2;
}

// solve_translation (util.py:388-423) with the row roll of main_RING.py:171-172: q, p float [C,H,W].
// corr_i[k] = 1/sqrt(W) sum_n q[i][n] p[i'][(n-k) mod W] in double (the reference: float FFTs), |.|, channel sum, fftshift, first
// argmax -> b[i] = W/2 - argmax; then least squares on A = [cos sin](linspace(0,2pi,H) + rot).  out3 = x, y, ||A[x y]-b||.
void orc_ring_solve_translation(const float* q, const float* p, int C, int H, int W, float rot_angle, int row_shift, float* b_out,
                                float* out3) {
  std::vector<double> corr(W);
  std::vector<float> b(H);
  const double scale = 1.0 / std::sqrt((double)W);
  for (int i = 0; i < H; i++) {
    int ip = ((i - row_shift) % H + H) % H;
    std::fill(corr.begin(), corr.end(), 0.0);
    for (int c = 0; c < C; c++) {
      const float* qr = q + ((size_t)c * H + i) * W;
      const float* pr = p + ((size_t)c * H + ip) * W;
      for (int k = 0; k < W; k++) {
        double acc = 0;
        for (int n = 0; n < W; n++) acc += (double)qr[n] * (double)pr[((n - k) % W + W) % W];
        corr[(k + W / 2) % W] += std::fabs(acc * scale);
      }
    }
    int am = 0;
    for (int j = 1; j < W; j++)
      if ((float)corr[j] > (float)corr[am]) am = j;
    b[i] = (float)(W / 2 - am);
    if (b_out) b_out[i] = b[i];
  }
  double scc = 0, sss = 0, scs = 0, scb = 0, ssb = 0;
  std::vector<double> cs(H), sn(H);
  for (int i = 0; i < H; i++) {
    float lin = (i == H - 1) ? (float)(2.0 * M_PI) : (float)((double)i * (2.0 * M_PI / (double)(H - 1)));
    float a = lin + rot_angle;
    cs[i] = (double)cosf(a);
    sn[i] = (double)sinf(a);
    scc += cs[i] * cs[i]; sss += sn[i] * sn[i]; scs += cs[i] * sn[i]; scb += cs[i] * b[i]; ssb += sn[i] * b[i];
  }
  double det = scc * sss - scs * scs;
  double x = (sss * scb - scs * ssb) / det, y = (scc * ssb - scs * scb) / det, err = 0;
  for (int i = 0; i < H; i++) {
    double r = cs[i] * x + sn[i] * y - b[i];
    err += r * r;
  }
  out3[0] = (float)x; out3[1] = (float)y; out3[2] = (float)std::sqrt(err);
}

// GlobalManager::calcRelOri (Mapping/src/global_manager/src/global_manager.cpp:2719-2762).  FFTW (absent third-party dependency;
// FFTW_BACKWARD = un-normalised sum with exp(+2 pi i (hu/H + wv/W))) is restated as a direct double DFT.  Returns the relative angle
// in degrees, *max_loc = flat arg-max of the real part after the cast to float (first maximum, Eigen maxCoeff).
float orc_disco_rel_ori(const float* ra, const float* ia, const float* rb, const float* ib, int H, int W, int* max_loc) {
  std::vector<double> cr((size_t)H * W), ci((size_t)H * W);
  for (int i = 0; i < H * W; i++) {
    cr[i] = (double)(ra[i] * rb[i] + ia[i] * ib[i]);
    ci[i] = (double)(ra[i] * ib[i] + rb[i] * ia[i]);
  }
  // separable: width transform per row, then height transform of each column (real part only)
  std::vector<double> rr((size_t)H * W), ri((size_t)H * W), tcw(W), tsw(W), tch(H), tsh(H);
  for (int w = 0; w < W; w++) { tcw[w] = std::cos(2.0 * M_PI * w / W); tsw[w] = std::sin(2.0 * M_PI * w / W); }
  for (int h = 0; h < H; h++) { tch[h] = std::cos(2.0 * M_PI * h / H); tsh[h] = std::sin(2.0 * M_PI * h / H); }
  for (int h = 0; h < H; h++)
    for (int v = 0; v < W; v++) {
      double re = 0, im = 0;
      for (int w = 0; w < W; w++) {
        int ph = (int)(((long)w * v) % W);
        re += cr[h * W + w] * tcw[ph] - ci[h * W + w] * tsw[ph];
        im += cr[h * W + w] * tsw[ph] + ci[h * W + w] * tcw[ph];
      }
      rr[h * W + v] = re;
      ri[h * W + v] = im;
    }
  float best = -INFINITY;
  int loc = 0;
  for (int u = 0; u < H; u++)
    for (int v = 0; v < W; v++) {
      double re = 0;
      for (int h = 0; h < H; h++) {
        int ph = (h * u) % H;
        re += rr[h * W + v] * tch[ph] - ri[h * W + v] * tsh[ph];
      }
      float val = (float)re;
      if (val > best) { best = val; loc = u * W + v; }
    }
  if (max_loc) *max_loc = loc;
  int hoff = loc / W;
  return (float)(loc - W * hoff) * 3.0f;
}
}
