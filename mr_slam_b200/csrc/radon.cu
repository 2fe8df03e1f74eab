// radon.cu -- parallel-beam Radon forward projection and the RING descriptor tail / matcher for sm_100a.
//
// Reference (paths relative to LoopDetection/):
//   torch-radon/src/forward.cu:12-124     radon_forward_kernel<parallel_beam=true>  -> k_radon_forward
//   torch-radon/src/texture.cu:114-189    texture upload, border addressing, linear filter (replaced by a shared-
//                                         memory resident image and a software bilinear fetch)
//   src/RING_ros/util.py:197-198          normalise + fft2(dim=-2, "ortho")          -> k_ring_stats + k_ring_tiring
//   src/RING_ros/util.py:362-374          fast_corr                                  -> k_ring_fast_corr
//
// Radon kernel: one block per (image, group of ANGLES_PER_BLOCK angles); the image (H*W*4 B, 57.6 KB for RING)
// is staged once in shared memory with coalesced float4 loads and every ray of the group marches through it
// there -- no texture objects, no cudaArray upload, deterministic fp32.  interp = 1 emulates the texture unit's
// 8-bit fractional weights so results track the reference's hardware-filtered numbers, interp = 0 is exact bilinear.
// On-chip bound (shared-memory gathers); HBM floor per image = read H*W*4 + write A*D*4.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mrslam {

constexpr int ANGLES_PER_BLOCK = 8;
constexpr int RADON_THREADS = 128;

template <bool SMEM>
__device__ __forceinline__ float texel(const float* __restrict__ img, int W, int H, int i, int j) {
  return (i >= 0 && i < W && j >= 0 && j < H) ? img[j * W + i] : 0.0f;  // cudaAddressModeBorder
}

template <bool SMEM>
__device__ __forceinline__ float tex2d(const float* __restrict__ img, int W, int H, float x, float y, int interp) {
  const float xb = x - 0.5f, yb = y - 0.5f;
  const float fi = floorf(xb), fj = floorf(yb);
  float a = xb - fi, b = yb - fj;
  if (interp) {
    a = rintf(a * 256.0f) / 256.0f;
    b = rintf(b * 256.0f) / 256.0f;
  }
  const int i = (int)fi, j = (int)fj;
  const float t00 = texel<SMEM>(img, W, H, i, j), t10 = texel<SMEM>(img, W, H, i + 1, j);
  const float t01 = texel<SMEM>(img, W, H, i, j + 1), t11 = texel<SMEM>(img, W, H, i + 1, j + 1);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  return fmaf(w11, t11, fmaf(w01, t01, fmaf(w10, t10, w00 * t00)));
}

template <bool SMEM>
__global__ void __launch_bounds__(RADON_THREADS) k_radon_forward(const float* __restrict__ img, int H, int W,
                                                                 const float* __restrict__ angles, int n_angles, int det_count,
                                                                 float det_spacing, int interp, int fast_trig,
                                                                 float* __restrict__ out) {
  extern __shared__ float s_img[];
  const int b = blockIdx.y;
  const float* im = img + (size_t)b * H * W;
  if (SMEM) {
    const int n = H * W;
    if ((n & 3) == 0 && ((size_t)im & 15) == 0) {
      const float4* src = reinterpret_cast<const float4*>(im);
      float4* dst = reinterpret_cast<float4*>(s_img);
      for (int i = threadIdx.x; i < n / 4; i += blockDim.x) dst[i] = src[i];
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) s_img[i] = im[i];
    }
    __syncthreads();
    im = s_img;
  }
  const int a0 = blockIdx.x * ANGLES_PER_BLOCK;
  const int a1 = min(a0 + ANGLES_PER_BLOCK, n_angles);
  const float fW = (float)W, fH = (float)H;
  for (int angle_id = a0; angle_id < a1; angle_id++) {
    const float angle = angles[angle_id];
    // accurate mode: double-precision value rounded to float (same bits as the CPU oracle); fast mode: the
    // reference kernel's intrinsics (forward.cu:46-47)
    const float cs = fast_trig ? __cosf(angle) : (float)cos((double)angle);
    const float sn = fast_trig ? __sinf(angle) : (float)sin((double)angle);
    for (int ray_id = threadIdx.x; ray_id < det_count; ray_id += blockDim.x) {
      float* o = out + ((size_t)b * n_angles + angle_id) * det_count + ray_id;
      // forward.cu:33-37 (parallel beam; Volume2D(): dx = dy = 0, sx = sy = 1)
      const float sx = ((float)ray_id - (float)det_count * 0.5f + 0.5f) * det_spacing;
      const float sy = hypotf(fW * 0.5f, fH * 0.5f);
      const float ex = sx, ey = -sy;
      float rsx = fmaf(sy, sn, sx * cs);
      float rsy = fmaf(sy, cs, -sx * sn);
      float rdx = fmaf(ey, sn, ex * cs) - rsx;
      float rdy = fmaf(ey, cs, -ex * sn) - rsy;
      const float vol_orig_x = 0.0f - 0.5f * fW * 1.0f;
      const float vol_orig_y = 0.0f - 0.5f * fH * 1.0f;
      rsx = (rsx - vol_orig_x) * 1.0f;
      rsy = (rsy - vol_orig_y) * 1.0f;
      // clip to the volume, forward.cu:63-80
      const float dx = rdx >= 0 ? (float)fmax((double)rdx, 1e-6) : (float)fmin((double)rdx, -1e-6);
      const float dy = rdy >= 0 ? (float)fmax((double)rdy, 1e-6) : (float)fmin((double)rdy, -1e-6);
      const float alpha_x_m = (-rsx) / dx, alpha_x_p = (fW - rsx) / dx;
      const float alpha_y_m = (-rsy) / dy, alpha_y_p = (fH - rsy) / dy;
      const float alpha_s = fmaxf(fminf(alpha_x_p, alpha_x_m), fminf(alpha_y_p, alpha_y_m));
      const float alpha_e = fminf(fmaxf(alpha_x_p, alpha_x_m), fmaxf(alpha_y_p, alpha_y_m));
      if ((double)alpha_s > (double)alpha_e - 1e-6) {
        *o = 0.0f;
        continue;
      }
      rsx = fmaf(rdx, alpha_s, rsx);
      rsy = fmaf(rdy, alpha_s, rsy);
      rdx *= (alpha_e - alpha_s);
      rdy *= (alpha_e - alpha_s);
      const float mx = fmaxf(fabsf(rdx), fabsf(rdy));
      const int n_steps = __float2int_rn(mx);
      const float vx = rdx / mx, vy = rdy / mx;
      const float nrm = hypotf(vx, vy);
      float step;
      if (fabsf(rdy) >= fabsf(rdx)) {
        const float y_inc = 0.5f - rsy + (float)__float2int_rn(rsy);
        step = y_inc / vy;
        step += (vy < 0) ? 1.0f : 0.0f;
      } else {
        const float x_inc = 0.5f - rsx + (float)__float2int_rn(rsx);
        step = x_inc / vx;
        step += (vx < 0) ? 1.0f : 0.0f;
      }
      rsx = fmaf(step, vx, rsx);
      rsy = fmaf(step, vy, rsy);
      float acc = 0.0f;
      for (int j = 0; j < n_steps; j++) {
        acc += tex2d<SMEM>(im, W, H, rsx, rsy, interp);
        rsx += vx;
        rsy += vy;
      }
      *o = acc * nrm;
    }
  }
}

// mean and unbiased std over each item of n elements (double accumulation); stats[item] = {mean, std}
__global__ void __launch_bounds__(256) k_ring_stats(const float* __restrict__ x, int n, float* __restrict__ stats) {
  __shared__ double s_a[8], s_b[8];
  const float* p = x + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < 8; w++) tot += s_a[w];
  const double mean = tot / (double)n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)p[i] - mean;
    q += d * d;
  }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) s_b[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tq = 0.0;
    for (int w = 0; w < 8; w++) tq += s_b[w];
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)sqrt(tq / (double)(n - 1));
  }
}

// TIRING[c,k,d] = 1/sqrt(A) sum_a ((x[c,a,d]-mean)/std) exp(-2 pi i k a / A); grid = (channels, items), block 256.
// The normalised sinogram of one channel (A*D floats) and the twiddle table live in shared memory.
__global__ void __launch_bounds__(256) k_ring_tiring(const float* __restrict__ sino, const float* __restrict__ stats, int C, int A, int D,
                                                    float* __restrict__ out) {
  extern __shared__ float sm[];
  float* xs = sm;              // A*D
  float* tw_c = sm + A * D;    // A
  float* tw_s = tw_c + A;      // A
  const int item = blockIdx.y, c = blockIdx.x;
  const float mean = stats[2 * item], sd = stats[2 * item + 1];
  const float* p = sino + ((size_t)item * C + c) * A * D;
  for (int i = threadIdx.x; i < A * D; i += blockDim.x) xs[i] = (p[i] - mean) / sd;
  for (int i = threadIdx.x; i < A; i += blockDim.x) {
    double sv, cv;
    sincospi(-2.0 * (double)i / (double)A, &sv, &cv);
    tw_c[i] = (float)cv;
    tw_s[i] = (float)sv;
  }
  __syncthreads();
  const float scale = (float)(1.0 / sqrt((double)A));
  float* o = out + ((size_t)item * C + c) * A * D * 2;
  // r2: the output frequencies of one image are split over gridDim.z blocks (each stages the whole image: 58 KB, L2 hits), so a
  // batch of 256 images gives 1024 blocks instead of 256 on 148 SMs
  const int per = (A + gridDim.z - 1) / gridDim.z, k0 = blockIdx.z * per, k1 = min(A, k0 + per);
  for (int idx = k0 * D + threadIdx.x; idx < k1 * D; idx += blockDim.x) {
    const int k = idx / D, d = idx - k * D;
    float re = 0.f, im = 0.f;
    int ph = 0;
    for (int a = 0; a < A; a++) {
      const float v = xs[a * D + d];
      re = fmaf(v, tw_c[ph], re);
      im = fmaf(v, tw_s[ph], im);
      ph += k;
      if (ph >= A) ph -= A;
    }
    o[2 * (size_t)idx] = re * scale;
    o[2 * (size_t)idx + 1] = im * scale;
  }
}

// fast_corr(query, db[j]) for every j: grid = n_db, block = 256.  corr[t] = sum_{c,d} | 1/sqrt(A) sum_k q*conj(b) e^{+2 pi i k t/A} |
// fast_corr (util.py:362-374), split over detector columns: block (j, z) forms the product a * conj(b) of database descriptor j
// for ITS slice of the detector axis only (r2: the whole 2 x A x D product took 115 KB of shared memory = one 8-warp block per
// SM; a slice of D / FC_SPLIT columns takes 29 KB, so seven blocks are resident), inverse-DFTs it along the angle axis and adds
// its part of sum_d |.| to corr[j][t]; k_ring_corr_argmax then applies fftshift + max.
constexpr int FC_SPLIT = 4;
__global__ void __launch_bounds__(256) k_ring_fast_corr(const float* __restrict__ q, const float* __restrict__ db, int C, int A, int D,
                                                       float* __restrict__ corr_acc /* [n_db][A], zeroed */) {
  extern __shared__ float sm[];
  const int Ds = (D + FC_SPLIT - 1) / FC_SPLIT;
  float* tw_c = sm;            // A
  float* tw_s = sm + A;        // A
  float* corr = sm + 2 * A;    // A
  float* pr = sm + 3 * A;      // A * Ds   product real (one channel at a time)
  float* pi = pr + A * Ds;     // A * Ds
  const int j = blockIdx.x, d0 = blockIdx.y * Ds, dn = min(Ds, D - d0);
  if (dn <= 0) return;
  for (int i = threadIdx.x; i < A; i += blockDim.x) {
    double sv, cv;
    sincospi(2.0 * (double)i / (double)A, &sv, &cv);
    tw_c[i] = (float)cv;
    tw_s[i] = (float)sv;
    corr[i] = 0.f;
  }
  const float scale = (float)(1.0 / sqrt((double)A));
  for (int c = 0; c < C; c++) {
    __syncthreads();
    const float* qa = q + (size_t)c * A * D * 2;
    const float* ba = db + ((size_t)j * C + c) * A * D * 2;
    for (int i = threadIdx.x; i < A * dn; i += blockDim.x) {
      const int a = i / dn, d = i - a * dn;
      const size_t g = (size_t)a * D + d0 + d;
      const float ar = qa[2 * g], ai = qa[2 * g + 1], br = ba[2 * g], bi = -ba[2 * g + 1];
      pr[a * Ds + d] = ar * br - ai * bi;
      pi[a * Ds + d] = fmaf(ar, bi, ai * br);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < A * dn; idx += blockDim.x) {
      const int t = idx / dn, d = idx - t * dn;
      float re = 0.f, im = 0.f;
      int ph = 0;
      for (int k = 0; k < A; k++) {
        const float xr = pr[k * Ds + d], xi = pi[k * Ds + d];
        re += xr * tw_c[ph] - xi * tw_s[ph];
        im += fmaf(xr, tw_s[ph], xi * tw_c[ph]);
        ph += t;
        if (ph >= A) ph -= A;
      }
      atomicAdd(&corr[t], sqrtf(re * re + im * im) * scale);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A; i += blockDim.x) atomicAdd(&corr_acc[(size_t)j * A + i], corr[i]);
}

// fftshift1d (util.py:306-337): roll by ceil(A/2); then max / argmax (first maximum); one warp per database descriptor
__global__ void k_ring_corr_argmax(const float* __restrict__ corr_acc, int n_db, int A, int D, float* __restrict__ dist, int* __restrict__ angle,
                                   float* __restrict__ corr_out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n_db) return;
  const float* corr = corr_acc + (size_t)j * A;
  const int nsh = (A + 1) / 2;
  float best = -INFINITY;
  int am = 0;
  for (int i = lane; i < A; i += 32) {
    const float v = corr[(i + nsh) % A];
    if (corr_out) corr_out[(size_t)j * A + i] = v;
    if (v > best) {  // (first maximum of this lane's positions: they are visited in increasing order)
      best = v;
      am = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(FULL, best, o);
    const int oa = __shfl_xor_sync(FULL, am, o);
    if (ob > best || (ob == best && oa < am)) {
      best = ob;
      am = oa;
    }
  }
  if (lane == 0) {
    dist[j] = 1.0f - best / (float)(0.15 * A * D);
    angle[j] = -am + A / 2;
  }
}

struct RadonScratch {
  DevBuf<float> stats, sino, shifts, corr;
  DevBuf<double> disco;
};
static RadonScratch& rs() {
  static RadonScratch s;
  return s;
}

static int radon_launch(const float* d_img, int batch, int H, int W, const float* d_angles, int n_angles, int det_count,
                        float det_spacing, int interp, int fast_trig, float* d_sino) {
  if (batch <= 0 || n_angles <= 0 || det_count <= 0) return MRSLAM_OK;
  cudaStream_t st = rt().stream;
  dim3 grid((n_angles + ANGLES_PER_BLOCK - 1) / ANGLES_PER_BLOCK, batch);
  size_t smem = (size_t)H * W * sizeof(float);
  if (smem <= 200 * 1024) {
    MR_CUDA(cudaFuncSetAttribute(k_radon_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_radon_forward<true><<<grid, RADON_THREADS, smem, st>>>(d_img, H, W, d_angles, n_angles, det_count, det_spacing, interp, fast_trig, d_sino);
  } else {
    k_radon_forward<false><<<grid, RADON_THREADS, 0, st>>>(d_img, H, W, d_angles, n_angles, det_count, det_spacing, interp, fast_trig, d_sino);
  }
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

static int tiring_launch(const float* d_sino, int items, int C, int A, int D, float* d_tiring) {
  if (items <= 0) return MRSLAM_OK;
  cudaStream_t st = rt().stream;
  RadonScratch& S = rs();
  MR_CHECK(S.stats.reserve((size_t)items * 2));
  k_ring_stats<<<items, 256, 0, st>>>(d_sino, C * A * D, S.stats.p);
  size_t smem = ((size_t)A * D + 2 * A) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("ring_tiring: sinogram %dx%d does not fit shared memory", A, D);
    return MRSLAM_ERR_ARG;
  }
  MR_CUDA(cudaFuncSetAttribute(k_ring_tiring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_ring_tiring<<<dim3(C, items, 1), 256, smem, st>>>(d_sino, S.stats.p, C, A, D, d_tiring);  // (a 4-way split re-stages the image 4x: measured slower)
  count_launch(2);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

}  // namespace mrslam

using namespace mrslam;

// solve_translation (src/RING_ros/util.py:388-423).  For every sinogram row i: circular cross-correlation of the query row with
// the (row-rolled) matched row along the detector axis, magnitude, summed over the channels, fftshift, first argmax ->
// b[i] = W/2 - argmax.  corr[k] = 1/sqrt(W) * sum_n q[n] p[(n - k) mod W] (what ifft(fft(q) conj(fft(p))) with "ortho" norms gives
// for real rows), evaluated directly: W = 120 makes the FFT pointless.  One block per row, one thread per lag.
__global__ void __launch_bounds__(128) k_ring_row_shift(const float* __restrict__ q, const float* __restrict__ p, int C, int H, int W,
                                                       int row_shift, float* __restrict__ b_out) {
  extern __shared__ float sm[];
  float* sq = sm;            // W
  float* sp = sm + W;        // W
  float* corr = sm + 2 * W;  // W (fftshift-ed order)
  const int i = blockIdx.x;
  int ip = (i - row_shift) % H;  // torch.roll(RING_matched, row_shift, dims=1)[i] = RING_matched[(i - row_shift) mod H]
  if (ip < 0) ip += H;
  const float scale = (float)(1.0 / sqrt((double)W));
  for (int k = threadIdx.x; k < W; k += blockDim.x) corr[k] = 0.f;
  for (int c = 0; c < C; c++) {
    __syncthreads();
    for (int n = threadIdx.x; n < W; n += blockDim.x) {
      sq[n] = q[((size_t)c * H + i) * W + n];
      sp[n] = p[((size_t)c * H + ip) * W + n];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < W; k += blockDim.x) {
      float acc = 0.f;
      int m = (W - k) % W;  // (n - k) mod W for n = 0
      for (int n = 0; n < W; n++) {
        acc = fmaf(sq[n], sp[m], acc);
        if (++m == W) m = 0;
      }
      // torch.fft.fftshift: out[j] = in[(j - W/2) mod W]  <=>  in[k] lands at j = (k + W/2) mod W
      corr[(k + W / 2) % W] += fabsf(acc * scale);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int am = 0;
    float best = corr[0];
    for (int j = 1; j < W; j++)
      if (corr[j] > best) {
        best = corr[j];
        am = j;
      }
    b_out[i] = (float)(W / 2 - am);
  }
}

// least squares of A [x y]^T = b with A = [cos(a_i) sin(a_i)], a_i = linspace(0, 2 pi, H)_i (float32, as numpy) + rot_angle
// (util.py:391-399, :488-506 `svd` pseudo-inverse == the normal equations for this full-rank 2-column system); double sums.
// out[0..2] = x, y, ||A [x y]^T - b||.
__global__ void k_ring_translation_lsq(const float* __restrict__ b, int H, float rot_angle, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double scc = 0, sss = 0, scs = 0, scb = 0, ssb = 0;
  for (int i = 0; i < H; i++) {
    // np.linspace(0, 2*pi, H).astype(float32): i * step in double (last element = stop exactly), then the cast
    const float lin = (i == H - 1) ? (float)(2.0 * M_PI) : (float)((double)i * (2.0 * M_PI / (double)(H - 1)));
    const float a = lin + rot_angle;
    const double c = (double)cosf(a), s = (double)sinf(a);
    scc += c * c; sss += s * s; scs += c * s; scb += c * (double)b[i]; ssb += s * (double)b[i];
  }
  const double det = scc * sss - scs * scs;
  const double x = (sss * scb - scs * ssb) / det, y = (scc * ssb - scs * scb) / det;
  double err = 0;
  for (int i = 0; i < H; i++) {
    const float lin = (i == H - 1) ? (float)(2.0 * M_PI) : (float)((double)i * (2.0 * M_PI / (double)(H - 1)));
    const float a = lin + rot_angle;
    const double r = (double)cosf(a) * x + (double)sinf(a) * y - (double)b[i];
    err += r * r;
  }
  out[0] = (float)x;
  out[1] = (float)y;
  out[2] = (float)sqrt(err);
}

// GlobalManager::calcRelOri (Mapping/src/global_manager/src/global_manager.cpp:2719-2762): yaw between two DiSCO signatures.
// Cross spectrum with the reference's (non-conjugate) imaginary part, un-normalised inverse 2-D DFT (FFTW_BACKWARD, double),
// first arg-max of the real part, column offset * 3 degrees.  (The reference's r0 normalisation lines :2737-2738 have no effect.)
// The 40 x 120 transform is evaluated directly and separably in double: stage 1 one block per (pair, row) does the width
// transform, stage 2 one block per pair does the height transform of the real part and the arg-max.
__global__ void __launch_bounds__(128) k_disco_rows(const float* __restrict__ ra, const float* __restrict__ ia, const float* __restrict__ rb,
                                                   const float* __restrict__ ib, int H, int W, double* __restrict__ row_re,
                                                   double* __restrict__ row_im) {
  extern __shared__ double smd[];
  double* cr = smd;
  double* ci = smd + W;
  double* tc = smd + 2 * W;
  double* ts = smd + 3 * W;
  const int pair = blockIdx.y, h = blockIdx.x;
  const float* rbp = rb + (size_t)pair * H * W;
  const float* ibp = ib + (size_t)pair * H * W;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const float a_r = ra[h * W + w], a_i = ia[h * W + w], b_r = rbp[h * W + w], b_i = ibp[h * W + w];
    cr[w] = (double)(a_r * b_r + a_i * b_i);  // float products and sums, then widened: real_a[i] * real_b[i] + ... is float math
    ci[w] = (double)(a_r * b_i + b_r * a_i);
    double sv, cv;
    sincospi(2.0 * (double)w / (double)W, &sv, &cv);
    tc[w] = cv;
    ts[w] = sv;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < W; v += blockDim.x) {
    double re = 0, im = 0;
    int ph = 0;
    for (int w = 0; w < W; w++) {
      re += cr[w] * tc[ph] - ci[w] * ts[ph];
      im += cr[w] * ts[ph] + ci[w] * tc[ph];
      ph += v;
      if (ph >= W) ph -= W;
    }
    row_re[((size_t)pair * H + h) * W + v] = re;
    row_im[((size_t)pair * H + h) * W + v] = im;
  }
}

__global__ void __launch_bounds__(256) k_disco_cols_argmax(const double* __restrict__ row_re, const double* __restrict__ row_im, int H, int W,
                                                          float* __restrict__ rel_angle, int* __restrict__ max_loc) {
  __shared__ float s_best[256];
  __shared__ int s_loc[256];
  const int pair = blockIdx.x;
  const double* rr = row_re + (size_t)pair * H * W;
  const double* ri = row_im + (size_t)pair * H * W;
  float best = -INFINITY;
  int loc = 0x7fffffff;
  for (int o = threadIdx.x; o < H * W; o += blockDim.x) {
    const int u = o / W, v = o - u * W;
    double re = 0;
    for (int h = 0; h < H; h++) {
      double sv, cv;
      sincospi(2.0 * (double)((h * u) % H) / (double)H, &sv, &cv);
      re += rr[h * W + v] * cv - ri[h * W + v] * sv;
    }
    const float val = (float)re;  // Eigen::VectorXf cross_real
    if (val > best || (val == best && o < loc)) {
      best = val;
      loc = o;
    }
  }
  s_best[threadIdx.x] = best;
  s_loc[threadIdx.x] = loc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) {
      const float ob = s_best[threadIdx.x + st];
      const int ol = s_loc[threadIdx.x + st];
      if (ob > s_best[threadIdx.x] || (ob == s_best[threadIdx.x] && ol < s_loc[threadIdx.x])) {
        s_best[threadIdx.x] = ob;
        s_loc[threadIdx.x] = ol;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int ml = s_loc[0];
    const int hoff = ml / W;
    const int woff = ml - W * hoff;
    rel_angle[pair] = (float)woff * 3.0f;
    if (max_loc) max_loc[pair] = ml;
  }
}

extern "C" {

int mrslam_radon_forward(const float* d_img, int batch, int height, int width, const float* d_angles, int n_angles, int det_count,
                         float det_spacing, int interp, int fast_trig, float* d_sino) {
  MR_CHECK(ensure_init());
  std::lock_guard<std::mutex> lk(rt().mu);
  return radon_launch(d_img, batch, height, width, d_angles, n_angles, det_count, det_spacing, interp, fast_trig, d_sino);
}

int mrslam_ring_tiring(const float* d_sino, int items, int channels, int n_angles, int det_count, float* d_tiring) {
  MR_CHECK(ensure_init());
  std::lock_guard<std::mutex> lk(rt().mu);
  return tiring_launch(d_sino, items, channels, n_angles, det_count, d_tiring);
}

int mrslam_ring_fast_corr(const float* d_query, const float* d_db, int n_db, int channels, int n_angles, int det_count, float* d_dist,
                          int* d_angle, float* d_corr_opt) {
  MR_CHECK(ensure_init());
  if (n_db <= 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  const int Ds = (det_count + FC_SPLIT - 1) / FC_SPLIT;
  size_t smem = ((size_t)3 * n_angles + 2 * (size_t)n_angles * Ds) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("ring_fast_corr: descriptor %dx%d does not fit shared memory", n_angles, det_count);
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(rs().corr.reserve((size_t)n_db * n_angles));
  MR_CUDA(cudaMemsetAsync(rs().corr.p, 0, sizeof(float) * (size_t)n_db * n_angles, rt().stream));
  MR_CUDA(cudaFuncSetAttribute(k_ring_fast_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_ring_fast_corr<<<dim3(n_db, FC_SPLIT), 256, smem, rt().stream>>>(d_query, d_db, channels, n_angles, det_count, rs().corr.p);
  k_ring_corr_argmax<<<(n_db + 3) / 4, 128, 0, rt().stream>>>(rs().corr.p, n_db, n_angles, det_count, d_dist, d_angle, d_corr_opt);
  count_launch(2);
  count_launch(1);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

int mrslam_ring_generate(const float* d_bev, int batch, int size, const float* d_angles, int n_angles, int interp, int fast_trig,
                         float* d_sino_opt, float* d_tiring) {
  MR_CHECK(ensure_init());
  std::lock_guard<std::mutex> lk(rt().mu);
  float* sino = d_sino_opt;
  if (!sino) {
    MR_CHECK(rs().sino.reserve((size_t)batch * n_angles * size));
    sino = rs().sino.p;
  }
  // generate_RING: ParallelBeam(num_sector, angles).forward(bev[1, ring, sector]) -> [1, n_angles, det = size]
  MR_CHECK(radon_launch(d_bev, batch, size, size, d_angles, n_angles, size, 1.0f, interp, fast_trig, sino));
  return tiring_launch(sino, batch, 1, n_angles, size, d_tiring);
}
int mrslam_ring_solve_translation(const float* d_query, const float* d_positive, int channels, int n_angles, int det_count, float rot_angle,
                                  int row_shift, float* d_row_shift_opt, float* xy_err3) {
  MR_CHECK(ensure_init());
  if (channels <= 0 || n_angles < 2 || det_count <= 0 || !d_query || !d_positive || !xy_err3) {
    set_error("ring_solve_translation: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(rt().mu);
  MR_CHECK(rs().shifts.reserve((size_t)n_angles + 4));
  float* b = d_row_shift_opt ? d_row_shift_opt : rs().shifts.p;
  float* out = rs().shifts.p + n_angles;
  k_ring_row_shift<<<n_angles, 128, 3 * (size_t)det_count * sizeof(float), rt().stream>>>(d_query, d_positive, channels, n_angles, det_count,
                                                                                         row_shift, b);
  k_ring_translation_lsq<<<1, 32, 0, rt().stream>>>(b, n_angles, rot_angle, out);
  count_launch(2);
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaMemcpyAsync(xy_err3, out, 3 * sizeof(float), cudaMemcpyDeviceToHost, rt().stream));
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  return MRSLAM_OK;
}
int mrslam_disco_rel_ori(const float* d_real_a, const float* d_imag_a, const float* d_real_b, const float* d_imag_b, int n_b, int height,
                         int width, float* d_rel_angle, int* d_max_loc_opt) {
  MR_CHECK(ensure_init());
  if (n_b < 0 || height <= 0 || width <= 0 || (n_b > 0 && (!d_real_a || !d_imag_a || !d_real_b || !d_imag_b || !d_rel_angle))) {
    set_error("disco_rel_ori: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  if (n_b == 0) return MRSLAM_OK;
  std::lock_guard<std::mutex> lk(rt().mu);
  const size_t plane = (size_t)n_b * height * width;
  MR_CHECK(rs().disco.reserve(2 * plane));
  double* re = rs().disco.p;
  double* im = rs().disco.p + plane;
  k_disco_rows<<<dim3(height, n_b), 128, 4 * (size_t)width * sizeof(double), rt().stream>>>(d_real_a, d_imag_a, d_real_b, d_imag_b, height, width,
                                                                                         re, im);
  k_disco_cols_argmax<<<n_b, 256, 0, rt().stream>>>(re, im, height, width, d_rel_angle, d_max_loc_opt);
  count_launch(2);
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}
}
