// gicp.cu -- host side of the FastGICP path: device-resident clouds, the lock-step batched
// Levenberg-Marquardt driver and the C-ABI entry points (include/mrslam_b200.h).
//
// Reference being replaced (paths relative to Mapping/src/fast_gicp/):
//   include/fast_gicp/gicp/impl/fast_gicp_impl.hpp        (covariances, correspondences, linearize, error)
//   include/fast_gicp/gicp/impl/lsq_registration_impl.hpp (LM / GN driver, convergence test)
//   include/fast_gicp/so3/so3.hpp                         (so3_exp)
// Host algebra (6x6 LDLT, so3_exp) is written here from the published formulas; it is not shared
// with oracle/.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

#include "gicp_kernels.cuh"

namespace mrslam {

// ---------------------------------------------------------------------------------------------
// device-resident cloud
// ---------------------------------------------------------------------------------------------
// ordering between host uploads (copy stream) and their unpack kernels (library stream), shared by all clouds
struct UploadSync {
  cudaEvent_t alloc_ev = nullptr;  // a staging buffer allocated on the library stream exists
  cudaEvent_t pack_ev = nullptr;   // recorded behind every unpack launch: staging buffers unpacked so far may be overwritten
  bool pack_recorded = false;
  unsigned long long seq = 0;
};
static UploadSync& upload_sync() {
  static UploadSync u;
  return u;
}

struct Cloud {
  int n = 0, npad = 0, ntiles = 0, ntiles_pad = 0, nsuper = 0, nsuper_pad = 0;
  DevBuf<float4> porig, psort;
  DevBuf<float> tb, sb, qp;
  DevBuf<unsigned long long> skeys;
  DevBuf<double> cov;
  DevBuf<int> knn_ids;  // [npad*k] neighbour lists (scratch between k_knn_search and k_cov)
  DevBuf<unsigned char> knn_fb;  // [ntiles] tiles k_knn_search_v2 hands back to k_knn_search
  bool cov_valid = false;
  bool built = false;   // spatial structures (psort / tb / sb) are produced lazily, batched with the other clouds of a call
  DevBuf<float> raw;    // staging of host input
  // Uploads run on the runtime's copy stream and are only ENQUEUED at setInput time: the library stream neither waits for them
  // nor unpacks them before the cloud is first used (build_clouds() waits for the newest upload of its batch once and unpacks
  // all of them with one launch).  A caller can therefore hand over the clouds of the next batch while the current one is still
  // being aligned; their uploads run under its kernels.  ev_up = this cloud's upload has landed (copy-stream order).
  cudaEvent_t ev_up = nullptr;
  int pending_stride = 0;  // > 0: raw holds an upload that has not been unpacked into porig yet
  unsigned long long upload_seq = 0;  // order of the uploads on the copy stream
  int refs = 1;

  // stage 1 (at setInput time): points into porig (device).  The Morton sort etc. is deferred to build_clouds().
  int set_sizes(int n_) {
    n = n_;
    npad = ((n + TILE - 1) / TILE) * TILE;
    ntiles = npad / TILE;
    ntiles_pad = ((ntiles + 31) / 32) * 32;
    if (ntiles_pad == 0) ntiles_pad = 32;
    nsuper = (ntiles + 31) / 32;
    nsuper_pad = ((nsuper + 31) / 32) * 32;
    if (nsuper_pad == 0) nsuper_pad = 32;
    cov_valid = false;
    built = false;
    pending_stride = 0;
    return porig.reserve(std::max(n, 1));
  }

  int set_from_device_raw(const float* d_raw, int n_, int stride) {
    cudaStream_t st = rt().stream;
    pending_stride = 0;
    n = n_;
    npad = ((n + TILE - 1) / TILE) * TILE;
    ntiles = npad / TILE;
    ntiles_pad = ((ntiles + 31) / 32) * 32;
    if (ntiles_pad == 0) ntiles_pad = 32;
    nsuper = (ntiles + 31) / 32;
    nsuper_pad = ((nsuper + 31) / 32) * 32;
    if (nsuper_pad == 0) nsuper_pad = 32;
    cov_valid = false;
    built = false;
    MR_CHECK(porig.reserve(std::max(n, 1)));
    if (n > 0) {
      ProfScope ps(PROF_CLOUD_BUILD);
      k_pack_points<<<(n + 255) / 256, 256, 0, st>>>(d_raw, n, stride, porig.p);
      count_launch(1);
      MR_CUDA(cudaGetLastError());
    }
    return MRSLAM_OK;
  }

  int set_from_host(const float* xyz, int n_, int stride) {
    if (n_ < 0 || stride < 3 || (n_ > 0 && !xyz)) {
      set_error("cloud: bad arguments (n=%d stride=%d)", n_, stride);
      return MRSLAM_ERR_ARG;
    }
    size_t nf = (size_t)n_ * stride;
    const float* before = raw.p;
    // an upload that was never consumed may still be in flight into the staging buffer that reserve() is about to free
    if (pending_stride > 0 && ev_up && raw.cap < std::max<size_t>(nf, 1)) MR_CUDA(cudaStreamWaitEvent(rt().stream, ev_up, 0));
    MR_CHECK(raw.reserve(std::max<size_t>(nf, 1)));
    MR_CHECK(set_sizes(n_));
    if (n_ > 0) {
      cudaStream_t up = rt().copy_stream, st = rt().stream;
      UploadSync& U = upload_sync();
      if (!ev_up) MR_CUDA(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
      if (!U.alloc_ev) {
        MR_CUDA(cudaEventCreateWithFlags(&U.alloc_ev, cudaEventDisableTiming));
        MR_CUDA(cudaEventCreateWithFlags(&U.pack_ev, cudaEventDisableTiming));
      }
      // the staging buffer must exist (a fresh allocation lives in library-stream order) and must no longer be read: its last
      // reader was the unpack kernel of an earlier batch, and every such kernel precedes the newest pack_ev record
      if (raw.p != before) {
        MR_CUDA(cudaEventRecord(U.alloc_ev, st));
        MR_CUDA(cudaStreamWaitEvent(up, U.alloc_ev, 0));
      } else if (U.pack_recorded) {
        MR_CUDA(cudaStreamWaitEvent(up, U.pack_ev, 0));
      }
      size_t bytes = ((size_t)(n_ - 1) * stride + 3) * sizeof(float);
      MR_CUDA(cudaMemcpyAsync(raw.p, xyz, bytes, cudaMemcpyHostToDevice, up));
      MR_CUDA(cudaEventRecord(ev_up, up));
      upload_seq = ++U.seq;
      pending_stride = stride;
    }
    return MRSLAM_OK;
  }

  void release() {
    if (pending_stride > 0 && ev_up) cudaStreamWaitEvent(rt().stream, ev_up, 0);  // frees are library-stream ordered
    pending_stride = 0;
    if (ev_up) cudaEventDestroy(ev_up);
    ev_up = nullptr;
    porig.release(); psort.release(); tb.release(); sb.release(); qp.release(); skeys.release(); cov.release(); knn_ids.release(); knn_fb.release(); raw.release();
  }
};

// Device buffers of destroyed clouds / registrations are recycled instead of cudaFreeAsync'ed: the reference constructs a
// fresh FastGICP (and fresh clouds) for every ICPCheck call, and ~25 stream-ordered frees + allocations per object were
// measured to cost ~7 ms per 32-pair batch on the stream.
static std::vector<Cloud*>& cloud_pool() {
  static std::vector<Cloud*> p;
  return p;
}
static Cloud* acquire_cloud() {
  std::vector<Cloud*>& p = cloud_pool();
  if (!p.empty()) {
    Cloud* c = p.back();
    p.pop_back();
    c->refs = 1;
    c->cov_valid = false;
    c->built = false;
    c->n = 0;
    return c;
  }
  return new Cloud();
}

// scratch of the batched build (process-global, grow-only)
struct BuildScratch {
  DevBuf<unsigned long long> keys, keys_alt;
  DevBuf<int> vals, vals_alt;
  DevBuf<unsigned char> cubtmp;
  DevBuf<unsigned> bbox;
  DevBuf<BuildDesc> d_desc;
  PinBuf<BuildDesc> h_desc;
  StageFence fence;
};
static BuildScratch& build_scratch() {
  static BuildScratch s;
  return s;
}

// stage 2: Morton order, tiles and boxes of every not-yet-built cloud in `clouds`, ONE pass over the whole batch:
// per-cloud bounding cube -> (cloud id, 48-bit Morton) keys -> one radix sort -> gather + tile boxes -> super-tile boxes.
// development aid (MRSLAM_BUILD_DIAG=1): device time of the stages of build_clouds, summed over the process, printed at exit
struct BuildDiag {
  bool on = getenv("MRSLAM_BUILD_DIAG") != nullptr;
  cudaEvent_t ev[6] = {};
  double ms[5] = {};
  double host_ms = 0.0;
  long calls = 0;
  void mark(int i, cudaStream_t st) {
    if (!on) return;
    if (!ev[i]) cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[i], st);
  }
  void collect() {
    if (!on) return;
    cudaEventSynchronize(ev[5]);
    for (int i = 0; i < 5; i++) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      ms[i] += t;
    }
    calls++;
  }
  void dump(const char* label) {
    if (on && calls)
      fprintf(stderr, "build diag [%s]: %ld calls; ms per call: host-setup-to-first-kernel %.3f | wait+unpack %.3f | bbox+keys %.3f | sort %.3f | gather+boxes %.3f\n",
              label ? label : "", calls, ms[0] / calls, ms[1] / calls, ms[2] / calls, ms[3] / calls, ms[4] / calls);
    for (double& m : ms) m = 0.0;
    calls = 0;
  }
};
static BuildDiag& build_diag() {
  static BuildDiag d;
  return d;
}

static int build_clouds(const std::vector<Cloud*>& clouds) {
  std::vector<Cloud*> todo;
  for (Cloud* c : clouds)
    if (c && !c->built && std::find(todo.begin(), todo.end(), c) == todo.end()) todo.push_back(c);
  if (todo.empty()) return MRSLAM_OK;
  cudaStream_t st = rt().stream;
  BuildScratch& S = build_scratch();
  ProfScope ps(PROF_CLOUD_BUILD);
  BuildDiag& D = build_diag();
  D.mark(0, st);
  const int nc = (int)todo.size();
  if (nc >= (1 << 15)) {
    set_error("too many clouds in one batch (%d)", nc);
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(S.fence.wait());
  MR_CHECK(S.h_desc.reserve(nc));
  MR_CHECK(S.d_desc.reserve(nc));
  long long total = 0;
  int max_n = 0, max_tiles_pad = 0, max_super_pad = 0;
  Cloud* newest_upload = nullptr;  // uploads complete in copy-stream order: waiting for the newest one covers the batch
  for (int i = 0; i < nc; i++) {
    Cloud* c = todo[i];
    MR_CHECK(c->psort.reserve(std::max(c->npad, 1)));
    MR_CHECK(c->tb.reserve((size_t)6 * c->ntiles_pad));
    MR_CHECK(c->sb.reserve((size_t)6 * c->nsuper_pad));
    MR_CHECK(c->qp.reserve(4));
    MR_CHECK(c->skeys.reserve(std::max(c->n, 1)));
    BuildDesc d;
    d.skeys = c->skeys.p;
    d.porig = c->porig.p; d.psort = c->psort.p; d.tb = c->tb.p; d.sb = c->sb.p; d.qp = c->qp.p;
    d.n = c->n; d.ntiles = c->ntiles; d.ntiles_pad = c->ntiles_pad; d.nsuper_pad = c->nsuper_pad;
    d.offset = total;
    d.raw = nullptr;
    d.raw_stride = 0;
    if (c->pending_stride > 0 && c->n > 0) {
      d.raw = c->raw.p;
      d.raw_stride = c->pending_stride;
      if (!newest_upload || c->upload_seq > newest_upload->upload_seq) newest_upload = c;
    }
    c->pending_stride = 0;
    S.h_desc.p[i] = d;
    total += c->n;
    max_n = std::max(max_n, c->n);
    max_tiles_pad = std::max(max_tiles_pad, c->ntiles_pad);
    max_super_pad = std::max(max_super_pad, c->nsuper_pad);
  }
  MR_CUDA(upload_params(S.d_desc.p, S.h_desc.p, sizeof(BuildDesc) * nc, st));
  MR_CHECK(S.fence.mark(st));
  D.mark(1, st);
  if (newest_upload) {
    UploadSync& U = upload_sync();
    MR_CUDA(cudaStreamWaitEvent(st, newest_upload->ev_up, 0));
    k_pack_pending<<<dim3(std::max(1, std::min((max_n + 255) / 256, 128)), nc), 256, 0, st>>>(S.d_desc.p);
    count_launch(1);
    MR_CUDA(cudaEventRecord(U.pack_ev, st));
    U.pack_recorded = true;
  }
  size_t tot = (size_t)std::max<long long>(total, 1);
  MR_CHECK(S.keys.reserve(tot)); MR_CHECK(S.keys_alt.reserve(tot));
  MR_CHECK(S.vals.reserve(tot)); MR_CHECK(S.vals_alt.reserve(tot));
  MR_CHECK(S.bbox.reserve((size_t)6 * nc));
  const int* order = S.vals.p;
  const unsigned long long* keys_sorted = S.keys.p;
  D.mark(2, st);
  if (total > 0) {
    k_bbox_init<<<(nc * 6 + 127) / 128, 128, 0, st>>>(S.bbox.p, nc);
    int gx = std::max(1, std::min((max_n + 255) / 256, 32));
    k_bbox<<<dim3(gx, nc), 256, 0, st>>>(S.d_desc.p, S.bbox.p);
    k_morton<<<dim3((max_n + 255) / 256, nc), 256, 0, st>>>(S.d_desc.p, S.bbox.p, S.keys.p, S.vals.p);
    count_launch(3);
    D.mark(3, st);
    int bits = 48;
    while ((1 << (bits - 48)) < nc) bits++;
    cub::DoubleBuffer<unsigned long long> dk(S.keys.p, S.keys_alt.p);
    cub::DoubleBuffer<int> dv(S.vals.p, S.vals_alt.p);
    size_t tmp = 0;
    MR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)total, KEY_DROP_BITS, bits, st));
    MR_CHECK(S.cubtmp.reserve(tmp));
    MR_CUDA(cub::DeviceRadixSort::SortPairs(S.cubtmp.p, tmp, dk, dv, (int)total, KEY_DROP_BITS, bits, st));
    order = dv.Current();
    keys_sorted = dk.Current();
  }
  D.mark(4, st);
  k_gather_tiles<<<dim3((max_tiles_pad * 32 + 127) / 128, nc), 128, 0, st>>>(S.d_desc.p, order, keys_sorted);
  k_super_boxes<<<dim3((max_super_pad * 32 + 127) / 128, nc), 128, 0, st>>>(S.d_desc.p);
  count_launch(2);
  MR_CUDA(cudaGetLastError());
  D.mark(5, st);
  D.collect();
  for (Cloud* c : todo) c->built = true;
  return MRSLAM_OK;
}

static bool scratch_stats_enabled();
static int dispatch_knn(int k, const KnnParams* d_params, int nproblems, int max_tiles, bool with_cov) {
  if (k < 1 || k > 32) {
    set_error("k_correspondences must be in 1..32 (got %d)", k);
    return MRSLAM_ERR_ARG;
  }
  if (max_tiles <= 0 || nproblems <= 0) return MRSLAM_OK;
  dim3 grid((max_tiles + 3) / 4, nproblems);
  {
    ProfScope ps(PROF_KNN_SEARCH);
    const bool v2 = k < KP && getenv("MRSLAM_KNN_V2");  // opt-in until the round-2 GPU comparison is in (gpurun_out/r2f)
    if (v2) {
      // k <= 15: staged lane-per-query-list kernel; the (rare) blocks it flags are redone by the general kernel
      const dim3 grid2(max_tiles, nproblems);  // one warp per query tile
      if (scratch_stats_enabled()) k_knn_search_v2<true><<<grid2, 32, 0, rt().stream>>>(d_params);
      else k_knn_search_v2<false><<<grid2, 32, 0, rt().stream>>>(d_params);
      count_launch(1);
    }
    static const int knn_wpb = getenv("MRSLAM_KNN_WPB") ? std::max(1, std::min(4, atoi(getenv("MRSLAM_KNN_WPB")))) : 4;
    const dim3 g1((max_tiles + knn_wpb - 1) / knn_wpb, nproblems);
    if (scratch_stats_enabled()) k_knn_search<true><<<g1, 32 * knn_wpb, 0, rt().stream>>>(d_params, v2 ? 1 : 0);
    else k_knn_search<false><<<g1, 32 * knn_wpb, 0, rt().stream>>>(d_params, v2 ? 1 : 0);
  }
  count_launch(1);
  if (with_cov) {
    ProfScope ps(PROF_COV);
    k_cov<<<dim3((max_tiles * TILE + 127) / 128, nproblems), 128, 0, rt().stream>>>(d_params);
    count_launch(1);
  }
  MR_CUDA(cudaGetLastError());
  return MRSLAM_OK;
}

// ---------------------------------------------------------------------------------------------
// small host algebra (double)
// ---------------------------------------------------------------------------------------------
struct Iso {
  double R[9];
  double t[3];
};
static Iso iso_identity() {
  Iso x;
  for (int i = 0; i < 9; i++) x.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  x.t[0] = x.t[1] = x.t[2] = 0;
  return x;
}
static Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
    r.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
  }
  return r;
}
// so3.hpp:58-77 (quaternion exponential) followed by Quaterniond::toRotationMatrix
static void so3_exp(const double om[3], double R[9]) {
  double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    double tq = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * tq;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * tq;
  } else {
    double theta = std::sqrt(theta_sq), half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  const double w = real, x = imag * om[0], y = imag * om[1], z = imag * om[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
               tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// symmetric 6x6 solve by LDL^T with diagonal pivoting (what Eigen::LDLT does)
static bool ldlt_solve6(const double Ain[36], const double rhs[6], double x[6]) {
  double A[36];
  memcpy(A, Ain, sizeof(A));
  int perm[6] = {0, 1, 2, 3, 4, 5};
  double L[36] = {0}, D[6];
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = std::fabs(A[7 * k]);
    for (int i = k + 1; i < 6; i++)
      if (std::fabs(A[7 * i]) > best) {
        best = std::fabs(A[7 * i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < 6; j++) std::swap(A[k * 6 + j], A[piv * 6 + j]);
      for (int j = 0; j < 6; j++) std::swap(A[j * 6 + k], A[j * 6 + piv]);
      for (int j = 0; j < k; j++) std::swap(L[k * 6 + j], L[piv * 6 + j]);
      std::swap(perm[k], perm[piv]);
    }
    D[k] = A[7 * k];
    L[7 * k] = 1.0;
    if (D[k] == 0.0 || !std::isfinite(D[k])) return false;
    for (int i = k + 1; i < 6; i++) L[i * 6 + k] = A[i * 6 + k] / D[k];
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j < 6; j++) A[i * 6 + j] -= L[i * 6 + k] * D[k] * L[j * 6 + k];
  }
  double y[6], z[6];
  for (int i = 0; i < 6; i++) y[i] = rhs[perm[i]];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < i; j++) y[i] -= L[i * 6 + j] * y[j];
  for (int i = 0; i < 6; i++) z[i] = y[i] / D[i];
  for (int i = 5; i >= 0; i--)
    for (int j = i + 1; j < 6; j++) z[i] -= L[j * 6 + i] * z[j];
  for (int i = 0; i < 6; i++) x[perm[i]] = z[i];
  return true;
}

// ---------------------------------------------------------------------------------------------
// registration handle
// ---------------------------------------------------------------------------------------------
struct Gicp {
  // parameters: fast_gicp_impl.hpp:9-24, lsq_registration_impl.hpp:9-22, pcl::Registration defaults
  int num_threads = 0;
  int k_corr = 20;
  int regularization = MRSLAM_REG_PLANE;
  double corr_dist_threshold = std::sqrt(std::numeric_limits<double>::max());
  int max_iterations = 64;
  double rotation_epsilon = 2e-3;
  double transformation_epsilon = 5e-4;
  int optimizer = MRSLAM_OPT_LM;
  int lm_max_iterations = 10;
  double lm_init_lambda_factor = 1e-9;
  bool debug_print = false;

  Cloud* src = nullptr;
  Cloud* tgt = nullptr;

  // per-source-point state of the last linearisation (sorted source order); recycled through pair_pool()
  struct PairBufs {
    DevBuf<int> corr, nnpos;
    DevBuf<float> sqd;
    DevBuf<double> mahal, partials;
  };
  PairBufs* pb = nullptr;
  bool have_seed = false;
  bool have_corr = false;

  // results
  Iso x0 = iso_identity();
  double final_H[36];
  bool converged = false;
  int nr_iterations = 0;
  int n_linearize = 0, n_error = 0;
  bool lm_failed = false;

  // LM scratch (batch driver)
  double lm_lambda = -1.0, nu = 2.0, y0 = 0.0;
  double H[36], b[6], d[6];
  Iso delta, xi;
  int lm_try = 0;
  int phase = 0;

  Gicp() {
    for (int i = 0; i < 36; i++) final_H[i] = (i % 7 == 0) ? 1.0 : 0.0;
  }
};

static void cloud_unref(Cloud* c) {
  if (!c) return;
  if (--c->refs == 0) {
    if (cloud_pool().size() < 1024) {
      cloud_pool().push_back(c);  // keep the device buffers for the next cloud
    } else {
      c->release();
      delete c;
    }
  }
}

// Shared pinned/device scratch of the batch driver (process-global: the reference creates a fresh
// FastGICP per ICPCheck call, global_manager.cpp:2437, so nothing here may be per-object).
struct BatchScratch {
  DevBuf<PairParams> d_pairs;
  DevBuf<KnnParams> d_knn;
  DevBuf<int> d_active;
  PinBuf<PairParams> h_pairs;
  PinBuf<KnnParams> h_knn;
  StageFence knn_fence;
  PinBuf<int> h_active;
  PinBuf<double> h_T;    // [n][12]  staging of the transforms of the next launch
  DevBuf<double> d_T;    // [n][12]
  PinBuf<double> h_out;  // [n][NRED] written by k_reduce_partials straight into pinned host memory
  DevBuf<unsigned long long> d_stat;
  bool stats = false;
};
static BatchScratch& scratch() {
  static BatchScratch s;
  return s;
}

static bool scratch_stats_enabled() { return scratch().stats; }
static std::vector<Gicp::PairBufs*>& pair_pool() {
  static std::vector<Gicp::PairBufs*> p;
  return p;
}
static int ensure_pair_buffers(Gicp* g) {
  if (!g->pb) {
    if (!pair_pool().empty()) {
      g->pb = pair_pool().back();
      pair_pool().pop_back();
    } else {
      g->pb = new Gicp::PairBufs();
    }
  }
  int npad = g->src->npad;
  MR_CHECK(g->pb->corr.reserve(std::max(npad, 1)));
  MR_CHECK(g->pb->nnpos.reserve(std::max(npad, 1)));
  MR_CHECK(g->pb->sqd.reserve(std::max(npad, 1)));
  MR_CHECK(g->pb->mahal.reserve((size_t)std::max(npad, 1) * 6));
  MR_CHECK(g->pb->partials.reserve((size_t)((g->src->ntiles + 3) / 4 + 1) * NRED));
  return MRSLAM_OK;
}

static PairParams make_pair_params(Gicp* g) {
  PairParams p;
  p.src_sort = g->src->psort.p;
  p.src_cov = g->src->cov.p;
  p.src_ntiles = g->src->ntiles;
  p.tgt_sort = g->tgt->psort.p;
  p.tgt_tb = g->tgt->tb.p;
  p.tgt_sb = g->tgt->sb.p;
  p.tgt_qp = g->tgt->qp.p;
  p.tgt_keys = g->tgt->skeys.p;
  p.tgt_n = g->tgt->n;
  p.tgt_cov = g->tgt->cov.p;
  p.tgt_ntiles = g->tgt->ntiles;
  p.tgt_ntiles_pad = g->tgt->ntiles_pad;
  p.tgt_nsuper = g->tgt->nsuper;
  p.tgt_nsuper_pad = g->tgt->nsuper_pad;
  p.corr = g->pb->corr.p;
  p.nnpos = g->pb->nnpos.p;
  p.sqd = g->pb->sqd.p;
  p.mahal = g->pb->mahal.p;
  p.partials = g->pb->partials.p;
  p.thr2 = g->corr_dist_threshold * g->corr_dist_threshold;
  p.use_seed = g->have_seed ? 1 : 0;
  return p;
}

static void iso_to_T12(const Iso& x, double* T) {
  for (int r = 0; r < 3; r++) {
    T[4 * r] = x.R[3 * r];
    T[4 * r + 1] = x.R[3 * r + 1];
    T[4 * r + 2] = x.R[3 * r + 2];
    T[4 * r + 3] = x.t[r];
  }
}

// covariances for every distinct cloud of the batch that lacks them: one launch per k-capacity
static int ensure_built(Gicp** hs, int n) {
  std::vector<Cloud*> cs;
  for (int i = 0; i < n; i++) {
    cs.push_back(hs[i]->src);
    cs.push_back(hs[i]->tgt);
  }
  return build_clouds(cs);
}

// which: bit 0 = source clouds, bit 1 = target clouds
static int compute_covariances(Gicp** hs, int n, int which = 3) {
  BatchScratch& S = scratch();
  std::vector<Cloud*> todo, all;
  std::vector<int> ks, regs;
  for (int i = 0; i < n; i++)
    for (int side = 0; side < 2; side++) {
      Cloud* c = side == 0 ? hs[i]->src : hs[i]->tgt;
      if (!(which & (1 << side)) || !c) continue;
      all.push_back(c);
      if (c->cov_valid) continue;
      if (std::find(todo.begin(), todo.end(), c) != todo.end()) continue;
      todo.push_back(c);
      ks.push_back(hs[i]->k_corr);
      regs.push_back(hs[i]->regularization);
    }
  MR_CHECK(build_clouds(all));
  if (todo.empty()) return MRSLAM_OK;
  MR_CHECK(S.knn_fence.wait());
  MR_CHECK(S.h_knn.reserve(todo.size()));
  MR_CHECK(S.d_knn.reserve(todo.size()));
  // group by k (normally a single group)
  std::vector<int> order(todo.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ks[a] < ks[b]; });
  size_t pos = 0;
  for (size_t i = 0; i < order.size(); i++) {
    Cloud* c = todo[order[i]];
    MR_CHECK(c->cov.reserve((size_t)std::max(c->npad, 1) * 6));
    MR_CHECK(c->knn_ids.reserve((size_t)std::max(c->npad, 1) * ks[order[i]]));
    MR_CHECK(c->knn_fb.reserve((size_t)c->ntiles + 4));
    KnnParams p;
    p.fb = c->knn_fb.p;
    p.n = c->n; p.ids = c->knn_ids.p;
    p.psort = c->psort.p; p.porig = c->porig.p; p.tb = c->tb.p; p.sb = c->sb.p;
    p.ntiles = c->ntiles; p.ntiles_pad = c->ntiles_pad; p.nsuper = c->nsuper; p.nsuper_pad = c->nsuper_pad;
    p.k = ks[order[i]]; p.reg = regs[order[i]];
    p.cov = c->cov.p; p.knn_idx = nullptr; p.knn_sqd = nullptr;
    p.stat_pairs = S.stats ? S.d_stat.p : nullptr;
    S.h_knn.p[pos++] = p;
  }
  MR_CUDA(upload_params(S.d_knn.p, S.h_knn.p, sizeof(KnnParams) * todo.size(), rt().stream));
  MR_CHECK(S.knn_fence.mark(rt().stream));
  size_t i0 = 0;
  while (i0 < order.size()) {
    size_t i1 = i0;
    int k = ks[order[i0]], maxt = 0;
    while (i1 < order.size() && ks[order[i1]] == k) {
      maxt = std::max(maxt, todo[order[i1]]->ntiles);
      i1++;
    }
    MR_CHECK(dispatch_knn(k, S.d_knn.p + i0, (int)(i1 - i0), maxt, true));
    i0 = i1;
  }
  for (Cloud* c : todo) c->cov_valid = true;
  return MRSLAM_OK;
}

static bool is_converged(const Gicp* g, const Iso& delta) {
  double rmax = 0, tmax = 0;
  for (int i = 0; i < 9; i++) rmax = std::max(rmax, std::fabs(delta.R[i] - ((i % 4 == 0) ? 1.0 : 0.0)) / g->rotation_epsilon);
  for (int i = 0; i < 3; i++) tmax = std::max(tmax, std::fabs(delta.t[i]) / g->transformation_epsilon);
  return std::max(rmax, tmax) < 1;
}

static void delta_from(const double d[6], Iso& delta) {
  so3_exp(d, delta.R);
  delta.t[0] = d[3];
  delta.t[1] = d[4];
  delta.t[2] = d[5];
}

static void unpack_H(const double* out, double H[36], double b[6], double* cost) {
  int o = 0;
  for (int r = 0; r < 6; r++)
    for (int c = r; c < 6; c++) {
      H[6 * r + c] = out[o];
      H[6 * c + r] = out[o];
      o++;
    }
  for (int r = 0; r < 6; r++) b[r] = out[21 + r];
  *cost = out[27];
}

// launch linearize (nn1 + reduce) for the listed problems at the transforms in S.h_T; blocks until results are in S.h_out
static int run_linearize(Gicp** hs, const std::vector<int>& ids) {
  BatchScratch& S = scratch();
  cudaStream_t st = rt().stream;
  int na = (int)ids.size();
  if (na == 0) return MRSLAM_OK;
  int maxt = 0;
  for (int a = 0; a < na; a++) {
    Gicp* g = hs[ids[a]];
    S.h_pairs.p[ids[a]] = make_pair_params(g);
    S.h_active.p[a] = ids[a];
    maxt = std::max(maxt, g->src->ntiles);
    for (int c = 0; c < NRED; c++) S.h_out.p[(size_t)ids[a] * NRED + c] = 0.0;
  }
  // whole descriptor table + active list (small): one async copy each
  int nmax = *std::max_element(ids.begin(), ids.end()) + 1;
  MR_CUDA(upload_params(S.d_pairs.p, S.h_pairs.p, sizeof(PairParams) * nmax, st));
  MR_CUDA(upload_params(S.d_active.p, S.h_active.p, sizeof(int) * na, st));
  MR_CUDA(upload_params(S.d_T.p, S.h_T.p, sizeof(double) * 12 * nmax, st));
  if (maxt > 0) {
    dim3 grid((maxt + 3) / 4, na);
    {
      ProfScope ps(PROF_NN1_SEARCH);
      static const int nn1_wpb = getenv("MRSLAM_NN1_WPB") ? std::max(1, std::min(4, atoi(getenv("MRSLAM_NN1_WPB")))) : 4;
      const dim3 g1((maxt + nn1_wpb - 1) / nn1_wpb, na);
      k_nn1_search<MODE_LINEARIZE><<<g1, 32 * nn1_wpb, 0, st>>>(S.d_pairs.p, S.d_active.p, S.d_T.p, 0.0, S.stats ? S.d_stat.p : nullptr);
    }
    {
      ProfScope ps(PROF_LINEARIZE);
      k_linearize<<<grid, 128, 0, st>>>(S.d_pairs.p, S.d_active.p, S.d_T.p);
    }
    k_reduce_partials<<<na, 32 * NRED, 0, st>>>(S.d_pairs.p, S.d_active.p, 0, NRED, S.h_out.p);
    count_launch(3);
  }
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaStreamSynchronize(st));
  for (int a = 0; a < na; a++) {
    Gicp* g = hs[ids[a]];
    g->have_seed = true;
    g->have_corr = true;
    g->n_linearize++;
  }
  return MRSLAM_OK;
}

static int run_error(Gicp** hs, const std::vector<int>& ids) {
  BatchScratch& S = scratch();
  cudaStream_t st = rt().stream;
  int na = (int)ids.size();
  if (na == 0) return MRSLAM_OK;
  int maxt = 0;
  for (int a = 0; a < na; a++) {
    S.h_active.p[a] = ids[a];
    maxt = std::max(maxt, hs[ids[a]]->src->ntiles);
    S.h_out.p[(size_t)ids[a] * NRED + 27] = 0.0;
    hs[ids[a]]->n_error++;
  }
  MR_CUDA(upload_params(S.d_active.p, S.h_active.p, sizeof(int) * na, st));
  {
    int nmax = *std::max_element(ids.begin(), ids.end()) + 1;
    MR_CUDA(upload_params(S.d_T.p, S.h_T.p, sizeof(double) * 12 * nmax, st));
  }
  if (maxt > 0) {
    dim3 grid((maxt + 3) / 4, na);
    {
      ProfScope ps(PROF_ERROR);
      k_error<<<grid, 128, 0, st>>>(S.d_pairs.p, S.d_active.p, S.d_T.p);
    }
    k_reduce_partials<<<na, 32, 0, st>>>(S.d_pairs.p, S.d_active.p, 27, 1, S.h_out.p);
    count_launch(2);
  }
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaStreamSynchronize(st));
  return MRSLAM_OK;
}

static int reserve_scratch(int n) {
  BatchScratch& S = scratch();
  MR_CHECK(S.d_pairs.reserve(n));
  MR_CHECK(S.d_active.reserve(n));
  MR_CHECK(S.h_pairs.reserve(n));
  MR_CHECK(S.h_active.reserve(n));
  MR_CHECK(S.h_T.reserve((size_t)n * 12));
  MR_CHECK(S.d_T.reserve((size_t)n * 12));
  MR_CHECK(S.h_out.reserve((size_t)n * NRED));
  MR_CHECK(S.d_stat.reserve(8));
  return MRSLAM_OK;
}

// LsqRegistration::computeTransformation for a batch, every LM phase advanced in lock step.
// lsq_registration_impl.hpp:53-79 (outer loop), :125-171 (step_lm), :102-122 (step_gn).
static int align_batch(Gicp** hs, int n, const float* guesses16) {
  BatchScratch& S = scratch();
  MR_CHECK(reserve_scratch(n));
  for (int i = 0; i < n; i++) {
    Gicp* g = hs[i];
    if (!g->src || !g->tgt) {
      set_error("align: input source/target not set");
      return MRSLAM_ERR_STATE;
    }
    MR_CHECK(ensure_pair_buffers(g));
  }
  // FastGICP::computeTransformation (fast_gicp_impl.hpp:103-112): covariances on demand
  MR_CHECK(compute_covariances(hs, n));

  std::vector<int> active;
  for (int i = 0; i < n; i++) {
    Gicp* g = hs[i];
    if (guesses16) {
      const float* G = guesses16 + (size_t)i * 16;
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) g->x0.R[3 * r + c] = (double)G[4 * r + c];
        g->x0.t[r] = (double)G[4 * r + 3];
      }
    } else {
      g->x0 = iso_identity();
    }
    g->lm_lambda = -1.0;
    g->converged = false;
    g->lm_failed = false;
    g->nr_iterations = 0;
    g->n_linearize = g->n_error = 0;
    g->have_seed = false;
    g->have_corr = false;
    if (g->max_iterations > 0) active.push_back(i);
  }

  int iter = 0;
  while (!active.empty()) {
    for (int id : active) {
      hs[id]->nr_iterations = iter;
      iso_to_T12(hs[id]->x0, S.h_T.p + (size_t)id * 12);
    }
    MR_CHECK(run_linearize(hs, active));
    std::vector<int> pending;  // problems still inside the LM inner loop
    std::vector<int> step_ok(n, 0);
    for (int id : active) {
      Gicp* g = hs[id];
      unpack_H(S.h_out.p + (size_t)id * NRED, g->H, g->b, &g->y0);
      if (g->optimizer == MRSLAM_OPT_GN) {
        double nb[6];
        for (int i = 0; i < 6; i++) nb[i] = -g->b[i];
        if (ldlt_solve6(g->H, nb, g->d)) {
          delta_from(g->d, g->delta);
          g->x0 = iso_mul(g->delta, g->x0);
          memcpy(g->final_H, g->H, sizeof(g->H));
          step_ok[id] = 1;
        }
        continue;
      }
      if (g->lm_lambda < 0.0) {
        double mx = 0;
        for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(g->H[7 * i]));
        g->lm_lambda = g->lm_init_lambda_factor * mx;
      }
      g->nu = 2.0;
      g->lm_try = 0;
      pending.push_back(id);
    }
    while (!pending.empty()) {
      std::vector<int> evals;
      for (int id : pending) {
        Gicp* g = hs[id];
        double A[36], nb[6];
        memcpy(A, g->H, sizeof(A));
        for (int i = 0; i < 6; i++) {
          A[7 * i] += g->lm_lambda;
          nb[i] = -g->b[i];
        }
        if (!ldlt_solve6(A, nb, g->d)) {
          g->lm_try = g->lm_max_iterations;  // cannot recover: fail this step
          continue;
        }
        delta_from(g->d, g->delta);
        g->xi = iso_mul(g->delta, g->x0);
        iso_to_T12(g->xi, S.h_T.p + (size_t)id * 12);
        evals.push_back(id);
      }
      MR_CHECK(run_error(hs, evals));
      std::vector<int> next;
      for (int id : pending) {
        Gicp* g = hs[id];
        if (g->lm_try >= g->lm_max_iterations) continue;  // failed solve
        double yi = S.h_out.p[(size_t)id * NRED + 27];
        double denom = 0;
        for (int i = 0; i < 6; i++) denom += g->d[i] * (g->lm_lambda * g->d[i] - g->b[i]);
        double rho = (g->y0 - yi) / denom;
        if (g->debug_print)
          fprintf(stdout, "lm[%d] try %d y0 %.9g yi %.9g rho %.6g lambda %.6g\n", id, g->lm_try, g->y0, yi, rho, g->lm_lambda);
        if (rho < 0) {
          if (is_converged(g, g->delta)) {
            step_ok[id] = 1;
            continue;
          }
          g->lm_lambda = g->nu * g->lm_lambda;
          g->nu = 2 * g->nu;
          g->lm_try++;
          if (g->lm_try < g->lm_max_iterations) next.push_back(id);
          continue;
        }
        g->x0 = g->xi;
        g->lm_lambda = g->lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
        memcpy(g->final_H, g->H, sizeof(g->H));
        step_ok[id] = 1;
      }
      pending.swap(next);
    }
    std::vector<int> still;
    for (int id : active) {
      Gicp* g = hs[id];
      if (!step_ok[id]) {
        fprintf(stderr, "lm not converged!!\n");
        g->lm_failed = true;
        continue;
      }
      g->converged = is_converged(g, g->delta);
      if (!g->converged && iter + 1 < g->max_iterations) still.push_back(id);
    }
    active.swap(still);
    iter++;
  }
  int status = MRSLAM_OK;
  for (int i = 0; i < n; i++)
    if (hs[i]->lm_failed) status = MRSLAM_ERR_NOTCONV;
  return status;
}

static int fitness_batch(Gicp** hs, int n, double max_range, double* out) {
  BatchScratch& S = scratch();
  cudaStream_t st = rt().stream;
  MR_CHECK(reserve_scratch(n));
  int maxt = 0;
  for (int i = 0; i < n; i++) {
    Gicp* g = hs[i];
    if (!g->src || !g->tgt) {
      set_error("getFitnessScore: input source/target not set");
      return MRSLAM_ERR_STATE;
    }
    MR_CHECK(ensure_pair_buffers(g));
    {
      Gicp* one[1] = {g};
      MR_CHECK(ensure_built(one, 1));
    }
    S.h_pairs.p[i] = make_pair_params(g);
    S.h_active.p[i] = i;
    // final_transformation_ is the float cast of x0; the kernel casts the doubles below to float itself
    Iso xf = g->x0;
    for (int k = 0; k < 9; k++) xf.R[k] = (double)(float)xf.R[k];
    for (int k = 0; k < 3; k++) xf.t[k] = (double)(float)xf.t[k];
    iso_to_T12(xf, S.h_T.p + (size_t)i * 12);
    maxt = std::max(maxt, g->src->ntiles);
    S.h_out.p[(size_t)i * NRED] = 0;
    S.h_out.p[(size_t)i * NRED + 1] = 0;
  }
  MR_CUDA(upload_params(S.d_pairs.p, S.h_pairs.p, sizeof(PairParams) * n, st));
  MR_CUDA(upload_params(S.d_active.p, S.h_active.p, sizeof(int) * n, st));
  MR_CUDA(upload_params(S.d_T.p, S.h_T.p, sizeof(double) * 12 * n, st));
  if (maxt > 0) {
    dim3 grid((maxt + 3) / 4, n);
    k_nn1_search<MODE_FITNESS><<<grid, 128, 0, st>>>(S.d_pairs.p, S.d_active.p, S.d_T.p, max_range, S.stats ? S.d_stat.p : nullptr);
    k_reduce_partials<<<n, 64, 0, st>>>(S.d_pairs.p, S.d_active.p, 0, 2, S.h_out.p);
    count_launch(2);
  }
  MR_CUDA(cudaGetLastError());
  MR_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < n; i++) {
    double sum = S.h_out.p[(size_t)i * NRED], cnt = S.h_out.p[(size_t)i * NRED + 1];
    out[i] = cnt > 0 ? sum / cnt : std::numeric_limits<double>::max();
  }
  return MRSLAM_OK;
}

}  // namespace mrslam

// =============================================================================================
// C-ABI
// =============================================================================================
using namespace mrslam;

#define LOCK() std::lock_guard<std::mutex> _lk(rt().mu)
#define REQUIRE(h)                          \
  do {                                      \
    if (!(h)) {                             \
      set_error("%s: null handle", __func__); \
      return MRSLAM_ERR_ARG;                \
    }                                       \
  } while (0)

extern "C" {

mrslam_cloud_t* mrslam_cloud_create_host(const float* xyz, int n, int stride_floats) {
  if (ensure_init() != MRSLAM_OK) return nullptr;
  LOCK();
  Cloud* c = acquire_cloud();
  if (c->set_from_host(xyz, n, stride_floats) != MRSLAM_OK) {
    c->release();
    delete c;
    return nullptr;
  }
  return reinterpret_cast<mrslam_cloud_t*>(c);
}
mrslam_cloud_t* mrslam_cloud_create_device(const float* d_xyz, int n, int stride_floats) {
  if (ensure_init() != MRSLAM_OK) return nullptr;
  LOCK();
  if (n < 0 || stride_floats < 3) {
    set_error("cloud_create_device: bad arguments");
    return nullptr;
  }
  Cloud* c = acquire_cloud();
  if (c->set_from_device_raw(d_xyz, n, stride_floats) != MRSLAM_OK) {
    c->release();
    delete c;
    return nullptr;
  }
  return reinterpret_cast<mrslam_cloud_t*>(c);
}
int mrslam_cloud_create_device_batch(const float* const* d_xyz, const int* n, int stride_floats, int count, mrslam_cloud_t** out) {
  if (count < 0 || (count > 0 && (!d_xyz || !n || !out)) || stride_floats < 3) {
    set_error("cloud_create_device_batch: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(ensure_init());
  LOCK();
  if (count == 0) return MRSLAM_OK;
  static PinBuf<PackDesc> h_desc;
  static DevBuf<PackDesc> d_desc;
  static StageFence fence;
  MR_CHECK(fence.wait());
  MR_CHECK(h_desc.reserve(count));
  MR_CHECK(d_desc.reserve(count));
  int max_n = 0;
  for (int i = 0; i < count; i++) {
    out[i] = nullptr;
    if (n[i] < 0 || (n[i] > 0 && !d_xyz[i])) {
      set_error("cloud_create_device_batch: bad cloud %d", i);
      for (int j = 0; j < i; j++) cloud_unref(reinterpret_cast<Cloud*>(out[j])), out[j] = nullptr;
      return MRSLAM_ERR_ARG;
    }
    Cloud* c = acquire_cloud();
    int st = c->set_sizes(n[i]);
    if (st != MRSLAM_OK) {
      c->release();
      delete c;
      for (int j = 0; j < i; j++) cloud_unref(reinterpret_cast<Cloud*>(out[j])), out[j] = nullptr;
      return st;
    }
    out[i] = reinterpret_cast<mrslam_cloud_t*>(c);
    PackDesc d;
    d.raw = d_xyz[i]; d.porig = c->porig.p; d.n = n[i]; d.stride = stride_floats;
    h_desc.p[i] = d;
    max_n = std::max(max_n, n[i]);
  }
  if (max_n > 0) {
    ProfScope ps(PROF_CLOUD_BUILD);
    MR_CUDA(upload_params(d_desc.p, h_desc.p, sizeof(PackDesc) * count, rt().stream));
    MR_CHECK(fence.mark(rt().stream));
    int gx = std::max(1, std::min((max_n + 255) / 256, 128));
    k_pack_points_batch<<<dim3(gx, count), 256, 0, rt().stream>>>(d_desc.p);
    count_launch(1);
    MR_CUDA(cudaGetLastError());
  }
  return MRSLAM_OK;
}
void mrslam_cloud_destroy(mrslam_cloud_t* h) {
  if (!h) return;
  LOCK();
  cloud_unref(reinterpret_cast<Cloud*>(h));
}
int mrslam_cloud_size(const mrslam_cloud_t* h) { return h ? reinterpret_cast<const Cloud*>(h)->n : 0; }

int mrslam_cloud_knn(mrslam_cloud_t* h, int k, int* idx_out, float* sqd_out) {
  REQUIRE(h);
  MR_CHECK(ensure_init());
  LOCK();
  Cloud* c = reinterpret_cast<Cloud*>(h);
  BatchScratch& S = scratch();
  MR_CHECK(build_clouds(std::vector<Cloud*>(1, c)));
  MR_CHECK(S.knn_fence.wait());
  MR_CHECK(S.h_knn.reserve(1));
  MR_CHECK(S.d_knn.reserve(1));
  DevBuf<int> di;
  DevBuf<float> dd;
  size_t tot = (size_t)std::max(c->n, 1) * k;
  MR_CHECK(di.reserve(tot));
  MR_CHECK(dd.reserve(tot));
  if (k < 1 || k > 32) {
    set_error("k must be in 1..32");
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(c->knn_ids.reserve((size_t)std::max(c->npad, 1) * k));
  MR_CHECK(c->knn_fb.reserve((size_t)c->ntiles + 4));
  KnnParams p;
  p.fb = c->knn_fb.p;
  p.n = c->n; p.ids = c->knn_ids.p;
  p.psort = c->psort.p; p.porig = c->porig.p; p.tb = c->tb.p; p.sb = c->sb.p;
  p.ntiles = c->ntiles; p.ntiles_pad = c->ntiles_pad; p.nsuper = c->nsuper; p.nsuper_pad = c->nsuper_pad; p.k = k; p.reg = 0;
  p.cov = nullptr; p.knn_idx = di.p; p.knn_sqd = dd.p; p.stat_pairs = nullptr;
  S.h_knn.p[0] = p;
  MR_CUDA(upload_params(S.d_knn.p, S.h_knn.p, sizeof(KnnParams), rt().stream));
  MR_CHECK(dispatch_knn(k, S.d_knn.p, 1, c->ntiles, false));
  if (c->n > 0) {
    MR_CUDA(cudaMemcpyAsync(idx_out, di.p, sizeof(int) * (size_t)c->n * k, cudaMemcpyDeviceToHost, rt().stream));
    if (sqd_out) MR_CUDA(cudaMemcpyAsync(sqd_out, dd.p, sizeof(float) * (size_t)c->n * k, cudaMemcpyDeviceToHost, rt().stream));
  }
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  di.release();
  dd.release();
  return MRSLAM_OK;
}

// build_neighbors_NN(pointcloud, k) of the RING++ generator (LoopDetection/src/RING_ros/util.py:163-170): kNN lists (self first),
// sample covariances, eigen-entropy and the 3-D / 2-D eigenvalues that voxelfeat.GPUFeatureExtractor consumes.  Host arrays.
int mrslam_ringpp_build_neighbors(const float* xyz, int n, int stride_floats, int k, int* indices_out, float* covs_out9, float* entropy_out,
                                  float* eigens_out5) {
  MR_CHECK(ensure_init());
  if (n < 0 || stride_floats < 3 || (n > 0 && (!xyz || !indices_out || !eigens_out5))) {
    set_error("ringpp_build_neighbors: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  if (k < 2 || k > 32) {
    set_error("ringpp_build_neighbors: k must be in 2..32 (got %d)", k);
    return MRSLAM_ERR_ARG;
  }
  if (n == 0) return MRSLAM_OK;
  LOCK();
  Cloud* c = acquire_cloud();
  int st = c->set_from_host(xyz, n, stride_floats);
  if (st == MRSLAM_OK) st = build_clouds(std::vector<Cloud*>(1, c));
  BatchScratch& S = scratch();
  DevBuf<int> di;
  DevBuf<float> dout;
  auto body = [&]() -> int {
    MR_CHECK(st);
    MR_CHECK(S.knn_fence.wait());
    MR_CHECK(S.h_knn.reserve(1));
    MR_CHECK(S.d_knn.reserve(1));
    MR_CHECK(di.reserve((size_t)n * k));
    MR_CHECK(dout.reserve((size_t)n * 15));
    MR_CHECK(c->knn_ids.reserve((size_t)std::max(c->npad, 1) * k));
    MR_CHECK(c->knn_fb.reserve((size_t)c->ntiles + 4));
    KnnParams p;
    p.fb = c->knn_fb.p;
    p.n = c->n; p.ids = c->knn_ids.p;
    p.psort = c->psort.p; p.porig = c->porig.p; p.tb = c->tb.p; p.sb = c->sb.p;
    p.ntiles = c->ntiles; p.ntiles_pad = c->ntiles_pad; p.nsuper = c->nsuper; p.nsuper_pad = c->nsuper_pad; p.k = k; p.reg = 0;
    p.cov = nullptr; p.knn_idx = di.p; p.knn_sqd = nullptr; p.stat_pairs = nullptr;
    S.h_knn.p[0] = p;
    MR_CUDA(upload_params(S.d_knn.p, S.h_knn.p, sizeof(KnnParams), rt().stream));
    MR_CHECK(S.knn_fence.mark(rt().stream));
    MR_CHECK(dispatch_knn(k, S.d_knn.p, 1, c->ntiles, false));
    float* cov = dout.p;
    float* ent = dout.p + (size_t)n * 9;
    float* eig = dout.p + (size_t)n * 10;
    k_ringpp_eigens<<<(n + 127) / 128, 128, 0, rt().stream>>>(c->porig.p, n, k, di.p, cov, ent, eig);
    count_launch(1);
    MR_CUDA(cudaGetLastError());
    MR_CUDA(cudaMemcpyAsync(indices_out, di.p, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToHost, rt().stream));
    if (covs_out9) MR_CUDA(cudaMemcpyAsync(covs_out9, cov, sizeof(float) * (size_t)n * 9, cudaMemcpyDeviceToHost, rt().stream));
    if (entropy_out) MR_CUDA(cudaMemcpyAsync(entropy_out, ent, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, rt().stream));
    MR_CUDA(cudaMemcpyAsync(eigens_out5, eig, sizeof(float) * (size_t)n * 5, cudaMemcpyDeviceToHost, rt().stream));
    MR_CUDA(cudaStreamSynchronize(rt().stream));
    return MRSLAM_OK;
  };
  st = body();
  di.release();
  dout.release();
  cloud_unref(c);
  return st;
}

mrslam_gicp_t* mrslam_gicp_create(void) {
  if (ensure_init() != MRSLAM_OK) return nullptr;
  return reinterpret_cast<mrslam_gicp_t*>(new Gicp());
}
void mrslam_gicp_destroy(mrslam_gicp_t* h) {
  if (!h) return;
  LOCK();
  Gicp* g = reinterpret_cast<Gicp*>(h);
  cloud_unref(g->src);
  cloud_unref(g->tgt);
  if (g->pb) {
    if (pair_pool().size() < 1024) {
      pair_pool().push_back(g->pb);
    } else {
      g->pb->corr.release(); g->pb->nnpos.release(); g->pb->sqd.release(); g->pb->mahal.release(); g->pb->partials.release();
      delete g->pb;
    }
  }
  delete g;
}

#define G(h) (reinterpret_cast<Gicp*>(h))
#define CG(h) (reinterpret_cast<const Gicp*>(h))

int mrslam_gicp_set_num_threads(mrslam_gicp_t* h, int n) { REQUIRE(h); G(h)->num_threads = n; return MRSLAM_OK; }
int mrslam_gicp_set_correspondence_randomness(mrslam_gicp_t* h, int k) {
  REQUIRE(h);
  if (k < 1 || k > 32) {
    set_error("setCorrespondenceRandomness: k must be in 1..32 (got %d)", k);
    return MRSLAM_ERR_ARG;
  }
  G(h)->k_corr = k;
  return MRSLAM_OK;
}
int mrslam_gicp_set_regularization_method(mrslam_gicp_t* h, int m) {
  REQUIRE(h);
  if (m < 0 || m > 4) {
    set_error("setRegularizationMethod: unknown method %d", m);
    return MRSLAM_ERR_ARG;
  }
  G(h)->regularization = m;
  return MRSLAM_OK;
}
int mrslam_gicp_set_max_correspondence_distance(mrslam_gicp_t* h, double d) { REQUIRE(h); G(h)->corr_dist_threshold = d; return MRSLAM_OK; }
int mrslam_gicp_set_maximum_iterations(mrslam_gicp_t* h, int n) { REQUIRE(h); G(h)->max_iterations = n; return MRSLAM_OK; }
int mrslam_gicp_set_transformation_epsilon(mrslam_gicp_t* h, double e) { REQUIRE(h); G(h)->transformation_epsilon = e; return MRSLAM_OK; }
int mrslam_gicp_set_rotation_epsilon(mrslam_gicp_t* h, double e) { REQUIRE(h); G(h)->rotation_epsilon = e; return MRSLAM_OK; }
int mrslam_gicp_set_initial_lambda_factor(mrslam_gicp_t* h, double f) { REQUIRE(h); G(h)->lm_init_lambda_factor = f; return MRSLAM_OK; }
int mrslam_gicp_set_lsq_optimizer(mrslam_gicp_t* h, int t) {
  REQUIRE(h);
  if (t != MRSLAM_OPT_LM && t != MRSLAM_OPT_GN) {
    set_error("unknown optimizer %d", t);
    return MRSLAM_ERR_ARG;
  }
  G(h)->optimizer = t;
  return MRSLAM_OK;
}
int mrslam_gicp_set_debug_print(mrslam_gicp_t* h, int on) { REQUIRE(h); G(h)->debug_print = on != 0; return MRSLAM_OK; }

static int set_cloud(Gicp* g, bool source, Cloud* c) {
  Cloud*& slot = source ? g->src : g->tgt;
  cloud_unref(slot);
  slot = c;
  g->have_seed = false;
  g->have_corr = false;
  return MRSLAM_OK;
}
static int set_input_host(mrslam_gicp_t* h, bool source, const float* xyz, int n, int stride) {
  REQUIRE(h);
  MR_CHECK(ensure_init());
  LOCK();
  Cloud* c = acquire_cloud();
  int s = c->set_from_host(xyz, n, stride);
  if (s != MRSLAM_OK) {
    c->release();
    delete c;
    return s;
  }
  return set_cloud(G(h), source, c);
}
int mrslam_gicp_set_input_source(mrslam_gicp_t* h, const float* xyz, int n, int stride) { return set_input_host(h, true, xyz, n, stride); }
int mrslam_gicp_set_input_target(mrslam_gicp_t* h, const float* xyz, int n, int stride) { return set_input_host(h, false, xyz, n, stride); }
int mrslam_gicp_set_input_source_cloud(mrslam_gicp_t* h, mrslam_cloud_t* c) {
  REQUIRE(h);
  REQUIRE(c);
  LOCK();
  reinterpret_cast<Cloud*>(c)->refs++;
  return set_cloud(G(h), true, reinterpret_cast<Cloud*>(c));
}
int mrslam_gicp_set_input_target_cloud(mrslam_gicp_t* h, mrslam_cloud_t* c) {
  REQUIRE(h);
  REQUIRE(c);
  LOCK();
  reinterpret_cast<Cloud*>(c)->refs++;
  return set_cloud(G(h), false, reinterpret_cast<Cloud*>(c));
}
int mrslam_gicp_swap_source_and_target(mrslam_gicp_t* h) {
  REQUIRE(h);
  LOCK();
  std::swap(G(h)->src, G(h)->tgt);  // covariances travel with the clouds (fast_gicp_impl.hpp:43-51)
  G(h)->have_seed = false;
  G(h)->have_corr = false;
  return MRSLAM_OK;
}
int mrslam_gicp_clear_source(mrslam_gicp_t* h) { REQUIRE(h); LOCK(); return set_cloud(G(h), true, nullptr); }
int mrslam_gicp_clear_target(mrslam_gicp_t* h) { REQUIRE(h); LOCK(); return set_cloud(G(h), false, nullptr); }

static int set_covs(mrslam_gicp_t* h, bool source, const double* cov16, int n) {
  REQUIRE(h);
  LOCK();
  Cloud* c = source ? G(h)->src : G(h)->tgt;
  if (!c || n != c->n || !cov16) {
    set_error("setCovariances: cloud not set or size mismatch");
    return MRSLAM_ERR_STATE;
  }
  MR_CHECK(build_clouds(std::vector<Cloud*>(1, c)));
  DevBuf<double> tmp;
  MR_CHECK(tmp.reserve((size_t)std::max(n, 1) * 16));
  MR_CHECK(c->cov.reserve((size_t)std::max(c->npad, 1) * 6));
  MR_CUDA(cudaMemcpyAsync(tmp.p, cov16, sizeof(double) * 16 * (size_t)n, cudaMemcpyHostToDevice, rt().stream));
  if (c->npad > 0) {
    k_cov_from_user<<<(c->npad + 255) / 256, 256, 0, rt().stream>>>(tmp.p, c->psort.p, c->npad, c->cov.p);
    count_launch(1);
  }
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  tmp.release();
  c->cov_valid = true;
  return MRSLAM_OK;
}
static int get_covs(mrslam_gicp_t* h, bool source, double* cov16, int n) {
  REQUIRE(h);
  LOCK();
  Gicp* g = G(h);
  Cloud* c = source ? g->src : g->tgt;
  if (!c || n != c->n || !cov16) {
    set_error("getCovariances: cloud not set or size mismatch");
    return MRSLAM_ERR_STATE;
  }
  MR_CHECK(build_clouds(std::vector<Cloud*>(1, c)));
  if (!c->cov_valid) {
    // compute for this cloud only
    Gicp tmpg = Gicp();
    tmpg.k_corr = g->k_corr;
    tmpg.regularization = g->regularization;
    tmpg.src = c;
    tmpg.tgt = c;
    Gicp* arr[1] = {&tmpg};
    int s = compute_covariances(arr, 1);
    tmpg.src = tmpg.tgt = nullptr;
    MR_CHECK(s);
  }
  DevBuf<double> tmp;
  MR_CHECK(tmp.reserve((size_t)std::max(n, 1) * 16));
  if (c->npad > 0) {
    k_cov_to_user<<<(c->npad + 255) / 256, 256, 0, rt().stream>>>(c->cov.p, c->psort.p, c->npad, tmp.p);
    count_launch(1);
  }
  MR_CUDA(cudaMemcpyAsync(cov16, tmp.p, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost, rt().stream));
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  tmp.release();
  return MRSLAM_OK;
}
int mrslam_gicp_set_source_covariances(mrslam_gicp_t* h, const double* c, int n) { return set_covs(h, true, c, n); }
int mrslam_gicp_set_target_covariances(mrslam_gicp_t* h, const double* c, int n) { return set_covs(h, false, c, n); }
int mrslam_gicp_get_source_covariances(mrslam_gicp_t* h, double* c, int n) { return get_covs(h, true, c, n); }
int mrslam_gicp_get_target_covariances(mrslam_gicp_t* h, double* c, int n) { return get_covs(h, false, c, n); }

int mrslam_gicp_align_batch(mrslam_gicp_t** handles, int n, const float* guesses16) {
  if (n <= 0) return MRSLAM_OK;
  if (!handles) {
    set_error("align_batch: null handle array");
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(ensure_init());
  LOCK();
  for (int i = 0; i < n; i++)
    if (!handles[i]) {
      set_error("align_batch: null handle %d", i);
      return MRSLAM_ERR_ARG;
    }
  return align_batch(reinterpret_cast<Gicp**>(handles), n, guesses16);
}
int mrslam_gicp_prepare_batch(mrslam_gicp_t** handles, int n, int which) {
  if (n < 0 || (n > 0 && !handles) || (which & ~3) || which == 0) {
    set_error("prepare_batch: bad arguments");
    return MRSLAM_ERR_ARG;
  }
  MR_CHECK(ensure_init());
  LOCK();
  std::vector<Gicp*> hs;
  for (int i = 0; i < n; i++) {
    REQUIRE(handles[i]);
    Gicp* g = G(handles[i]);
    if (((which & 1) && !g->src) || ((which & 2) && !g->tgt)) {
      set_error("prepare_batch: registration %d has no %s cloud", i, ((which & 1) && !g->src) ? "source" : "target");
      return MRSLAM_ERR_STATE;
    }
    hs.push_back(g);
  }
  if (hs.empty()) return MRSLAM_OK;
  MR_CHECK(reserve_scratch((int)hs.size()));
  return compute_covariances(hs.data(), (int)hs.size(), which);
}

int mrslam_gicp_align(mrslam_gicp_t* h, const float* guess16) { return mrslam_gicp_align_batch(&h, 1, guess16); }

int mrslam_gicp_get_final_transformation_d(const mrslam_gicp_t* h, double* T) {
  REQUIRE(h);
  const Iso& x = CG(h)->x0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T[4 * r + c] = x.R[3 * r + c];
    T[4 * r + 3] = x.t[r];
  }
  T[12] = T[13] = T[14] = 0;
  T[15] = 1;
  return MRSLAM_OK;
}
int mrslam_gicp_get_final_transformation(const mrslam_gicp_t* h, float* T) {
  REQUIRE(h);
  double Td[16];
  mrslam_gicp_get_final_transformation_d(h, Td);
  for (int i = 0; i < 16; i++) T[i] = (float)Td[i];
  return MRSLAM_OK;
}
int mrslam_gicp_has_converged(const mrslam_gicp_t* h) { return h ? (CG(h)->converged ? 1 : 0) : 0; }
int mrslam_gicp_get_num_iterations(const mrslam_gicp_t* h) { return h ? CG(h)->nr_iterations : 0; }
int mrslam_gicp_get_final_hessian(const mrslam_gicp_t* h, double* H36) {
  REQUIRE(h);
  memcpy(H36, CG(h)->final_H, sizeof(double) * 36);
  return MRSLAM_OK;
}
int mrslam_gicp_get_eval_counts(const mrslam_gicp_t* h, int* nl, int* ne) {
  REQUIRE(h);
  if (nl) *nl = CG(h)->n_linearize;
  if (ne) *ne = CG(h)->n_error;
  return MRSLAM_OK;
}

int mrslam_gicp_get_fitness_score_batch(mrslam_gicp_t** handles, int n, double max_range, double* out) {
  if (n <= 0) return MRSLAM_OK;
  MR_CHECK(ensure_init());
  LOCK();
  return fitness_batch(reinterpret_cast<Gicp**>(handles), n, max_range, out);
}
double mrslam_gicp_get_fitness_score(mrslam_gicp_t* h, double max_range) {
  double out = std::numeric_limits<double>::max();
  if (mrslam_gicp_get_fitness_score_batch(&h, 1, max_range, &out) != MRSLAM_OK) return -1.0;
  return out;
}

int mrslam_gicp_get_aligned(mrslam_gicp_t* h, float* xyz_out, int stride) {
  REQUIRE(h);
  LOCK();
  Gicp* g = G(h);
  if (!g->src || stride < 3 || !xyz_out) {
    set_error("get_aligned: source not set / bad arguments");
    return MRSLAM_ERR_STATE;
  }
  int n = g->src->n;
  if (n == 0) return MRSLAM_OK;
  DevBuf<float> out, dT;
  MR_CHECK(out.reserve((size_t)n * 3));
  MR_CHECK(dT.reserve(12));
  float Tf[12];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Tf[4 * r + c] = (float)g->x0.R[3 * r + c];
    Tf[4 * r + 3] = (float)g->x0.t[r];
  }
  MR_CUDA(cudaMemcpyAsync(dT.p, Tf, sizeof(Tf), cudaMemcpyHostToDevice, rt().stream));
  k_transform_out<<<(n + 255) / 256, 256, 0, rt().stream>>>(g->src->porig.p, n, dT.p, out.p, 3);
  count_launch(1);
  if (stride == 3) {
    MR_CUDA(cudaMemcpyAsync(xyz_out, out.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, rt().stream));
  } else {
    MR_CUDA(cudaMemcpy2DAsync(xyz_out, sizeof(float) * stride, out.p, sizeof(float) * 3, sizeof(float) * 3, n, cudaMemcpyDeviceToHost,
                              rt().stream));
  }
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  out.release();
  dT.release();
  return MRSLAM_OK;
}

static void T16_to_iso(const double* T, Iso& x) {
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) x.R[3 * r + c] = T[4 * r + c];
    x.t[r] = T[4 * r + 3];
  }
}

int mrslam_gicp_evaluate_cost(mrslam_gicp_t* h, const float* T16, double* H36, double* b6, double* cost) {
  REQUIRE(h);
  MR_CHECK(ensure_init());
  LOCK();
  Gicp* g = G(h);
  if (!g->src || !g->tgt) {
    set_error("evaluateCost: inputs not set");
    return MRSLAM_ERR_STATE;
  }
  Gicp* arr[1] = {g};
  MR_CHECK(reserve_scratch(1));
  MR_CHECK(ensure_pair_buffers(g));
  MR_CHECK(compute_covariances(arr, 1));
  Iso x;
  double Td[16];
  for (int i = 0; i < 16; i++) Td[i] = (double)T16[i];
  T16_to_iso(Td, x);
  iso_to_T12(x, scratch().h_T.p);
  std::vector<int> ids(1, 0);
  g->have_seed = false;
  MR_CHECK(run_linearize(arr, ids));
  double H[36], b[6], c;
  unpack_H(scratch().h_out.p, H, b, &c);
  if (H36) memcpy(H36, H, sizeof(H));
  if (b6) memcpy(b6, b, sizeof(b));
  if (cost) *cost = c;
  return MRSLAM_OK;
}

int mrslam_gicp_compute_error(mrslam_gicp_t* h, const double* T16, double* cost) {
  REQUIRE(h);
  LOCK();
  Gicp* g = G(h);
  if (!g->have_corr) {
    set_error("compute_error: no correspondences yet (call evaluate_cost/align first)");
    return MRSLAM_ERR_STATE;
  }
  Gicp* arr[1] = {g};
  Iso x;
  T16_to_iso(T16, x);
  BatchScratch& S = scratch();
  S.h_pairs.p[0] = make_pair_params(g);
  MR_CUDA(upload_params(S.d_pairs.p, S.h_pairs.p, sizeof(PairParams), rt().stream));
  iso_to_T12(x, S.h_T.p);
  std::vector<int> ids(1, 0);
  MR_CHECK(run_error(arr, ids));
  *cost = S.h_out.p[27];
  return MRSLAM_OK;
}

int mrslam_gicp_debug_last_correspondences(mrslam_gicp_t* h, int* corr, float* sq_dist, double* mahal9) {
  REQUIRE(h);
  LOCK();
  Gicp* g = G(h);
  if (!g->have_corr) {
    set_error("no linearisation has been run yet");
    return MRSLAM_ERR_STATE;
  }
  int n = g->src->n, npad = g->src->npad;
  if (n == 0) return MRSLAM_OK;
  DevBuf<int> dc;
  DevBuf<float> dd;
  DevBuf<double> dm;
  MR_CHECK(dc.reserve(n));
  MR_CHECK(dd.reserve(n));
  MR_CHECK(dm.reserve((size_t)n * 9));
  k_debug_corr<<<(npad + 255) / 256, 256, 0, rt().stream>>>(g->src->psort.p, g->tgt->psort.p, npad, g->pb->corr.p, g->pb->sqd.p, g->pb->mahal.p,
                                                            corr ? dc.p : nullptr, sq_dist ? dd.p : nullptr, mahal9 ? dm.p : nullptr);
  count_launch(1);
  if (corr) MR_CUDA(cudaMemcpyAsync(corr, dc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, rt().stream));
  if (sq_dist) MR_CUDA(cudaMemcpyAsync(sq_dist, dd.p, sizeof(float) * n, cudaMemcpyDeviceToHost, rt().stream));
  if (mahal9) MR_CUDA(cudaMemcpyAsync(mahal9, dm.p, sizeof(double) * 9 * (size_t)n, cudaMemcpyDeviceToHost, rt().stream));
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  dc.release();
  dd.release();
  dm.release();
  return MRSLAM_OK;
}

// instrumentation used by bench.py's roofline block: number of distance evaluations since the last reset
void mrslam_build_diag_dump(const char* label) { build_diag().dump(label); }
int mrslam_gicp_stats_enable(int on) {
  MR_CHECK(ensure_init());
  LOCK();
  BatchScratch& S = scratch();
  MR_CHECK(S.d_stat.reserve(8));
  MR_CUDA(cudaMemsetAsync(S.d_stat.p, 0, sizeof(unsigned long long) * 8, rt().stream));
  S.stats = on != 0;
  return MRSLAM_OK;
}
// which: 0 = kNN distance evaluations, 1 = NN-1 distance evaluations, 2 = k-min insertion rounds
unsigned long long mrslam_gicp_stats_get(int which) {
  LOCK();
  BatchScratch& S = scratch();
  unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (!S.d_stat.p || which < 0 || which > 7) return 0;
  cudaMemcpyAsync(v, S.d_stat.p, sizeof(v), cudaMemcpyDeviceToHost, rt().stream);
  cudaStreamSynchronize(rt().stream);
  return v[which];
}
unsigned long long mrslam_gicp_stats_pairs(void) { return mrslam_gicp_stats_get(0) + mrslam_gicp_stats_get(1); }
}
