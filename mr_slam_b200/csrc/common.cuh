// common.cuh -- shared host/device helpers of libmrslam_b200 (sm_100a only).
// Build policy: every TU is compiled with -fmad=false; fused multiply-adds are written
// explicitly (fmaf / fma) so the arithmetic is the same expression tree as the CPU oracle's.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mrslam_b200.h"

namespace mrslam {

// ---- error plumbing (no exceptions across the C-ABI) -------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int g_status;

#define MR_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      mrslam::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MRSLAM_ERR_CUDA;                                                                  \
    }                                                                                          \
  } while (0)

#define MR_CUDA_NULL(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      mrslam::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return nullptr;                                                                          \
    }                                                                                          \
  } while (0)

#define MR_CHECK(call)            \
  do {                            \
    int _s = (call);              \
    if (_s != MRSLAM_OK) return _s; \
  } while (0)

// ---- runtime singleton: device, stream, memory pool, launch counter ------------------------
struct Runtime {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host -> device uploads of input clouds: they overlap kernels already queued on `stream`
  int sm_count = 148;
  std::atomic<unsigned long long> launches{0};
  std::mutex mu;  // serialises host-side drivers that share the stream
};
Runtime& rt();
int ensure_init();  // MRSLAM_OK or error (no device -> MRSLAM_ERR_CUDA; never a CPU fallback)

// ---- optional per-kernel device timing (bench.py's roofline block): CUDA events on the library stream ----
enum ProfId { PROF_KNN_SEARCH = 0, PROF_COV, PROF_NN1_SEARCH, PROF_LINEARIZE, PROF_ERROR, PROF_CLOUD_BUILD, PROF_ELEV_PROCESS,
              PROF_ELEV_FUSE, PROF_COUNT };
bool prof_enabled();
void prof_begin(int id);
void prof_end(int id);
struct ProfScope {
  int id;
  bool on;
  explicit ProfScope(int i) : id(i), on(prof_enabled()) {
    if (on) prof_begin(id);
  }
  ~ProfScope() {
    if (on) prof_end(id);
  }
};

inline void count_launch(int n = 1) { rt().launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// stream-ordered device buffer (cudaMallocAsync pool; grow-only)
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return MRSLAM_OK;
    if (p) MR_CUDA(cudaFreeAsync(p, rt().stream));
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    MR_CUDA(cudaMallocAsync((void**)&p, want * sizeof(T), rt().stream));
    cap = want;
    return MRSLAM_OK;
  }
  void release() {
    if (p) cudaFreeAsync(p, rt().stream);
    p = nullptr;
    cap = 0;
  }
};

// pinned host buffer (grow-only)
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return MRSLAM_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 64;
    MR_CUDA(cudaMallocHost((void**)&p, want * sizeof(T)));
    cap = want;
    return MRSLAM_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// Guards a pinned staging buffer that an async H2D copy reads: wait() before the host rewrites it, mark() after the copy
// has been queued.  (Only matters when two batched calls follow each other without a stream synchronisation in between.)
struct StageFence {
  cudaEvent_t ev = nullptr;
  bool pending = false;
  int wait() {
    if (pending) {
      MR_CUDA(cudaEventSynchronize(ev));
      pending = false;
    }
    return MRSLAM_OK;
  }
  int mark(cudaStream_t st) {
    if (!ev) MR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    MR_CUDA(cudaEventRecord(ev, st));
    pending = true;
    return MRSLAM_OK;
  }
};

// ---- device helpers --------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

// Small parameter blocks (descriptor tables, pair lists, transforms) reach the device through a kernel that reads the pinned --
// hence device-visible -- host copy, NOT through the copy engine: an H2D cudaMemcpyAsync on the library stream queues in the
// DMA engine behind every bulk upload handed to it before (the clouds of this or of the next batch), which stalled the stream
// for the length of those uploads (~2.5 ms per 32-pair step, gpurun_out/r2ac).  The source must stay untouched until the
// launch has run, exactly as with cudaMemcpyAsync from pinned memory.
static __global__ void k_fetch_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src_pinned, size_t nwords) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src_pinned[i];
}
static inline cudaError_t upload_params(void* dst, const void* src_pinned, size_t bytes, cudaStream_t st) {
  const size_t nw = bytes / 4;  // every parameter type is a multiple of four bytes
  if (nw == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)std::min<size_t>((nw + 255) / 256, 32);
  k_fetch_words<<<blocks, 256, 0, st>>>(reinterpret_cast<unsigned*>(dst), reinterpret_cast<const unsigned*>(src_pinned), nw);
  return cudaGetLastError();
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// fixed-order butterfly: bitwise deterministic
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// squared distance, the one convention shared with the oracle: fmaf(dz,dz, fmaf(dy,dy, dx*dx))
__device__ __forceinline__ float dist2(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = qx - px, dy = qy - py, dz = qz - pz;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
// float transform of a point: fmaf(m2,z, fmaf(m1,y, m0*x)) + m3  (row r of a row-major 3x4)
__device__ __forceinline__ float xform_row(const float* m, float x, float y, float z) {
  return fmaf(m[2], z, fmaf(m[1], y, m[0] * x)) + m[3];
}
// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
#endif

}  // namespace mrslam
