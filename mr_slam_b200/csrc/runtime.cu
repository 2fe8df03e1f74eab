// runtime.cu -- process-global runtime of libmrslam_b200: device selection, the library stream,
// the stream-ordered memory pool, error strings, launch counter.
#include "common.cuh"

namespace mrslam {

thread_local int g_status = 0;
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

Runtime& rt() {
  static Runtime r;
  return r;
}

// ---- profiling: event pairs per launch group, resolved lazily ----
struct ProfRec {
  int id;
  cudaEvent_t a, b;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_free_events;
static double g_prof_ms[PROF_COUNT];
static long long g_prof_n[PROF_COUNT];
static cudaEvent_t g_open[PROF_COUNT];
bool prof_enabled() { return g_prof; }
static cudaEvent_t get_event() {
  if (!g_free_events.empty()) {
    cudaEvent_t e = g_free_events.back();
    g_free_events.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int id) {
  g_open[id] = get_event();
  cudaEventRecord(g_open[id], rt().stream);
}
void prof_end(int id) {
  cudaEvent_t b = get_event();
  cudaEventRecord(b, rt().stream);
  g_recs.push_back(ProfRec{id, g_open[id], b});
}
static void prof_resolve() {
  cudaStreamSynchronize(rt().stream);
  for (ProfRec& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      g_prof_ms[r.id] += ms;
      g_prof_n[r.id]++;
    }
    g_free_events.push_back(r.a);
    g_free_events.push_back(r.b);
  }
  g_recs.clear();
}

static int init_device(int device) {
  Runtime& R = rt();
  std::lock_guard<std::mutex> lk(R.mu);
  if (R.device >= 0 && (device < 0 || device == R.device)) return MRSLAM_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); libmrslam_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MRSLAM_ERR_CUDA;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= count) {
    set_error("device %d out of range (%d devices)", device, count);
    return MRSLAM_ERR_ARG;
  }
  MR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return MRSLAM_ERR_CUDA;
  }
  if (R.stream) {
    cudaStreamDestroy(R.stream);
    R.stream = nullptr;
  }
  MR_CUDA(cudaStreamCreateWithFlags(&R.stream, cudaStreamNonBlocking));
  if (R.copy_stream) {
    cudaStreamDestroy(R.copy_stream);
    R.copy_stream = nullptr;
  }
  MR_CUDA(cudaStreamCreateWithFlags(&R.copy_stream, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  MR_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thresh = ~0ull;  // keep freed blocks cached: handles are created per loop-closure check
  MR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  R.sm_count = prop.multiProcessorCount;
  R.device = device;
  return MRSLAM_OK;
}

int ensure_init() {
  if (rt().device >= 0) {
    // the calling thread may not have the device current yet
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != rt().device) cudaSetDevice(rt().device);
    return MRSLAM_OK;
  }
  return init_device(-1);
}

}  // namespace mrslam

using namespace mrslam;

extern "C" {
int mrslam_init(int device) { return init_device(device); }
const char* mrslam_last_error(void) { return g_err; }
const char* mrslam_version(void) { return "mrslam_b200 0.1 (sm_100a)"; }
int mrslam_device_synchronize(void) {
  MR_CHECK(ensure_init());
  MR_CUDA(cudaStreamSynchronize(rt().stream));
  return MRSLAM_OK;
}
void* mrslam_stream(void) { return ensure_init() == MRSLAM_OK ? (void*)rt().stream : nullptr; }
unsigned long long mrslam_kernel_launch_count(void) { return rt().launches.load(); }
// per-kernel device time (CUDA events on the launching stream); ids: see ProfId in common.cuh
int mrslam_prof_enable(int on) {
  MR_CHECK(ensure_init());
  std::lock_guard<std::mutex> lk(rt().mu);
  prof_resolve();
  for (int i = 0; i < PROF_COUNT; i++) {
    g_prof_ms[i] = 0;
    g_prof_n[i] = 0;
  }
  g_prof = on != 0;
  if (g_prof) {  // pre-create the event pool so that no cudaEventCreate lands inside a measured region
    while (g_free_events.size() < 2048) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_free_events.push_back(e);
    }
  }
  return MRSLAM_OK;
}
int mrslam_prof_get(int id, double* total_ms, long long* count) {
  if (id < 0 || id >= PROF_COUNT) return MRSLAM_ERR_ARG;
  std::lock_guard<std::mutex> lk(rt().mu);
  prof_resolve();
  if (total_ms) *total_ms = g_prof_ms[id];
  if (count) *count = g_prof_n[id];
  return MRSLAM_OK;
}
}
