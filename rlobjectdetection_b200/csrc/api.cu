// Version / error strings, launch counter and the optional per-kernel CUDA-event profiler of
// the C ABI (include/rlod.h).
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "rlod_common.cuh"

namespace rlod {

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profile{0};
static std::atomic<int> g_profile_only{-1};
static std::mutex g_mu;
struct EvPair {
  cudaEvent_t a, b;
};
static std::vector<EvPair> g_events[RLOD_KERNEL_COUNT];

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// events are recycled: cudaEventCreate can block for milliseconds when the driver grows its pool,
// and the first launches after a synchronise are the ones nobody can hide behind
static std::map<int, std::vector<cudaEvent_t>> g_free_events;  // per device

static bool take_event(cudaEvent_t *e) {
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto &pool = g_free_events[dev];
    if (!pool.empty()) {
      *e = pool.back();
      pool.pop_back();
      return true;
    }
  }
  return cudaEventCreate(e) == cudaSuccess;
}

ProfScope::ProfScope(int kernel_id, cudaStream_t st) : id_(kernel_id), st_(st), on_(false) {
  note_launch(1);
  if (!g_profile.load(std::memory_order_relaxed) || id_ < 0 || id_ >= RLOD_KERNEL_COUNT) return;
  const int only = g_profile_only.load(std::memory_order_relaxed);
  if (only >= 0 && only != id_) return;
  if (!take_event(&a_)) return;
  if (!take_event(&b_)) {
    cudaEventDestroy(a_);
    return;
  }
  on_ = true;
  cudaEventRecord(a_, st_);
}

ProfScope::~ProfScope() {
  if (!on_) return;
  cudaEventRecord(b_, st_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_events[id_].push_back(EvPair{a_, b_});
}

}  // namespace rlod

using namespace rlod;

RLOD_API int rlod_version(void) { return 100; }

RLOD_API const char *rlod_error_string(int code) {
  switch (code) {
    case RLOD_OK:
      return "ok";
    case RLOD_EINVAL:
      return "invalid argument";
    case RLOD_ENOSPC:
      return "workspace too small";
    case RLOD_EUNSUPPORTED:
      return "shape not supported by the sm_100a kernels";
    default:
      return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

RLOD_API long long rlod_launch_count(void) { return g_launches.load(); }

RLOD_API int rlod_profile_enable(int on) { return g_profile.exchange(on ? 1 : 0); }

RLOD_API int rlod_profile_only(int kernel_id) {
  if (kernel_id >= RLOD_KERNEL_COUNT) return RLOD_EINVAL;
  g_profile_only.store(kernel_id < 0 ? -1 : kernel_id);
  return RLOD_OK;
}

RLOD_API int rlod_profile_collect(int kernel_id, double *total_ms, int *launches) {
  if (kernel_id < 0 || kernel_id >= RLOD_KERNEL_COUNT || !total_ms || !launches) return RLOD_EINVAL;
  std::vector<EvPair> ev;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    ev.swap(g_events[kernel_id]);
  }
  double sum = 0.;
  int n = 0, rc = RLOD_OK;
  for (auto &e : ev) {
    float ms = 0.f;
    cudaError_t err = cudaEventSynchronize(e.b);
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e.a, e.b);
    if (err == cudaSuccess) {
      sum += ms;
      ++n;
    } else {
      rc = (int)err;
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaGetDevice(&dev);  // collected on the device the launches ran on (one device per process in practice)
    auto &pool = g_free_events[dev];
    for (auto &e : ev) pool.push_back(e.a), pool.push_back(e.b);
  }
  *total_ms = sum;
  *launches = n;
  return rc;
}
