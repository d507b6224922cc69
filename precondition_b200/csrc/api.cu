// Library-wide helpers: error string, version, device capability.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace pc {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::mutex g_mu;
static std::atomic<long long> g_launches{0}, g_gemm_launches{0};
static bool g_timing = false;
static double g_gemm_flops = 0.0;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;

void count_launch(int n) { g_launches += n; }
bool gemm_timing_enabled() { return g_timing; }
void gemm_count(int n) { g_gemm_launches += n; }
void gemm_add_flops(double f) {
  std::lock_guard<std::mutex> l(g_mu);
  g_gemm_flops += f;
}
void gemm_timing_record(cudaEvent_t a, cudaEvent_t b) {
  std::lock_guard<std::mutex> l(g_mu);
  g_events.emplace_back(a, b);
}
}  // namespace pc

extern "C" {
void pc_stats_reset(int enable_gemm_timing) {
  std::lock_guard<std::mutex> l(pc::g_mu);
  for (auto& e : pc::g_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  pc::g_events.clear();
  pc::g_launches = 0;
  pc::g_gemm_launches = 0;
  pc::g_gemm_flops = 0.0;
  pc::g_timing = enable_gemm_timing != 0;
}
void pc_stats_get(pc_stats* out) {
  std::lock_guard<std::mutex> l(pc::g_mu);
  out->kernel_launches = pc::g_launches;
  out->gemm_launches = pc::g_gemm_launches;
  out->gemm_flops = pc::g_gemm_flops;
  double ms = 0.0;
  for (auto& e : pc::g_events) {
    float t = 0.f;
    if (cudaEventSynchronize(e.second) == cudaSuccess &&
        cudaEventElapsedTime(&t, e.first, e.second) == cudaSuccess)
      ms += t;
  }
  out->gemm_ms = ms;
}
int pc_version(void) { return 100; }
const char* pc_last_error(void) { return pc::g_error; }
int pc_device_supports_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return 0;
  return major == 10 ? 1 : 0;
}
}
