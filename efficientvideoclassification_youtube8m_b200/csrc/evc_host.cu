#include "evc_host.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

namespace evc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  std::snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return EVC_ERR_CUDA;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, what);
  return EVC_OK;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// Per-device state (a process may drive several devices, one host thread each): the SM count and the
// shared-memory opt-in of every kernel are looked up by the CURRENT device of the calling thread.
constexpr int kMaxDevices = 64;
int num_sms() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) dev = 0;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
int opt_in_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({func, dev})) return EVC_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  done.insert({func, dev});
  return EVC_OK;
}

}  // namespace evc

extern "C" const char* evc_last_error(void) { return evc::g_err; }
extern "C" long long evc_launch_count(void) { return evc::g_launches.load(); }
extern "C" int evc_version(void) { return 1; }
