#include "evc_host.h"

#include <atomic>
#include <cstdio>
#include <cstring>

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
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace evc

extern "C" const char* evc_last_error(void) { return evc::g_err; }
extern "C" long long evc_launch_count(void) { return evc::g_launches.load(); }
extern "C" int evc_version(void) { return 1; }
