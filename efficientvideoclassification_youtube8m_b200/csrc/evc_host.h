// Host-side helpers shared by the translation units of libevc: error reporting
// (thread-local message behind evc_last_error), launch accounting, device properties.
#pragma once
#include <cuda_runtime.h>

#define EVC_OK 0
#define EVC_ERR_ARG (-1)
#define EVC_ERR_CUDA (-2)
#define EVC_ERR_UNSUPPORTED (-3)

namespace evc {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* what);  // cudaGetLastError() -> EVC code
void count_launch();
int num_sms();
}  // namespace evc
