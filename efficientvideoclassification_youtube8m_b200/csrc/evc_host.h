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
int num_sms();                                      // of the calling thread's current device
int opt_in_smem(const void* func, int bytes);       // once per (kernel, device): MaxDynamicSharedMemorySize
// elementwise BasicLSTM cell kernels (evc_kernels.cu) used by the split-K recurrence paths
int launch_lstm_cell_fwd(const float* z_part, int S, long long part_stride, const float* bias, const float* c_prev,
                         const void* h_prev, const int* seq_len, int t, int rows, int H, float* c_out, void* h_out,
                         void* gates, cudaStream_t stream, const void* h_prev_lo = nullptr, void* h_out_lo = nullptr,
                         void* gates_lo = nullptr);
int launch_lstm_cell_bwd(const float* dh_part, int S, long long part_stride, const void* gates, const float* c_prev,
                         const float* dh_ext, long long ld_dh_ext, const float* dh_pass_in, long long ld_dh_pass_in,
                         const float* dc_in, long long ld_dc_in, const int* seq_len, int t, int rows, int H,
                         void* dz_out, float* dc_out, float* dh_pass_out, float* dbias, cudaStream_t stream,
                         const void* gates_lo = nullptr, void* dz_lo = nullptr);
}  // namespace evc
