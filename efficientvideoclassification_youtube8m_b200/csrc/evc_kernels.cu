// Bandwidth-bound kernels of the H-LSTM teacher-student path (sm_100a):
// frame normalise/gather/pack, sequence lengths, sampler index rules, state packing,
// MoE mixture + cross-entropy, loss gradients, representation loss, bias column sums,
// per-variable clip + TF-Adam, exact top-k.  Reference formulas are cited per kernel
// (paths relative to /root/reference/code_student_uniform).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "evc_host.h"

using namespace evc;

namespace {

// All kernels of this file are launched with programmatic stream serialization: each starts with
// griddepcontrol.launch_dependents / griddepcontrol.wait (PDL_PROLOGUE) so that its launch latency
// overlaps the tail of the previous kernel while it still observes all of that kernel's writes.
#define PDL_PROLOGUE()                                                   \
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");        \
  asm volatile("griddepcontrol.wait;" ::: "memory")

template <typename... KArgs, typename... Args>
static cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {  // sh: 32 floats
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  return sh[0];
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float sigm_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) { return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f; }
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
__device__ __forceinline__ void unpack4_bf16(uint2 w, float* v) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
// Split-bf16 ("precise" mode) operand planes: v = hi + lo with hi = bf16(v), lo = bf16(v - hi); the GEMMs then
// form A_hi*B_hi + A_hi*B_lo + A_lo*B_hi in f32 (~16 mantissa bits per operand instead of 8).
__device__ __forceinline__ void store4_split(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, const float* v) {
  const uint2 h = pack4_bf16(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<uint2*>(hi + off) = h;
  if (lo != nullptr) {
    float q[4];
    unpack4_bf16(h, q);
    *reinterpret_cast<uint2*>(lo + off) = pack4_bf16(v[0] - q[0], v[1] - q[1], v[2] - q[2], v[3] - q[3]);
  }
}
__device__ __forceinline__ void load4_split(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long off, float* v) {
  unpack4_bf16(*reinterpret_cast<const uint2*>(hi + off), v);
  if (lo != nullptr) {
    float q[4];
    unpack4_bf16(*reinterpret_cast<const uint2*>(lo + off), q);
    v[0] += q[0]; v[1] += q[1]; v[2] += q[2]; v[3] += q[3];
  }
}
__device__ __forceinline__ void store1_split(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float v) {
  const __nv_bfloat16 h = __float2bfloat16(v);
  hi[off] = h;
  if (lo != nullptr) lo[off] = __float2bfloat16(v - __bfloat162float(h));
}


// --------------------------------------------------------------------------------------
// train.py:256 l2_normalize + train.py:265-272 gather (or model_utils.py:34-58 gather_nd),
// written in the chunk-time-major bf16 layout the LSTM GEMMs read:
//   out[(tt * C + chunk) * B + b][D]  with frame k = chunk*ell + tt  (C chunks of ell frames)
// One warp per selected frame.
__global__ void frames_pack_kernel(const float* __restrict__ src, int B, int T, int D,
                                   const int* __restrict__ frame_idx, int idx_per_batch, int K, int C,
                                   int normalize, __nv_bfloat16* __restrict__ out_bf16,
                                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_lo) {
  PDL_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * K) return;
  const int b = warp / K, k = warp % K;
  const int f = frame_idx ? (idx_per_batch ? frame_idx[b * K + k] : frame_idx[k]) : k;
  const float4* s = reinterpret_cast<const float4*>(src + (static_cast<long long>(b) * T + f) * D);
  const int n4 = D >> 2;
  float scale = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int i = lane; i < n4; i += 32) {
      const float4 v = __ldg(s + i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    ss = warp_sum(ss);
    scale = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  }
  const int ell = K / C;
  const int chunk = k / ell, tt = k % ell;
  __nv_bfloat16* ob = out_bf16 ? out_bf16 + ((static_cast<long long>(tt) * C + chunk) * B + b) * D : nullptr;
  float4* of = out_f32 ? reinterpret_cast<float4*>(out_f32 + (static_cast<long long>(b) * K + k) * D) : nullptr;
  for (int i = lane; i < n4; i += 32) {
    float4 v = __ldg(s + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (ob) {
      const float q[4] = {v.x, v.y, v.z, v.w};
      store4_split(ob, out_lo ? out_lo + (ob - out_bf16) : nullptr, 4 * i, q);
    }
    if (of) of[i] = v;
  }
}

// Same, reading the uint8 features as the YT8M tfrecords store them: readers.py:160-172 decodes the
// bytes, utils.Dequantize (utils.py:9-25, max 2 / min -2) maps q -> q*(4/255) + (4/512 - 2) in
// float32, and resize_axis zero-pads frames >= num_frames.  Keeping the features uint8 up to this
// kernel moves 4x fewer bytes over PCIe / HBM (SURVEY 8f "next #1").
__global__ void frames_pack_u8_kernel(const uint8_t* __restrict__ src, const int* __restrict__ num_frames, int B,
                                      int T, int D, const int* __restrict__ frame_idx, int idx_per_batch, int K,
                                      int C, int normalize, __nv_bfloat16* __restrict__ out_bf16,
                                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_lo) {
  PDL_PROLOGUE();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * K) return;
  const int b = warp / K, k = warp % K;
  const int f = frame_idx ? (idx_per_batch ? frame_idx[b * K + k] : frame_idx[k]) : k;
  const bool padded = f >= num_frames[b];
  const uchar4* s = reinterpret_cast<const uchar4*>(src + (static_cast<long long>(b) * T + f) * D);
  const int n4 = D >> 2;
  const float scalar = 4.0f / 255.0f, bias = 4.0f / 512.0f - 2.0f;
  float scale = 1.f;
  if (normalize && !padded) {
    float ss = 0.f;
    for (int i = lane; i < n4; i += 32) {
      const uchar4 q = __ldg(s + i);
      const float x0 = __fadd_rn(__fmul_rn(q.x, scalar), bias), x1 = __fadd_rn(__fmul_rn(q.y, scalar), bias);
      const float x2 = __fadd_rn(__fmul_rn(q.z, scalar), bias), x3 = __fadd_rn(__fmul_rn(q.w, scalar), bias);
      ss += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
    }
    ss = warp_sum(ss);
    scale = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  }
  const int ell = K / C;
  const int chunk = k / ell, tt = k % ell;
  __nv_bfloat16* ob = out_bf16 ? out_bf16 + ((static_cast<long long>(tt) * C + chunk) * B + b) * D : nullptr;
  float4* of = out_f32 ? reinterpret_cast<float4*>(out_f32 + (static_cast<long long>(b) * K + k) * D) : nullptr;
  for (int i = lane; i < n4; i += 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!padded) {
      const uchar4 q = __ldg(s + i);
      v.x = __fadd_rn(__fmul_rn(q.x, scalar), bias) * scale;
      v.y = __fadd_rn(__fmul_rn(q.y, scalar), bias) * scale;
      v.z = __fadd_rn(__fmul_rn(q.z, scalar), bias) * scale;
      v.w = __fadd_rn(__fmul_rn(q.w, scalar), bias) * scale;
    }
    if (ob) {
      const float q[4] = {v.x, v.y, v.z, v.w};
      store4_split(ob, out_lo ? out_lo + (ob - out_bf16) : nullptr, 4 * i, q);
    }
    if (of) of[i] = v;
  }
}

// train.py:263-264  int64( (n / 300) * int(300/every_n) ) in float64
__global__ void num_frames_student_kernel(const int* __restrict__ nf, int B, int max_frames, int m,
                                          long long* __restrict__ out) {
  PDL_PROLOGUE();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double q = static_cast<double>(nf[b]) / static_cast<double>(max_frames);
  out[b] = static_cast<long long>(q * static_cast<double>(m));  // truncation toward zero
}

// frame_level_models.py:240,256 (teacher) / :309,327 (student)
//   len_l1[c*B+b] = min(ell, max(0, n - ell*c));  len_l2[b] = int32(ceil(float32(n)/ell))
__global__ void lstm_lengths_kernel(const void* __restrict__ nf, int is64, int B, int C, int ell,
                                    int* __restrict__ len_l1, int* __restrict__ len_l2) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i / B, b = i % B;
  const long long n = is64 ? static_cast<const long long*>(nf)[b] : static_cast<const int*>(nf)[b];
  long long v = n - static_cast<long long>(ell) * c;
  v = v < 0 ? 0 : v;
  v = v > ell ? ell : v;
  len_l1[i] = static_cast<int>(v);
  if (c == 0) len_l2[b] = static_cast<int>(ceilf(static_cast<float>(n) / static_cast<float>(ell)));
}

// model_utils.py:49-53  int32(u * float32(n))
__global__ void random_frame_index_kernel(const float* __restrict__ u, const int* __restrict__ nf, int B, int K,
                                          int* __restrict__ idx) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  // (a video with num_frames = 0 yields index 0: the zero-padded first frame, never an address before the row)
  idx[i] = max(static_cast<int>(__fmul_rn(u[i], static_cast<float>(nf[i / K]))), 0);
}
// model_utils.py:23-33  start = int32(u*float32(max(n-K,0)+1)); idx = min(start+k, n-1)
__global__ void random_sequence_index_kernel(const float* __restrict__ u, const int* __restrict__ nf, int B, int K,
                                             int* __restrict__ idx) {
  PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i % K;
  const int n = nf[b];
  const int ms = max(n - K, 0);
  const int start = static_cast<int>(__fmul_rn(u[b], static_cast<float>(ms + 1)));
  idx[i] = max(min(start + k, n - 1), 0);   // n = 0 (empty video): the reference's -1 would address before the row
}

__global__ void sampled_lengths_kernel(const int* __restrict__ nf, int B, int K, long long* __restrict__ out) {
  PDL_PROLOGUE();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = nf[b] > 0 ? K : 0;
}

// tf.random_uniform([..], dtype=float32) of the samplers (model_utils.py:25-27,50-51): Philox4x32-10 counter-based
// generator [TF random_distributions.h PhiloxRandom], 4 floats per counter, converted the way TF's
// Uint32ToFloat does (23 random mantissa bits -> [1,2) - 1 = [0,1)).  The stream is a pure function of
// (seed, offset + element index); TF's own keys/counters derive from graph-level seeds and are not reproduced.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__global__ void random_uniform_kernel(unsigned long long seed, unsigned long long offset, float* __restrict__ out,
                                      long long n) {
  PDL_PROLOGUE();
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // group of 4 outputs
  if (q * 4 >= n) return;
  const unsigned long long ctr = offset + static_cast<unsigned long long>(q);
  uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (q * 4 + j < n) out[q * 4 + j] = __uint_as_float((127u << 23) | (c[j] & 0x7FFFFFu)) - 1.0f;
}

// MultiRNNCell(state_is_tuple=False) state = [c0|h0|c1|h1]: gather the final
// (c, h) of both cells into the 4H-wide row that RNN_L2 / MoE / L_REP consume.
__global__ void state_pack_kernel(const float* __restrict__ c0, const __nv_bfloat16* __restrict__ h0,
                                  const float* __restrict__ c1, const __nv_bfloat16* __restrict__ h1, long long n,
                                  int H, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32,
                                  const __nv_bfloat16* __restrict__ h0_lo, const __nv_bfloat16* __restrict__ h1_lo,
                                  __nv_bfloat16* __restrict__ out_lo) {
  PDL_PROLOGUE();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / H;
  const int u = static_cast<int>(i % H);
  const float vc0 = c0[i], vc1 = c1[i];
  float vh0 = __bfloat162float(h0[i]), vh1 = __bfloat162float(h1[i]);
  if (h0_lo) { vh0 += __bfloat162float(h0_lo[i]); vh1 += __bfloat162float(h1_lo[i]); }
  const long long o = r * 4 * H + u;
  if (out_bf16) {
    store1_split(out_bf16, out_lo, o, vc0);
    store1_split(out_bf16, out_lo, o + H, vh0);
    store1_split(out_bf16, out_lo, o + 2 * H, vc1);
    store1_split(out_bf16, out_lo, o + 3 * H, vh1);
  }
  if (out_f32) {
    out_f32[o] = vc0;
    out_f32[o + H] = vh0;
    out_f32[o + 2 * H] = vc1;
    out_f32[o + 3 * H] = vh1;
  }
}

// BasicLSTMCell forward on split-K partial pre-activations (small-row recurrence steps):
//   z = sum_s z_part[s] + bias ; i,j,f,o ; c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j) ; h' = tanh(c')*sigmoid(o)
// with the dynamic_rnn copy-through for rows past their sequence_length.  Thread = (row, 4 units).
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ z_part, int S, long long part_stride,
                                     const float* __restrict__ bias, const float* __restrict__ c_prev,
                                     const __nv_bfloat16* __restrict__ h_prev, const int* __restrict__ seq_len,
                                     int t, int rows, int H, float* __restrict__ c_out,
                                     __nv_bfloat16* __restrict__ h_out, __nv_bfloat16* __restrict__ gates,
                                     const __nv_bfloat16* __restrict__ h_prev_lo, __nv_bfloat16* __restrict__ h_out_lo,
                                     __nv_bfloat16* __restrict__ gates_lo) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL (see evc_ptx.cuh)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int hq = H >> 2;
  if (idx >= static_cast<long long>(rows) * hq) return;
  const int r = static_cast<int>(idx / hq);
  const int u = static_cast<int>(idx % hq) * 4;
  const long long off = static_cast<long long>(r) * H + u;
  float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c_prev != nullptr) cp = *reinterpret_cast<const float4*>(c_prev + off);
  if (t >= seq_len[r]) {
    *reinterpret_cast<float4*>(c_out + off) = cp;
    uint2 hp = make_uint2(0u, 0u);
    if (h_prev != nullptr) hp = *reinterpret_cast<const uint2*>(h_prev + off);
    *reinterpret_cast<uint2*>(h_out + off) = hp;
    if (h_out_lo != nullptr) {
      uint2 hl = make_uint2(0u, 0u);
      if (h_prev_lo != nullptr) hl = *reinterpret_cast<const uint2*>(h_prev_lo + off);
      *reinterpret_cast<uint2*>(h_out_lo + off) = hl;
    }
    return;
  }
  float z[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + g * H + u));
    z[g][0] = b.x; z[g][1] = b.y; z[g][2] = b.z; z[g][3] = b.w;
  }
  for (int s = 0; s < S; ++s) {
    const float* zp = z_part + s * part_stride + static_cast<long long>(r) * 4 * H + u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 a = *reinterpret_cast<const float4*>(zp + g * H);
      z[g][0] += a.x; z[g][1] += a.y; z[g][2] += a.z; z[g][3] += a.w;
    }
  }
  const float cpa[4] = {cp.x, cp.y, cp.z, cp.w};
  float gi[4], gj[4], gf[4], go[4], cn[4], hn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    gi[k] = sigm_(z[0][k]);
    gj[k] = tanh_(z[1][k]);
    gf[k] = sigm_(z[2][k] + 1.0f);
    go[k] = sigm_(z[3][k]);
    cn[k] = cpa[k] * gf[k] + gi[k] * gj[k];
    hn[k] = tanh_(cn[k]) * go[k];
  }
  *reinterpret_cast<float4*>(c_out + off) = make_float4(cn[0], cn[1], cn[2], cn[3]);
  store4_split(h_out, h_out_lo, off, hn);
  if (gates != nullptr) {
    const long long go_ = static_cast<long long>(r) * 4 * H + u;
    store4_split(gates, gates_lo, go_ + 0 * H, gi);
    store4_split(gates, gates_lo, go_ + 1 * H, gj);
    store4_split(gates, gates_lo, go_ + 2 * H, gf);
    store4_split(gates, gates_lo, go_ + 3 * H, go);
  }
}

// BasicLSTMCell backward at step t: dh = sum_s dh_part[s] (= dz_{t+1} Wh^T) + dh_ext + pass-through;
// gate gradients dz_t (bf16), carried dc, masked pass-through, and the bias gradient db += column sums of dz_t
// (nullable; fused here so that dz is not read a second time by a column-sum pass).
// Block = 8 rows x 32 unit-quads (128 units of one column block), one (row, quad) per thread -- the full-occupancy
// shape that moves 36 B per element at HBM rate; the 8 rows of a block are summed through shared memory and the
// block issues 512 warp-coalesced float atomics (16 per warp).  (A grid-stride variant that kept the sums in
// registers over ~9 rows per thread ran at 3.0 instead of 5.2 TB/s: half the occupancy, no loads in flight across
// iterations.)
__global__ void __launch_bounds__(256, 5)
lstm_cell_bwd_kernel(const float* __restrict__ dh_part, int S, long long part_stride,
                     const __nv_bfloat16* __restrict__ gates, const float* __restrict__ c_prev,
                     const float* __restrict__ dh_ext, long long ld_dh_ext,
                     const float* __restrict__ dh_pass_in, long long ld_dh_pass_in,
                     const float* __restrict__ dc_in, long long ld_dc_in,
                     const int* __restrict__ seq_len, int t, int rows, int H,
                     __nv_bfloat16* __restrict__ dz_out, float* __restrict__ dc_out,
                     float* __restrict__ dh_pass_out, float* __restrict__ dbias,
                     const __nv_bfloat16* __restrict__ gates_lo, __nv_bfloat16* __restrict__ dz_lo, int red_vec) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL (see evc_ptx.cuh)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ __align__(16) float red[8][4][132];               // [row lane][gate][128 units + pad]
  const int ql = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int u = (blockIdx.x * 32 + ql) * 4;      // first of this thread's 4 units
  const int r = blockIdx.y * 8 + rl;
  bool contributes = false;
  if (r < rows) {
    const long long off = static_cast<long long>(r) * H + u;
    const long long zoff = static_cast<long long>(r) * 4 * H + u;
    const int len = seq_len[r];
    const bool live = t < len;
    float dh[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < S; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(dh_part + s * part_stride + off);
      dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w;
    }
    if (dh_ext != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(dh_ext + static_cast<long long>(r) * ld_dh_ext + u);
      dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w;
    }
    if (t + 1 >= len && dh_pass_in != nullptr) {   // row was masked at step t+1 (or t is the last step)
      const float4 a = *reinterpret_cast<const float4*>(dh_pass_in + static_cast<long long>(r) * ld_dh_pass_in + u);
      dh[0] += a.x; dh[1] += a.y; dh[2] += a.z; dh[3] += a.w;
    }
    if (dc_in != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(dc_in + static_cast<long long>(r) * ld_dc_in + u);
      dc[0] = a.x; dc[1] = a.y; dc[2] = a.z; dc[3] = a.w;
    }
    if (!live) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint2*>(dz_out + zoff + g * H) = make_uint2(0u, 0u);
        if (dz_lo != nullptr) *reinterpret_cast<uint2*>(dz_lo + zoff + g * H) = make_uint2(0u, 0u);
      }
      *reinterpret_cast<float4*>(dc_out + off) = make_float4(dc[0], dc[1], dc[2], dc[3]);
      *reinterpret_cast<float4*>(dh_pass_out + off) = make_float4(dh[0], dh[1], dh[2], dh[3]);
    } else {
      float gi[4], gj[4], gf[4], go[4], cp[4] = {0.f, 0.f, 0.f, 0.f};
      load4_split(gates, gates_lo, zoff + 0 * H, gi);
      load4_split(gates, gates_lo, zoff + 1 * H, gj);
      load4_split(gates, gates_lo, zoff + 2 * H, gf);
      load4_split(gates, gates_lo, zoff + 3 * H, go);
      if (c_prev != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(c_prev + off);
        cp[0] = a.x; cp[1] = a.y; cp[2] = a.z; cp[3] = a.w;
      }
      float dz[4][4], dco[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float cn = cp[k] * gf[k] + gi[k] * gj[k];
        const float tc = tanh_(cn);
        const float dcn = dc[k] + dh[k] * go[k] * (1.f - tc * tc);
        dz[0][k] = dcn * gj[k] * gi[k] * (1.f - gi[k]);
        dz[1][k] = dcn * gi[k] * (1.f - gj[k] * gj[k]);
        dz[2][k] = dcn * cp[k] * gf[k] * (1.f - gf[k]);
        dz[3][k] = dh[k] * tc * go[k] * (1.f - go[k]);
        dco[k] = dcn * gf[k];
      }
      contributes = dbias != nullptr;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint2 packed = pack4_bf16(dz[g][0], dz[g][1], dz[g][2], dz[g][3]);
        *reinterpret_cast<uint2*>(dz_out + zoff + g * H) = packed;
        // the bias gradient sums the values the weight-gradient GEMMs see: the bf16-rounded dz (+ its residual plane)
        float q[4];
        unpack4_bf16(packed, q);
        if (dz_lo != nullptr) {
          const uint2 pl = pack4_bf16(dz[g][0] - q[0], dz[g][1] - q[1], dz[g][2] - q[2], dz[g][3] - q[3]);
          *reinterpret_cast<uint2*>(dz_lo + zoff + g * H) = pl;
          float ql4[4];
          unpack4_bf16(pl, ql4);
#pragma unroll
          for (int k = 0; k < 4; ++k) q[k] += ql4[k];
        }
        if (dbias != nullptr) *reinterpret_cast<float4*>(&red[rl][g][ql * 4]) = make_float4(q[0], q[1], q[2], q[3]);
      }
      *reinterpret_cast<float4*>(dc_out + off) = make_float4(dco[0], dco[1], dco[2], dco[3]);
    }
  }
  if (dbias == nullptr) return;
  if (!contributes) {
#pragma unroll
    for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(&red[rl][g][ql * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < 128) {           // 4 gates x 32 quads: one 16-byte vector reduction per thread (red.global.v4.f32)
    const int g = threadIdx.x >> 5, c = (threadIdx.x & 31) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r8 = 0; r8 < 8; ++r8) {
      const float4 a = *reinterpret_cast<const float4*>(&red[r8][g][c]);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    float* dst = dbias + g * H + blockIdx.x * 128 + c;
    if (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f || acc.w != 0.f) {
      if (red_vec && (reinterpret_cast<uintptr_t>(dbias) & 15) == 0) {
        atomicAdd(reinterpret_cast<float4*>(dst), acc);      // sm_90+: one 16-byte RED
      } else {
        atomicAdd(dst + 0, acc.x); atomicAdd(dst + 1, acc.y); atomicAdd(dst + 2, acc.z); atomicAdd(dst + 3, acc.w);
      }
    }
  }
}

// f32 [R,C] -> bf16 [R,ld] (columns C..ld-1 zero): bf16 operand copies of weights / activations
__global__ void cast_bf16_kernel(const float* __restrict__ src, long long R, int C, int ld,
                                 __nv_bfloat16* __restrict__ dst, __nv_bfloat16* __restrict__ dst_lo) {
  PDL_PROLOGUE();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * ld) return;
  const long long r = i / ld;
  const int c = static_cast<int>(i % ld);
  store1_split(dst, dst_lo, i, (c < C) ? src[r * C + c] : 0.f);
}

// --------------------------------------------------------------------------------------
// video_level_models.py:437-447 + losses.py:90-97.  One block per video.
//   p[b,c] = sum_{m<M} softmax(G[b,c,0..M])[m] * sigmoid(E[b,c,m]);  CE row sum (eps = 1e-5)
template <int MAXM>
__device__ __forceinline__ float moe_class(const float* g, const float* e, int M, float* gate, float* sig) {
  float mx = g[0];
  for (int m = 1; m <= M; ++m) mx = fmaxf(mx, g[m]);
  float den = 0.f;
  for (int m = 0; m <= M; ++m) { gate[m] = __expf(g[m] - mx); den += gate[m]; }
  const float inv = 1.0f / den;
  float p = 0.f;
  for (int m = 0; m < M; ++m) { gate[m] *= inv; sig[m] = sigmoidf_(e[m]); p += gate[m] * sig[m]; }
  gate[M] *= inv;
  return p;
}

__global__ void __launch_bounds__(1024) moe_mix_fwd_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ E,
                                   long long lde, int V, int M, float* __restrict__ p_out) {
  PDL_PROLOGUE();
  const int b = blockIdx.x;
  const float* g = G + b * ldg;
  const float* e = E + b * lde;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float gate[9], sig[8];
    p_out[static_cast<long long>(b) * V + c] = moe_class<8>(g + c * (M + 1), e + c * M, M, gate, sig);
  }
}

// Backward of the mixture: dG_k = g_k (s_k - p) dp  (s_M = 0 for the dummy expert),
//                          dE_m = g_m s_m (1 - s_m) dp        (bf16 outputs = GEMM operands)
__global__ void moe_mix_bwd_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ E,
                                   long long lde, const float* __restrict__ dP, int V, int M,
                                   __nv_bfloat16* __restrict__ dG, long long lddg, __nv_bfloat16* __restrict__ dE,
                                   long long ldde, __nv_bfloat16* __restrict__ dG_lo, __nv_bfloat16* __restrict__ dE_lo) {
  PDL_PROLOGUE();
  const int b = blockIdx.x;
  const float* g = G + b * ldg;
  const float* e = E + b * lde;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float gate[9], sig[8];
    const float pc = moe_class<8>(g + c * (M + 1), e + c * M, M, gate, sig);
    const float dp = dP[static_cast<long long>(b) * V + c];
    for (int m = 0; m <= M; ++m) {
      const float s = (m < M) ? sig[m] : 0.f;
      store1_split(dG, dG_lo, b * lddg + c * (M + 1) + m, gate[m] * (s - pc) * dp);
    }
    for (int m = 0; m < M; ++m) store1_split(dE, dE_lo, b * ldde + c * M + m, gate[m] * sig[m] * (1.f - sig[m]) * dp);
  }
}

// losses.py:90-97 CrossEntropyLoss rows and train.py:398-402 KL(Categorical(probs=pT) ||
// Categorical(probs=pS)) rows (both are renormalised), plus the gradient of
//   ce_scale * CE_row + kl_scale * KL_row   w.r.t. the student predictions:
//   dCE/dp = -(y/(p+eps)) + (1-y)/(1-p+eps) ;  dKL/dp_c = -pT_hat_c / p_c + 1 / sum(pS)
__global__ void ce_kl_loss_kernel(const float* __restrict__ P, const float* __restrict__ PT,
                                  const uint8_t* __restrict__ labels, int V, float ce_scale, float kl_scale,
                                  float* __restrict__ ce_rows, float* __restrict__ kl_rows,
                                  float* __restrict__ dP) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* p = P + static_cast<long long>(b) * V;
  const float* pt = PT ? PT + static_cast<long long>(b) * V : nullptr;
  const uint8_t* y8 = labels ? labels + static_cast<long long>(b) * V : nullptr;
  float inv_sT = 0.f, inv_sS = 0.f;
  if (pt) {
    float st = 0.f, ss = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) { st += pt[c]; ss += p[c]; }
    st = block_sum(st, sh);
    ss = block_sum(ss, sh);
    inv_sT = 1.0f / st;
    inv_sS = 1.0f / ss;
  }
  float kl = 0.f, ce = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    const float pc = p[c];
    float dp = 0.f;
    if (y8) {
      const float y = y8[c] ? 1.f : 0.f;
      ce -= y * __logf(pc + 1e-5f) + (1.f - y) * __logf(1.f - pc + 1e-5f);
      dp += ce_scale * (-(y / (pc + 1e-5f)) + (1.f - y) / (1.f - pc + 1e-5f));
    }
    if (pt) {
      const float th = pt[c] * inv_sT;
      dp += kl_scale * (-th / pc + inv_sS);
      if (th > 0.f) kl += th * (__logf(th) - __logf(pc * inv_sS));
    }
    if (dP) dP[static_cast<long long>(b) * V + c] = dp;
  }
  if (ce_rows) {
    ce = block_sum(ce, sh);
    if (threadIdx.x == 0) ce_rows[b] = ce;
  }
  if (kl_rows) {
    kl = block_sum(kl, sh);
    if (threadIdx.x == 0) kl_rows[b] = kl;
  }
}

// Fused classifier head of the training step (video_level_models.py:437-447 + losses.py:90-97 +
// train.py:398-402): mixture forward, CE row, KL(teacher || student) row, d(loss)/dp and the
// gradients w.r.t. the gate / expert logits in ONE launch per model (one block per video).
//   pass 1: p[c] (written, kept in L2) and, for the KL normalisers, sum p and sum pT
//   pass 2: CE/KL terms, dp = ce_scale*dCE/dp + kl_scale*dKL/dp, dG/dE (bf16 GEMM operands)
__global__ void __launch_bounds__(1024) moe_mix_loss_kernel(const float* __restrict__ G, long long ldg, const float* __restrict__ E,
                                    long long lde, const float* __restrict__ PT,
                                    const uint8_t* __restrict__ labels, int V, int M, float ce_scale,
                                    float kl_scale, float* __restrict__ P, float* __restrict__ ce_rows,
                                    float* __restrict__ kl_rows, __nv_bfloat16* __restrict__ dG, long long lddg,
                                    __nv_bfloat16* __restrict__ dE, long long ldde,
                                    __nv_bfloat16* __restrict__ dG_lo, __nv_bfloat16* __restrict__ dE_lo) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* g = G + b * ldg;
  const float* e = E + b * lde;
  float* p = P + static_cast<long long>(b) * V;
  const float* pt = PT ? PT + static_cast<long long>(b) * V : nullptr;
  const uint8_t* y8 = labels + static_cast<long long>(b) * V;
  float ss = 0.f, st = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float gate[9], sig[8];
    const float pc = moe_class<8>(g + c * (M + 1), e + c * M, M, gate, sig);
    p[c] = pc;
    ss += pc;
    if (pt) st += pt[c];
  }
  float inv_sT = 0.f, inv_sS = 0.f;
  if (pt) {
    ss = block_sum(ss, sh);
    st = block_sum(st, sh);
    inv_sS = 1.0f / ss;
    inv_sT = 1.0f / st;
  }
  float ce = 0.f, kl = 0.f;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    float gate[9], sig[8];
    const float pc = moe_class<8>(g + c * (M + 1), e + c * M, M, gate, sig);
    const float y = y8[c] ? 1.f : 0.f;
    ce -= y * __logf(pc + 1e-5f) + (1.f - y) * __logf(1.f - pc + 1e-5f);
    float dp = ce_scale * (-(y / (pc + 1e-5f)) + (1.f - y) / (1.f - pc + 1e-5f));
    if (pt) {
      const float th = pt[c] * inv_sT;
      dp += kl_scale * (-th / pc + inv_sS);
      if (th > 0.f) kl += th * (__logf(th) - __logf(pc * inv_sS));
    }
    for (int m = 0; m <= M; ++m) {
      const float sgm = (m < M) ? sig[m] : 0.f;
      store1_split(dG, dG_lo, b * lddg + c * (M + 1) + m, gate[m] * (sgm - pc) * dp);
    }
    for (int m = 0; m < M; ++m) store1_split(dE, dE_lo, b * ldde + c * M + m, gate[m] * sig[m] * (1.f - sig[m]) * dp);
  }
  ce = block_sum(ce, sh);
  if (threadIdx.x == 0) ce_rows[b] = ce;
  if (kl_rows) {
    kl = block_sum(kl, sh);
    if (threadIdx.x == 0) kl_rows[b] = kl;
  }
}

// out[0] = scale * sum_i rows[i]  (reduce_mean / reduce_sum over the batch); single block
__global__ void reduce_rows_kernel(const float* __restrict__ rows, int n, float scale, float* __restrict__ out) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += rows[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

// [TF adam.py _prepare/_finish] t += 1 ; lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
__global__ void adam_lr_kernel(long long* __restrict__ step, float lr, float b1, float b2,
                               float* __restrict__ lr_t) {
  PDL_PROLOGUE();
  const long long t = step[0] + 1;
  step[0] = t;
  const double td = static_cast<double>(t);
  lr_t[0] = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), td)) /
                               (1.0 - pow(static_cast<double>(b1), td)));
}

// train.py:359-362  L_REP rows = sum_j (t - s)^2 ;  dS = grad_scale * (s - t)
__global__ void rep_loss_kernel(const float* __restrict__ t_state, const float* __restrict__ s_state, int S,
                                float grad_scale, float* __restrict__ rows, float* __restrict__ d_s) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const long long o = static_cast<long long>(b) * S + j;
    const float d = s_state[o] - t_state[o];
    acc += d * d;
    if (d_s) d_s[o] = grad_scale * d;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) rows[b] = acc;
}

// bias gradients: out[n] += sum_r X[r, n]  (bf16 X with row pitch ld).  out must be zeroed.
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, long long R, int N, long long ld,
                                   long long rows_per_block, float* __restrict__ out) {
  PDL_PROLOGUE();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = blockIdx.y * rows_per_block;
  const long long r1 = (r0 + rows_per_block < R) ? r0 + rows_per_block : R;
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += __bfloat162float(X[r * ld + n]);
  atomicAdd(out + n, acc);
}

// out[0] += sum (g + wd*w)^2     (slim clip_gradient_norms: per-variable l2 norm; wd*w is the
// gradient of penalty * l2_regularizer(1e-8)(w), video_level_models.py:428,434 / train.py:324)
__global__ void sumsq_kernel(const float* __restrict__ g, const float* __restrict__ w, float wd, long long n,
                             float* __restrict__ out, float* __restrict__ out_wsq) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  float acc = 0.f, wacc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 a = *reinterpret_cast<const float4*>(g + i);
      if (w != nullptr) {
        const float4 ww = *reinterpret_cast<const float4*>(w + i);
        a.x += wd * ww.x; a.y += wd * ww.y; a.z += wd * ww.z; a.w += wd * ww.w;
        wacc += ww.x * ww.x + ww.y * ww.y + ww.z * ww.z + ww.w * ww.w;
      }
      acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    } else {
      for (long long j = i; j < n; ++j) {
        const float a = g[j] + (w != nullptr ? wd * w[j] : 0.f);
        acc += a * a;
        if (w != nullptr) wacc += w[j] * w[j];
      }
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, acc);
  if (out_wsq != nullptr) {
    wacc = block_sum(wacc, sh);
    if (threadIdx.x == 0) atomicAdd(out_wsq, wacc);
  }
}

// [TF clip_ops.clip_by_norm] g * c * min(rsqrt(sum g^2), 1/c), then [TF ApplyAdam]
//   m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= lr_t * m / (sqrt(v) + eps)
// and refresh of the bf16 operand copy of w (row pitch ld_shadow, `cols` columns per row).
// The squared norm of the regularised gradient g + wd*w is  *normsq  (a sumsq pass)  +  the parts taken where they
// were cheap (all nullable):  *normsq_fused = sum g^2 from the weight-gradient GEMM's epilogue,  *reg_cross = <g, w>
// (evc_reg_cross)  and  *reg_wsq = sum w^2:   |g + wd w|^2 = |g|^2 + 2 wd <g,w> + wd^2 |w|^2.
// wsq_out (nullable): += sum of the squares of the UPDATED weights (next step's reg_wsq / regulariser value).
__global__ void clip_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ normsq, float clip,
                                 float wd, const float* __restrict__ lr_t, float b1, float b2, float eps,
                                 __nv_bfloat16* __restrict__ shadow, int cols, long long ld_shadow,
                                 __nv_bfloat16* __restrict__ shadow_lo, const float* __restrict__ normsq_fused,
                                 const float* __restrict__ reg_cross, const float* __restrict__ reg_wsq,
                                 float* __restrict__ wsq_out) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  float scale = 1.f;
  if (clip > 0.f) {
    float ns = *normsq;
    if (normsq_fused != nullptr) ns += *normsq_fused;
    if (reg_cross != nullptr) ns += 2.f * wd * *reg_cross;
    if (reg_wsq != nullptr) ns += wd * wd * *reg_wsq;
    scale = (ns > 0.f) ? clip * fminf(rsqrtf(ns), 1.0f / clip) : 1.f;
  }
  const float lr = *lr_t;
  const float c1 = 1.f - b1, c2 = 1.f - b2;
  float wacc = 0.f;
  // 4 parameters per thread and iteration (every tensor size and `cols` is a multiple of 4)
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const long long i = i4 << 2;
    float4 wv = *reinterpret_cast<const float4*>(w + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    float wa[4] = {wv.x, wv.y, wv.z, wv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gi = (ga[k] + wd * wa[k]) * scale;
      ma[k] += (gi - ma[k]) * c1;
      va[k] += (gi * gi - va[k]) * c2;
      wa[k] -= lr * ma[k] / (sqrtf(va[k]) + eps);
      wacc = fmaf(wa[k], wa[k], wacc);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    *reinterpret_cast<float4*>(w + i) = make_float4(wa[0], wa[1], wa[2], wa[3]);
    if (shadow) store4_split(shadow, shadow_lo, (i / cols) * ld_shadow + (i % cols), wa);
  }
  if (wsq_out != nullptr) {
    wacc = block_sum(wacc, sh);
    if (threadIdx.x == 0) atomicAdd(wsq_out, wacc);
  }
}

// <g, w> of a fully connected layer without touching g or w:  g = X^T dL (the weight-gradient GEMM) and
// logits = X w + bias, hence  <g, w> = <X^T dL, w> = <dL, X w> = sum_b sum_n dL[b,n] * (logits[b,n] - bias[n]).
// dL: bf16 gradient w.r.t. the logits as the weight-gradient GEMM reads it (+ residual plane in split-bf16 mode).
// out[0] += the sum.  One block per row.
__global__ void reg_cross_kernel(const float* __restrict__ logits, long long ld_logits,
                                 const __nv_bfloat16* __restrict__ dl, const __nv_bfloat16* __restrict__ dl_lo,
                                 long long ld_dl, const float* __restrict__ bias, int N, int vec,
                                 float* __restrict__ out) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  const float* lrow = logits + static_cast<long long>(blockIdx.x) * ld_logits;
  const __nv_bfloat16* drow = dl + static_cast<long long>(blockIdx.x) * ld_dl;
  const __nv_bfloat16* drow_lo = dl_lo ? dl_lo + static_cast<long long>(blockIdx.x) * ld_dl : nullptr;
  float acc = 0.f;
  const int t = blockIdx.y * blockDim.x + threadIdx.x, nt = gridDim.y * blockDim.x;
  int n0 = 0;
  if (vec) {                       // 4 columns per thread and iteration: 16-byte logits, 8-byte gradient loads
    const int n4 = N >> 2;
    for (int q = t; q < n4; q += nt) {
      const float4 l = *reinterpret_cast<const float4*>(lrow + 4 * q);
      float d[4];
      load4_split(drow, drow_lo, 4 * q, d);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(bias + 4 * q));
      acc = fmaf(d[0], l.x - b.x, acc);
      acc = fmaf(d[1], l.y - b.y, acc);
      acc = fmaf(d[2], l.z - b.z, acc);
      acc = fmaf(d[3], l.w - b.w, acc);
    }
    n0 = n4 << 2;
  }
  for (int n = n0 + t; n < N; n += nt) {
    float d = __bfloat162float(drow[n]);
    if (drow_lo != nullptr) d += __bfloat162float(drow_lo[n]);
    acc = fmaf(d, lrow[n] - (bias != nullptr ? __ldg(bias + n) : 0.f), acc);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(out, acc);
}

// eval_util.py:118-124 top_k_triplets: the k largest predictions of a video.  Exact; the
// reference's argpartition leaves ties at the boundary implementation-defined, here the
// lower class index wins and the output is ordered by value descending.  One block per video.
__global__ void topk_kernel(const float* __restrict__ P, int V, int k, const uint8_t* __restrict__ labels,
                            int* __restrict__ idx_out, float* __restrict__ val_out, uint8_t* __restrict__ lab_out) {
  PDL_PROLOGUE();
  extern __shared__ unsigned long long keys[];  // V keys + 32 scratch
  unsigned long long* red = keys + V;
  const int b = blockIdx.x;
  const float* p = P + static_cast<long long>(b) * V;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    uint32_t u = __float_as_uint(p[c]);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone map float -> uint
    keys[c] = (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(0xFFFFFFFFu - c);
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = threadIdx.x; c < V; c += blockDim.x) best = keys[c] > best ? keys[c] : best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
      }
      if (threadIdx.x == 0) {
        const int c = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(v & 0xFFFFFFFFull));
        idx_out[b * k + r] = c;
        val_out[b * k + r] = p[c];
        if (lab_out) lab_out[b * k + r] = labels ? labels[static_cast<long long>(b) * V + c] : 0;
        keys[c] = 0ull;  // remove from later rounds
      }
    }
    __syncthreads();
  }
}

// eval_util.py:34-59 calculate_precision_at_equal_recall_rate, per video, without a sort: a labelled class c is
// among the video's top-num_labels predictions iff fewer than num_labels classes rank before it under the top-k
// kernel's order (value descending, lower class index first).  rows[b] = |{labelled c in the top-n with p > 0}| / n
// (0 for a video without labels), npos[b] = n, class_pos[c] += 1 for every labelled class (nullable; the
// per-class positives of eval_util.py:114).  One block per video, one warp per labelled class.
__global__ void video_perr_kernel(const float* __restrict__ P, const uint8_t* __restrict__ labels, int V,
                                  float* __restrict__ perr_rows, int* __restrict__ npos_rows,
                                  int* __restrict__ class_pos) {
  PDL_PROLOGUE();
  extern __shared__ int lab_list[];           // labelled classes of this video (<= V entries)
  __shared__ int s_count, s_hits;
  const int b = blockIdx.x;
  const float* p = P + static_cast<long long>(b) * V;
  const uint8_t* y = labels + static_cast<long long>(b) * V;
  if (threadIdx.x == 0) { s_count = 0; s_hits = 0; }
  __syncthreads();
  for (int c = threadIdx.x; c < V; c += blockDim.x)
    if (y[c]) {
      lab_list[atomicAdd(&s_count, 1)] = c;
      if (class_pos) atomicAdd(class_pos + c, 1);
    }
  __syncthreads();
  const int n = s_count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int e = warp; e < n; e += nwarps) {
    const int c = lab_list[e];
    const float pc = p[c];
    int before = 0;
    for (int j = lane; j < V; j += 32) {
      const float pj = p[j];
      before += (pj > pc || (pj == pc && j < c)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0 && before < n && pc > 0.f) atomicAdd(&s_hits, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    perr_rows[b] = n > 0 ? static_cast<float>(s_hits) / static_cast<float>(n) : 0.f;
    npos_rows[b] = n;
  }
}

// eval_util.py:61-79 calculate_gap on the batch's pooled top-k triplets (average_precision_calculator.py:166-232:
// sort by prediction descending, AP = sum over hits of precision-at-rank / numpos) by rank counting instead of a
// sort: for a hit i, rank = 1 + #{j before i}, positives so far = 1 + #{hits j before i}, "before" = larger
// prediction, ties by (class, video) ascending -- the order the reference's regrouping by class feeds its sort.
// acc[0] += sum of precisions (double).  Thread = one triplet; only hits loop.
__global__ void gap_rank_kernel(const int* __restrict__ idx, const float* __restrict__ val,
                                const uint8_t* __restrict__ lab, int n, int k, double* __restrict__ acc) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (i < n && lab[i]) {
    const float vi = val[i];
    const int ci = idx[i], bi = i / k;
    int before = 0, hits_before = 0;
    for (int j = 0; j < n; ++j) {
      const float vj = val[j];
      bool first = vj > vi;
      if (vj == vi) {
        const int cj = idx[j], bj = j / k;
        first = cj < ci || (cj == ci && bj < bi);
      }
      before += first ? 1 : 0;
      hits_before += (first && lab[j]) ? 1 : 0;
    }
    contrib = static_cast<float>(hits_before + 1) / static_cast<float>(before + 1);
  }
  contrib = block_sum(contrib, sh);
  if (threadIdx.x == 0 && contrib != 0.f) atomicAdd(acc, static_cast<double>(contrib));
}

// out[0] = mean hit@1 (label of every video's top prediction, eval_util.py:17-31), out[1] = mean PERR,
// out[2] = GAP = acc / sum(npos) (0 without positives); sums[0..3] += (n videos, hit sum, perr sum, loss sum)
// when non-null (the epoch accumulators of EvaluationMetrics).  Single block.
__global__ void batch_metrics_finalize_kernel(const uint8_t* __restrict__ lab, int B, int k,
                                              const float* __restrict__ perr_rows, const int* __restrict__ npos_rows,
                                              const float* __restrict__ loss_rows, double* __restrict__ acc,
                                              float* __restrict__ out, double* __restrict__ sums) {
  PDL_PROLOGUE();
  __shared__ float sh[32];
  float hit = 0.f, perr = 0.f, npos = 0.f, loss = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    hit += lab[static_cast<long long>(b) * k] ? 1.f : 0.f;
    perr += perr_rows[b];
    npos += static_cast<float>(npos_rows[b]);
    if (loss_rows) loss += loss_rows[b];
  }
  hit = block_sum(hit, sh);
  perr = block_sum(perr, sh);
  npos = block_sum(npos, sh);
  loss = block_sum(loss, sh);
  if (threadIdx.x == 0) {
    out[0] = B > 0 ? hit / B : 0.f;
    out[1] = B > 0 ? perr / B : 0.f;
    out[2] = npos > 0.f ? static_cast<float>(acc[0] / static_cast<double>(npos)) : 0.f;
    out[3] = B > 0 ? loss / B : 0.f;
    acc[0] = 0.0;                                   // ready for the next batch
    if (sums) { sums[0] += B; sums[1] += hit; sums[2] += perr; sums[3] += loss; }
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  PDL_PROLOGUE();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

inline int grid_for(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define EVC_STREAM(s) static_cast<cudaStream_t>(s)

namespace evc {
int launch_lstm_cell_fwd(const float* z_part, int S, long long part_stride, const float* bias, const float* c_prev,
                         const void* h_prev, const int* seq_len, int t, int rows, int H, float* c_out, void* h_out,
                         void* gates, cudaStream_t stream, const void* h_prev_lo, void* h_out_lo, void* gates_lo) {
  const long long n = static_cast<long long>(rows) * (H / 4);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((n + 255) / 256));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, lstm_cell_fwd_kernel, z_part, S, part_stride, bias, c_prev,
                     static_cast<const __nv_bfloat16*>(h_prev), seq_len, t, rows, H, c_out,
                     static_cast<__nv_bfloat16*>(h_out), static_cast<__nv_bfloat16*>(gates),
                     static_cast<const __nv_bfloat16*>(h_prev_lo), static_cast<__nv_bfloat16*>(h_out_lo),
                     static_cast<__nv_bfloat16*>(gates_lo));
  count_launch();
  return check_launch("lstm_cell_fwd");
}
int launch_lstm_cell_bwd(const float* dh_part, int S, long long part_stride, const void* gates, const float* c_prev,
                         const float* dh_ext, long long ld_dh_ext, const float* dh_pass_in, long long ld_dh_pass_in,
                         const float* dc_in, long long ld_dc_in, const int* seq_len, int t, int rows, int H,
                         void* dz_out, float* dc_out, float* dh_pass_out, float* dbias, cudaStream_t stream,
                         const void* gates_lo, void* dz_lo) {
  if (H % 128 != 0) return set_error(EVC_ERR_ARG, "lstm_cell_bwd: H must be a multiple of 128");
  const int col_blocks = H / 128;
  const int row_groups = (rows + 7) / 8;
  if (row_groups > 65535) return set_error(EVC_ERR_UNSUPPORTED, "lstm_cell_bwd: more than 524280 rows");
  static int red_vec = -1;   // EVC_BIAS_RED=1: scalar float atomics for the bias sums (A/B experiment; default 16-byte REDs)
  if (red_vec < 0) {
    const char* e = getenv("EVC_BIAS_RED");
    red_vec = (e && atoi(e) == 1) ? 0 : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(col_blocks), static_cast<unsigned>(row_groups));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, lstm_cell_bwd_kernel, dh_part, S, part_stride, static_cast<const __nv_bfloat16*>(gates),
                     c_prev, dh_ext, ld_dh_ext, dh_pass_in, ld_dh_pass_in, dc_in, ld_dc_in, seq_len, t, rows, H,
                     static_cast<__nv_bfloat16*>(dz_out), dc_out, dh_pass_out, dbias,
                     static_cast<const __nv_bfloat16*>(gates_lo), static_cast<__nv_bfloat16*>(dz_lo), red_vec);
  count_launch();
  return check_launch("lstm_cell_bwd");
}
}  // namespace evc

extern "C" int evc_frames_pack(const float* src, int B, int T, int D, const int* frame_idx, int idx_per_batch,
                               int K, int num_chunks, int normalize, void* out_bf16, float* out_f32, void* out_lo,
                               void* stream) {
  if (B <= 0 || K <= 0) return set_error(EVC_ERR_ARG, "frames_pack: empty batch");
  if (D % 4 != 0) return set_error(EVC_ERR_ARG, "frames_pack: feature size must be a multiple of 4");
  if (num_chunks <= 0 || K % num_chunks != 0)
    return set_error(EVC_ERR_ARG, "frames_pack: number of frames must split evenly into chunks (tf.split)");
  const long long warps = static_cast<long long>(B) * K;
  const int block = 256;
  const int grid = static_cast<int>((warps * 32 + block - 1) / block);
  pdl_launch(frames_pack_kernel, dim3(grid), dim3(block), 0, EVC_STREAM(stream), src, B, T, D, frame_idx, idx_per_batch, K, num_chunks,
                                                             normalize, static_cast<__nv_bfloat16*>(out_bf16),
                                                             out_f32, static_cast<__nv_bfloat16*>(out_lo));
  count_launch();
  return check_launch("frames_pack");
}

extern "C" int evc_frames_pack_u8(const unsigned char* src, const int* num_frames, int B, int T, int D,
                                  const int* frame_idx, int idx_per_batch, int K, int num_chunks, int normalize,
                                  void* out_bf16, float* out_f32, void* out_lo, void* stream) {
  if (B <= 0 || K <= 0) return set_error(EVC_ERR_ARG, "frames_pack_u8: empty batch");
  if (D % 4 != 0) return set_error(EVC_ERR_ARG, "frames_pack_u8: feature size must be a multiple of 4");
  if (num_frames == nullptr) return set_error(EVC_ERR_ARG, "frames_pack_u8: num_frames required (zero padding)");
  if (num_chunks <= 0 || K % num_chunks != 0)
    return set_error(EVC_ERR_ARG, "frames_pack_u8: number of frames must split evenly into chunks (tf.split)");
  const long long warps = static_cast<long long>(B) * K;
  const int block = 256;
  const int grid = static_cast<int>((warps * 32 + block - 1) / block);
  pdl_launch(frames_pack_u8_kernel, dim3(grid), dim3(block), 0, EVC_STREAM(stream), src, num_frames, B, T, D, frame_idx, idx_per_batch, K,
                                                                num_chunks, normalize,
                                                                static_cast<__nv_bfloat16*>(out_bf16), out_f32,
                                                                static_cast<__nv_bfloat16*>(out_lo));
  count_launch();
  return check_launch("frames_pack_u8");
}

extern "C" int evc_num_frames_student(const int* num_frames, int B, int max_frames, int every_n, long long* out,
                                      void* stream) {
  if (every_n <= 0) return set_error(EVC_ERR_ARG, "num_frames_student: every_n must be positive");
  pdl_launch(num_frames_student_kernel, dim3((B + 127) / 128), dim3(128), 0, EVC_STREAM(stream), num_frames, B, max_frames,
                                                                             max_frames / every_n, out);
  count_launch();
  return check_launch("num_frames_student");
}

extern "C" int evc_lstm_lengths(const void* num_frames, int is_int64, int B, int num_chunks, int chunk_len,
                                int* len_l1, int* len_l2, void* stream) {
  const int n = B * num_chunks;
  pdl_launch(lstm_lengths_kernel, dim3((n + 127) / 128), dim3(128), 0, EVC_STREAM(stream), num_frames, is_int64, B, num_chunks,
                                                                       chunk_len, len_l1, len_l2);
  count_launch();
  return check_launch("lstm_lengths");
}

extern "C" int evc_random_frame_index(const float* u, const int* num_frames, int B, int K, int* idx, void* stream) {
  pdl_launch(random_frame_index_kernel, dim3((B * K + 127) / 128), dim3(128), 0, EVC_STREAM(stream), u, num_frames, B, K, idx);
  count_launch();
  return check_launch("random_frame_index");
}
extern "C" int evc_random_sequence_index(const float* u, const int* num_frames, int B, int K, int* idx,
                                         void* stream) {
  pdl_launch(random_sequence_index_kernel, dim3((B * K + 127) / 128), dim3(128), 0, EVC_STREAM(stream), u, num_frames, B, K, idx);
  count_launch();
  return check_launch("random_sequence_index");
}

extern "C" int evc_sampled_lengths(const int* num_frames, int B, int K, long long* out, void* stream) {
  pdl_launch(sampled_lengths_kernel, dim3((B + 127) / 128), dim3(128), 0, EVC_STREAM(stream), num_frames, B, K, out);
  count_launch();
  return check_launch("sampled_lengths");
}

extern "C" int evc_random_uniform(unsigned long long seed, unsigned long long offset, float* out, long long n,
                                  void* stream) {
  if (n <= 0) return EVC_OK;
  const long long groups = (n + 3) / 4;
  pdl_launch(random_uniform_kernel, dim3(static_cast<unsigned>((groups + 127) / 128)), dim3(128), 0, EVC_STREAM(stream),
             seed, offset, out, n);
  count_launch();
  return check_launch("random_uniform");
}

extern "C" int evc_state_pack(const float* c0, const void* h0, const float* c1, const void* h1, int rows, int H,
                              void* out_bf16, float* out_f32, const void* h0_lo, const void* h1_lo, void* out_lo,
                              void* stream) {
  if ((h0_lo == nullptr) != (h1_lo == nullptr)) return set_error(EVC_ERR_ARG, "state_pack: both h lo planes or none");
  const long long n = static_cast<long long>(rows) * H;
  pdl_launch(state_pack_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, EVC_STREAM(stream), 
      c0, static_cast<const __nv_bfloat16*>(h0), c1, static_cast<const __nv_bfloat16*>(h1), n, H,
      static_cast<__nv_bfloat16*>(out_bf16), out_f32, static_cast<const __nv_bfloat16*>(h0_lo),
      static_cast<const __nv_bfloat16*>(h1_lo), static_cast<__nv_bfloat16*>(out_lo));
  count_launch();
  return check_launch("state_pack");
}

extern "C" int evc_cast_bf16(const float* src, long long rows, int cols, int ld, void* dst, void* dst_lo,
                             void* stream) {
  if (ld < cols) return set_error(EVC_ERR_ARG, "cast_bf16: ld < cols");
  const long long n = rows * ld;
  pdl_launch(cast_bf16_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, EVC_STREAM(stream), 
      src, rows, cols, ld, static_cast<__nv_bfloat16*>(dst), static_cast<__nv_bfloat16*>(dst_lo));
  count_launch();
  return check_launch("cast_bf16");
}

extern "C" int evc_moe_mix_fwd(const float* G, long long ldg, const float* E, long long lde, int B, int V, int M,
                               float* p_out, void* stream) {
  if (M < 1 || M > 8) return set_error(EVC_ERR_UNSUPPORTED, "moe_mix: 1 <= num_mixtures <= 8");
  if (B <= 0) return EVC_OK;
  pdl_launch(moe_mix_fwd_kernel, dim3(B), dim3(1024), 0, EVC_STREAM(stream), G, ldg, E, lde, V, M, p_out);
  count_launch();
  return check_launch("moe_mix_fwd");
}

extern "C" int evc_moe_mix_bwd(const float* G, long long ldg, const float* E, long long lde, const float* dP, int B,
                               int V, int M, void* dG, long long lddg, void* dE, long long ldde, void* dG_lo,
                               void* dE_lo, void* stream) {
  if (M < 1 || M > 8) return set_error(EVC_ERR_UNSUPPORTED, "moe_mix_bwd: 1 <= num_mixtures <= 8");
  if (B <= 0) return EVC_OK;
  pdl_launch(moe_mix_bwd_kernel, dim3(B), dim3(256), 0, EVC_STREAM(stream), G, ldg, E, lde, dP, V, M, static_cast<__nv_bfloat16*>(dG),
                                                        lddg, static_cast<__nv_bfloat16*>(dE), ldde,
                                                        static_cast<__nv_bfloat16*>(dG_lo), static_cast<__nv_bfloat16*>(dE_lo));
  count_launch();
  return check_launch("moe_mix_bwd");
}

extern "C" int evc_ce_kl_loss(const float* P, const float* PT, const unsigned char* labels, int B, int V,
                              float ce_scale, float kl_scale, float* ce_rows, float* kl_rows, float* dP,
                              void* stream) {
  if (labels == nullptr && PT == nullptr) return set_error(EVC_ERR_ARG, "ce_kl_loss: labels or teacher needed");
  if (B <= 0) return EVC_OK;
  pdl_launch(ce_kl_loss_kernel, dim3(B), dim3(256), 0, EVC_STREAM(stream), P, PT, labels, V, ce_scale, kl_scale, ce_rows, kl_rows, dP);
  count_launch();
  return check_launch("ce_kl_loss");
}

extern "C" int evc_moe_mix_loss(const float* G, long long ldg, const float* E, long long lde, const float* PT,
                                const unsigned char* labels, int B, int V, int M, float ce_scale, float kl_scale,
                                float* P, float* ce_rows, float* kl_rows, void* dG, long long lddg, void* dE,
                                long long ldde, void* dG_lo, void* dE_lo, void* stream) {
  if (M < 1 || M > 8) return set_error(EVC_ERR_UNSUPPORTED, "moe_mix_loss: 1 <= num_mixtures <= 8");
  if (labels == nullptr || ce_rows == nullptr) return set_error(EVC_ERR_ARG, "moe_mix_loss: labels and ce_rows required");
  if (B <= 0) return EVC_OK;
  // one block per video, 1024 threads: with 256 videos the grid is only 1.7 blocks per SM, so the parallelism has to
  // come from inside the block (measured 70 us -> see profiles/r02 at 256 threads: latency-bound at 16 warps per SM)
  pdl_launch(moe_mix_loss_kernel, dim3(B), dim3(1024), 0, EVC_STREAM(stream), G, ldg, E, lde, PT, labels, V, M, ce_scale,
             kl_scale, P, ce_rows, kl_rows, static_cast<__nv_bfloat16*>(dG), lddg, static_cast<__nv_bfloat16*>(dE), ldde,
             static_cast<__nv_bfloat16*>(dG_lo), static_cast<__nv_bfloat16*>(dE_lo));
  count_launch();
  return check_launch("moe_mix_loss");
}

extern "C" int evc_reduce_rows(const float* rows, int n, float scale, float* out, void* stream) {
  pdl_launch(reduce_rows_kernel, dim3(1), dim3(256), 0, EVC_STREAM(stream), rows, n, scale, out);
  count_launch();
  return check_launch("reduce_rows");
}

extern "C" int evc_adam_lr(long long* step, float lr, float beta1, float beta2, float* lr_t, void* stream) {
  pdl_launch(adam_lr_kernel, dim3(1), dim3(1), 0, EVC_STREAM(stream), step, lr, beta1, beta2, lr_t);
  count_launch();
  return check_launch("adam_lr");
}

extern "C" int evc_rep_loss(const float* teacher_state, const float* student_state, int B, int S, float grad_scale,
                            float* rows, float* d_student, void* stream) {
  pdl_launch(rep_loss_kernel, dim3(B), dim3(256), 0, EVC_STREAM(stream), teacher_state, student_state, S, grad_scale, rows, d_student);
  count_launch();
  return check_launch("rep_loss");
}

extern "C" int evc_colsum_bf16(const void* X, long long rows, int N, long long ld, float* out, void* stream) {
  long long rpb = 256;
  long long gy = (rows + rpb - 1) / rpb;
  if (gy > 2048) { gy = 2048; rpb = (rows + gy - 1) / gy; gy = (rows + rpb - 1) / rpb; }
  dim3 grid((N + 127) / 128, static_cast<unsigned>(gy));
  pdl_launch(colsum_bf16_kernel, dim3(grid), dim3(128), 0, EVC_STREAM(stream), static_cast<const __nv_bfloat16*>(X), rows, N, ld, rpb,
                                                           out);
  count_launch();
  return check_launch("colsum_bf16");
}

extern "C" int evc_fill_f32(float* p, long long n, float value, void* stream) {
  pdl_launch(fill_f32_kernel, dim3(grid_for(n, 256, 148 * 8)), dim3(256), 0, EVC_STREAM(stream), p, n, value);
  count_launch();
  return check_launch("fill_f32");
}

extern "C" int evc_sumsq(const float* g, const float* w, float weight_decay, long long n, float* out,
                         float* out_wsq, void* stream) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) || (w && (reinterpret_cast<uintptr_t>(w) & 15)))
    return set_error(EVC_ERR_ARG, "sumsq: pointers must be 16-byte aligned");
  pdl_launch(sumsq_kernel, dim3(grid_for((n + 3) / 4, 256, 148 * 8)), dim3(256), 0, EVC_STREAM(stream), g, w, w ? weight_decay : 0.f, n,
                                                                                    out, w ? out_wsq : nullptr);
  count_launch();
  return check_launch("sumsq");
}

static int clip_adam_impl(float* w, const float* g, float* m, float* v, long long n, const float* normsq,
                          float clip_norm, float weight_decay, const float* lr_t, float beta1, float beta2, float eps,
                          void* shadow_bf16, int cols, long long ld_shadow, void* shadow_lo, const float* normsq_fused,
                          const float* reg_cross, const float* reg_wsq, float* wsq_out, void* stream) {
  if (shadow_bf16 && cols <= 0) return set_error(EVC_ERR_ARG, "clip_adam: cols required with a bf16 copy");
  if (n % 4 != 0 || (shadow_bf16 && (cols % 4 != 0 || ld_shadow % 4 != 0)) ||
      ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15))
    return set_error(EVC_ERR_ARG, "clip_adam: tensors must be 16-byte aligned with sizes/cols multiple of 4");
  if (normsq == nullptr && clip_norm > 0.f) return set_error(EVC_ERR_ARG, "clip_adam: normsq required when clipping");
  pdl_launch(clip_adam_kernel, dim3(grid_for(n / 4, 256, 148 * 8)), dim3(256), 0, EVC_STREAM(stream), 
      w, g, m, v, n, normsq, clip_norm, weight_decay, lr_t, beta1, beta2, eps,
      static_cast<__nv_bfloat16*>(shadow_bf16), cols > 0 ? cols : 1, ld_shadow,
      static_cast<__nv_bfloat16*>(shadow_lo), normsq_fused, reg_cross, reg_wsq, wsq_out);
  count_launch();
  return check_launch("clip_adam");
}

extern "C" int evc_clip_adam(float* w, const float* g, float* m, float* v, long long n, const float* normsq,
                             float clip_norm, float weight_decay, const float* lr_t, float beta1, float beta2,
                             float eps, void* shadow_bf16, int cols, long long ld_shadow, void* shadow_lo,
                             void* stream) {
  return clip_adam_impl(w, g, m, v, n, normsq, clip_norm, weight_decay, lr_t, beta1, beta2, eps, shadow_bf16, cols,
                        ld_shadow, shadow_lo, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int evc_clip_adam_fused(float* w, const float* g, float* m, float* v, long long n, const float* normsq,
                                   float clip_norm, float weight_decay, const float* lr_t, float beta1, float beta2,
                                   float eps, void* shadow_bf16, int cols, long long ld_shadow, void* shadow_lo,
                                   const float* normsq_fused, const float* reg_cross, const float* reg_wsq,
                                   float* wsq_out, void* stream) {
  return clip_adam_impl(w, g, m, v, n, normsq, clip_norm, weight_decay, lr_t, beta1, beta2, eps, shadow_bf16, cols,
                        ld_shadow, shadow_lo, normsq_fused, reg_cross, reg_wsq, wsq_out, stream);
}

extern "C" int evc_reg_cross(const float* logits, long long ld_logits, const void* dlogits, const void* dlogits_lo,
                             long long ld_dlogits, const float* bias, int B, int N, float* out, void* stream) {
  if (B <= 0 || N <= 0) return set_error(EVC_ERR_ARG, "reg_cross: empty problem");
  // vector path: every row of the three operands starts on a 16- (f32) / 8-byte (bf16) boundary
  const int vec = ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0 &&
                  ((reinterpret_cast<uintptr_t>(dlogits) | reinterpret_cast<uintptr_t>(dlogits_lo)) & 7) == 0 &&
                  ld_logits % 4 == 0 && ld_dlogits % 4 == 0;
  const int chunks = N >= 8192 ? 4 : (N >= 2048 ? 2 : 1);
  pdl_launch(reg_cross_kernel, dim3(B, chunks), dim3(256), 0, EVC_STREAM(stream), logits, ld_logits,
             static_cast<const __nv_bfloat16*>(dlogits), static_cast<const __nv_bfloat16*>(dlogits_lo), ld_dlogits, bias,
             N, vec, out);
  count_launch();
  return check_launch("reg_cross");
}

extern "C" int evc_batch_metrics(const float* P, const unsigned char* labels, int B, int V, int k, const int* idx,
                                 const float* val, const unsigned char* lab, const float* loss_rows,
                                 float* perr_rows, int* npos_rows, int* class_pos, double* acc, float* out,
                                 double* sums, void* stream) {
  if (B <= 0) return EVC_OK;
  if (k <= 0) return set_error(EVC_ERR_ARG, "batch_metrics: k must be a positive integer");
  if (!P || !labels || !idx || !val || !lab || !perr_rows || !npos_rows || !acc || !out)
    return set_error(EVC_ERR_ARG, "batch_metrics: null argument");
  pdl_launch(video_perr_kernel, dim3(B), dim3(256), static_cast<size_t>(V) * sizeof(int), EVC_STREAM(stream), P, labels,
             V, perr_rows, npos_rows, class_pos);
  count_launch();
  const int n = B * k;
  pdl_launch(gap_rank_kernel, dim3((n + 127) / 128), dim3(128), 0, EVC_STREAM(stream), idx, val, lab, n, k, acc);
  count_launch();
  pdl_launch(batch_metrics_finalize_kernel, dim3(1), dim3(256), 0, EVC_STREAM(stream), lab, B, k, perr_rows, npos_rows,
             loss_rows, acc, out, sums);
  count_launch();
  return check_launch("batch_metrics");
}

extern "C" int evc_topk(const float* P, int B, int V, int k, const unsigned char* labels, int* idx_out,
                        float* val_out, unsigned char* lab_out, void* stream) {
  if (k <= 0) return set_error(EVC_ERR_ARG, "topk: k must be a positive integer");  // eval_util.py:103-104
  if (k > V) return set_error(EVC_ERR_ARG, "topk: k > num_classes (clamp k = min(k, V) on the host)");
  const size_t smem = (static_cast<size_t>(V) + 32) * sizeof(unsigned long long);
  if (smem > 200 * 1024) return set_error(EVC_ERR_UNSUPPORTED, "topk: num_classes too large for one block");
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(topk_kernel), 200 * 1024)) return rc;
  pdl_launch(topk_kernel, dim3(B), dim3(256), smem, EVC_STREAM(stream), P, V, k, labels, idx_out, val_out, lab_out);
  count_launch();
  return check_launch("topk");
}
