// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] (f32 in TMEM) = A[M,K] * B[K,N]      bf16 operands, f32 accumulation
//
// * operands reach shared memory by TMA (SWIZZLE_128B boxes), the MMA is issued by one
//   elected thread (tcgen05.mma cta_group::1, 128 x BN x 16), accumulators live in TMEM
//   (double buffered so that the epilogue of tile i overlaps the main loop of tile i+1);
// * either operand may be K-major or MN-major in global memory, so the reference's weight
//   layout (kernel [in+H, 4H], gates [4H, V(M+1)], ...) is used as it is for forward,
//   dgrad and wgrad without transposed copies;
// * A may be the K-concatenation of two tensors ([x_t | h_{t-1}] of BasicLSTMCell);
// * epilogues: plain store (+bias, f32/bf16, optional split-K atomics), the BasicLSTM cell
//   forward (gate non-linearities, dynamic_rnn length mask, state update) and its backward
//   twin (gate gradients from the recurrent dgrad).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lanes 32*(warp%4) .. +31).
#pragma once
#include "evc_ptx.cuh"

namespace evc {

enum : int { EPI_STORE = 0, EPI_LSTM_FWD = 1, EPI_LSTM_BWD = 2 };

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmArgs {
  int M, N;            // output extent (rows, columns) used for masking
  int tiles_m, tiles_n, split_k;
  int kb_total;        // number of 64-wide k blocks (A1 part + A2 part)
  int kb_a1;           // k blocks taken from tensor map A1; the remainder comes from A2
  int kb_per_split;
  // ---- EPI_STORE
  void* C;
  long long ldc;
  int c_bf16;          // 0: float, 1: bf16
  int atomic_add;      // 1: red.add.f32 into C (split-K or accumulate)
  const float* bias;   // [N] or null
  // ---- EPI_LSTM_FWD / BWD (row r, hidden unit u; H = N/4 for fwd, N for bwd)
  int H;
  int t;                        // time step, row is live iff t < seq_len[r]
  const int* seq_len;           // [M]
  const float* c_prev;          // [M,H] or null (zero state)
  const __nv_bfloat16* h_prev;  // [M,H] or null
  float* c_out;                 // [M,H]
  __nv_bfloat16* h_out;         // [M,H]
  __nv_bfloat16* gates;         // [M,4H] post-activation i,j,f,o (fwd: out or null, bwd: in)
  // bwd only
  const float* dh_ext; long long ld_dh_ext;   // gradient arriving from above at step t (or null)
  const float* dh_pass_in; long long ld_dh_pass_in;
  const float* dc_in; long long ld_dc_in;
  float* dh_pass_out;           // [M,H]
  float* dc_out;                // [M,H]
  __nv_bfloat16* dz_out;        // [M,4H]
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int ACC_STAGES = 2;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
__device__ __forceinline__ void unpack8_bf16(const uint4& r, float* v) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(p[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load16_f32(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 f = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}
__device__ __forceinline__ void store16_f32(float* p, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* p, float* v) {
  uint4 a = *reinterpret_cast<const uint4*>(p);
  uint4 b = *reinterpret_cast<const uint4*>(p + 8);
  unpack8_bf16(a, v);
  unpack8_bf16(b, v + 8);
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* p, const float* v) {
  *reinterpret_cast<uint4*>(p) = pack8_bf16(v);
  *reinterpret_cast<uint4*>(p + 8) = pack8_bf16(v + 8);
}

template <int A_MN, int B_MN, int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + Cfg::ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + Cfg::ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_work = args.tiles_m * args.tiles_n * args.split_k;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int m_blk = w % args.tiles_m;
        const int n_blk = (w / args.tiles_m) % args.tiles_n;
        const int ks = w / (args.tiles_m * args.tiles_n);
        const int kb0 = ks * args.kb_per_split;
        const int kb1 = min(args.kb_total, kb0 + args.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const bool first = kb < args.kb_a1;
          const CUtensorMap* ta = first ? &tmA1 : &tmA2;
          const int ka = (first ? kb : kb - args.kb_a1) * BK;
          if (A_MN) {
            tma_load_2d(sa, ta, &full_bar[stage], m_blk * BM, ka);
            tma_load_2d(sa + 8192, ta, &full_bar[stage], m_blk * BM + 64, ka);
          } else {
            tma_load_2d(sa, ta, &full_bar[stage], ka, m_blk * BM);
          }
          const int kbk = kb * BK;
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) {
              int n = (EPI == EPI_LSTM_FWD) ? (i * args.H + n_blk * 64) : (n_blk * BN + i * 64);
              tma_load_2d(sb + i * 8192, &tmB, &full_bar[stage], n, kbk);
            }
          } else {
            tma_load_2d(sb, &tmB, &full_bar[stage], kbk, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int ks = w / (args.tiles_m * args.tiles_n);
        const int kb0 = ks * args.kb_per_split;
        const int kb1 = min(args.kb_total, kb0 + args.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb1 > kb0) umma_commit(&tfull_bar[as]);
        else mbar_arrive(&tfull_bar[as]);
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int m_blk = w % args.tiles_m;
      const int n_blk = (w / args.tiles_m) % args.tiles_n;
      const int ks = w / (args.tiles_m * args.tiles_n);
      const bool has_acc = min(args.kb_total, (ks + 1) * args.kb_per_split) > ks * args.kb_per_split;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < args.M;

      if constexpr (EPI == EPI_STORE) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c0;
          if (row_ok && col0 < args.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (args.bias != nullptr && ks == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < args.N) v[j] += __ldg(args.bias + col0 + j);
            }
            const bool full = (col0 + 32 <= args.N);
            if (args.atomic_add) {
              float* cp = reinterpret_cast<float*>(args.C) + static_cast<long long>(row) * args.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || col0 + j < args.N) atomicAdd(cp + j, v[j]);
            } else if (args.c_bf16) {
              __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(args.C) + static_cast<long long>(row) * args.ldc + col0;
              if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
                store16_bf16(cp, v);
                store16_bf16(cp + 16, v + 16);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < args.N) cp[j] = __float2bfloat16(v[j]);
              }
            } else {
              float* cp = reinterpret_cast<float*>(args.C) + static_cast<long long>(row) * args.ldc + col0;
              if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
                store16_f32(cp, v);
                store16_f32(cp + 16, v + 16);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < args.N) cp[j] = v[j];
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_LSTM_FWD) {
        // accumulator columns: [g*64 + u], gate g in (i, j, f, o), unit u of this tile
        const int H = args.H;
        const bool live = row_ok && (args.t < __ldg(args.seq_len + (row_ok ? row : 0)));
#pragma unroll 1
        for (int cu = 0; cu < 64; cu += 16) {
          uint32_t ri[16], rj[16], rf[16], ro[16];
          tmem_ld16(taddr + 0 * 64 + cu, ri);
          tmem_ld16(taddr + 1 * 64 + cu, rj);
          tmem_ld16(taddr + 2 * 64 + cu, rf);
          tmem_ld16(taddr + 3 * 64 + cu, ro);
          tmem_ld_wait();
          const int u0 = n_blk * 64 + cu;
          if (row_ok) {
            const long long off = static_cast<long long>(row) * H + u0;
            float cp[16];
            if (args.c_prev != nullptr) load16_f32(args.c_prev + off, cp);
            else {
#pragma unroll
              for (int j = 0; j < 16; ++j) cp[j] = 0.f;
            }
            if (live) {
              float gi[16], gj[16], gf[16], go[16], cn[16], hn[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float zi = __uint_as_float(ri[j]) + __ldg(args.bias + 0 * H + u0 + j);
                const float zj = __uint_as_float(rj[j]) + __ldg(args.bias + 1 * H + u0 + j);
                const float zf = __uint_as_float(rf[j]) + __ldg(args.bias + 2 * H + u0 + j);
                const float zo = __uint_as_float(ro[j]) + __ldg(args.bias + 3 * H + u0 + j);
                gi[j] = sigmoid_f(zi);
                gj[j] = tanh_f(zj);
                gf[j] = sigmoid_f(zf + 1.0f);  // forget_bias = 1.0 added at use
                go[j] = sigmoid_f(zo);
                cn[j] = cp[j] * gf[j] + gi[j] * gj[j];
                hn[j] = tanh_f(cn[j]) * go[j];
              }
              store16_f32(args.c_out + off, cn);
              store16_bf16(args.h_out + off, hn);
              if (args.gates != nullptr) {
                __nv_bfloat16* gp = args.gates + static_cast<long long>(row) * 4 * H + u0;
                store16_bf16(gp + 0 * H, gi);
                store16_bf16(gp + 1 * H, gj);
                store16_bf16(gp + 2 * H, gf);
                store16_bf16(gp + 3 * H, go);
              }
            } else {
              // dynamic_rnn: rows past their sequence_length keep their state
              store16_f32(args.c_out + off, cp);
              uint4 a = make_uint4(0, 0, 0, 0), b = a;
              if (args.h_prev != nullptr) {
                a = *reinterpret_cast<const uint4*>(args.h_prev + off);
                b = *reinterpret_cast<const uint4*>(args.h_prev + off + 8);
              }
              *reinterpret_cast<uint4*>(args.h_out + off) = a;
              *reinterpret_cast<uint4*>(args.h_out + off + 8) = b;
            }
          }
        }
      } else {  // EPI_LSTM_BWD : accumulator = dz_{t+1} * Wh^T, columns = hidden units of this tile
        const int H = args.H;
        const int len = row_ok ? __ldg(args.seq_len + row) : 0;
        const bool live = row_ok && (args.t < len);
        const bool had_pass = (args.t + 1 >= len);  // row was masked at step t+1 (or t is the last step)
#pragma unroll 1
        for (int cu = 0; cu < BN; cu += 16) {
          uint32_t racc[16];
          if (has_acc) {
            tmem_ld16(taddr + cu, racc);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) racc[j] = 0u;
          }
          const int u0 = n_blk * BN + cu;
          if (row_ok && u0 < H) {
            const long long off = static_cast<long long>(row) * H + u0;
            float dh[16], dc[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) dh[j] = __uint_as_float(racc[j]);
            if (args.dh_ext != nullptr) {
              float e[16];
              load16_f32(args.dh_ext + static_cast<long long>(row) * args.ld_dh_ext + u0, e);
#pragma unroll
              for (int j = 0; j < 16; ++j) dh[j] += e[j];
            }
            if (had_pass && args.dh_pass_in != nullptr) {
              float e[16];
              load16_f32(args.dh_pass_in + static_cast<long long>(row) * args.ld_dh_pass_in + u0, e);
#pragma unroll
              for (int j = 0; j < 16; ++j) dh[j] += e[j];
            }
            if (args.dc_in != nullptr) load16_f32(args.dc_in + static_cast<long long>(row) * args.ld_dc_in + u0, dc);
            else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dc[j] = 0.f;
            }
            __nv_bfloat16* zp = args.dz_out + static_cast<long long>(row) * 4 * H + u0;
            if (live) {
              float gi[16], gj[16], gf[16], go[16], cp[16];
              const __nv_bfloat16* gp = args.gates + static_cast<long long>(row) * 4 * H + u0;
              load16_bf16(gp + 0 * H, gi);
              load16_bf16(gp + 1 * H, gj);
              load16_bf16(gp + 2 * H, gf);
              load16_bf16(gp + 3 * H, go);
              if (args.c_prev != nullptr) load16_f32(args.c_prev + off, cp);
              else {
#pragma unroll
                for (int j = 0; j < 16; ++j) cp[j] = 0.f;
              }
              float di[16], dj[16], df[16], dout[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float cn = cp[j] * gf[j] + gi[j] * gj[j];
                const float tc = tanh_f(cn);
                dout[j] = dh[j] * tc * go[j] * (1.f - go[j]);
                const float dcn = dc[j] + dh[j] * go[j] * (1.f - tc * tc);
                di[j] = dcn * gj[j] * gi[j] * (1.f - gi[j]);
                dj[j] = dcn * gi[j] * (1.f - gj[j] * gj[j]);
                df[j] = dcn * cp[j] * gf[j] * (1.f - gf[j]);
                dc[j] = dcn * gf[j];
              }
              store16_bf16(zp + 0 * H, di);
              store16_bf16(zp + 1 * H, dj);
              store16_bf16(zp + 2 * H, df);
              store16_bf16(zp + 3 * H, dout);
              store16_f32(args.dc_out + off, dc);
            } else {
              const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                *reinterpret_cast<uint4*>(zp + g * H) = z;
                *reinterpret_cast<uint4*>(zp + g * H + 8) = z;
              }
              store16_f32(args.dc_out + off, dc);
              store16_f32(args.dh_pass_out + off, dh);
            }
          }
        }
      }
      // release the accumulator stage to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace evc
