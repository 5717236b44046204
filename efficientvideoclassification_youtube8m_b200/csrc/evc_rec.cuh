// Persistent multi-CTA BasicLSTM recurrence with the recurrent weights RESIDENT in shared memory
// (small-row regime: RNN_L2 of both models, the student's RNN_L1; frame_level_models.py:247-257,318-328).
//
// One launch runs all T steps of one cell.  The input half of BasicLSTMCell's matmul does not depend on the
// recurrence, so it is hoisted: Zx[t] = x_t * Wx + bias for ALL steps is one large GEMM (evc_gemm.cu), and only
//   z_t = Zx[t] + h_{t-1} * Wh        (Wh = rows Kx.. of the kernel, [H, 4H])
// is sequential.  Wh is split by hidden unit: CTA (g, s) owns the 4 gate columns of units [16 s, 16 s + 16), i.e. a
// [H x 64] slice (128 KB at H = 1024) that it loads ONCE by TMA and keeps in shared memory for all T steps; the
// H/16 slices x G row groups fill the SMs (128 CTAs at H = 1024).  Per step and 128-row tile a CTA streams only
// the h_{t-1} tile (bf16 [128 x H], from L2) through a 4-stage TMA ring into tcgen05.mma (128 x 64 x 16, f32
// accumulator in TMEM), and its 4 epilogue warps (thread = row) add Zx, apply the gate non-linearities, the
// c/h update and the dynamic_rnn length mask, and write h_t / c_t / gates for exactly their 16 units -- no f32
// pre-activation slab ever goes through memory.
//
// Step synchronisation is per 128-row tile, not grid wide: ready[t][m] counts the slices that have published
// h_t of tile m (release: fence.proxy.async + __threadfence + atomicAdd); a producer acquires ready[t-1][m] ==
// #slices before its TMA reads h_{t-1} of tile m.  A CTA that serves several tiles (rows > 128 * G) overlaps the
// epilogue of tile m with the MMAs of tile m+1 (one TMEM accumulator per tile).
//
// All CTAs must be co-resident (grid <= #SMs, one CTA per SM by shared memory); the host never lets two of
// these grids run concurrently (engine.py serialises them with an event).  Waits are bounded: a CTA that spins
// longer than ~4 s traps instead of hanging the device.
#pragma once
#include "evc_gemm.cuh"

namespace evc {

constexpr int REC_UNITS = 16;                 // hidden units per CTA -> 4 * 16 = 64 accumulator columns
constexpr int REC_BN = 4 * REC_UNITS;
constexpr int REC_STAGES = 4;                 // A ring: 4 x [128 rows x 64 k] bf16 = 64 KB
constexpr int REC_A_BYTES = BM * BK * 2;      // 16 KB
constexpr int REC_W_KB_BYTES = BK * REC_BN * 2;   // one 64-deep k block of the resident slice: 8 KB
constexpr int REC_MAX_TILES = 8;              // TMEM: 8 accumulators x 64 columns = 512
constexpr int REC_THREADS = 192;

struct RecArgs {
  int rows, H, T;
  int tiles_m;                   // 128-row tiles
  int groups;                    // CTAs per slice (row groups); gridDim.x = groups * (H / 16)
  int tiles_per_cta;             // ceil(tiles_m / groups) <= REC_MAX_TILES
  const float* zx;               // [T][rows][4H] f32: x_t * Wx + bias
  const int* seq_len;            // [rows]
  float* c_all;                  // [(T+1), rows, H]
  __nv_bfloat16* h_all;          // [(T+1), rows, H]
  __nv_bfloat16* gates_all;      // [T, rows, 4H] or null
  unsigned int* ready;           // [T][tiles_m] zero-initialised arrival counters
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

inline size_t rec_smem_bytes(int H) {
  return static_cast<size_t>(H / BK) * REC_W_KB_BYTES + REC_STAGES * REC_A_BYTES + 512;
}

// Wp[s][k][g*16 + u] = W[(Kx + k)][g*H + 16 s + u]: the slice of one CTA as a dense [H x 64] bf16 matrix whose
// rows are exactly one SWIZZLE_128B line (TMA box 64 x 64).
__global__ void lstm_pack_wh_kernel(const __nv_bfloat16* __restrict__ Wh, int H, __nv_bfloat16* __restrict__ Wp) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per 8 output elements (16 bytes): out index = ((s*H + k)*4 + g)*2 + half
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(H / REC_UNITS) * H * 8;
  if (i >= n) return;
  const int half = static_cast<int>(i & 1);
  const int g = static_cast<int>((i >> 1) & 3);
  const long long sk = i >> 3;
  const int k = static_cast<int>(sk % H);
  const int s = static_cast<int>(sk / H);
  const uint4 v = *reinterpret_cast<const uint4*>(Wh + static_cast<long long>(k) * 4 * H + g * H + s * REC_UNITS + half * 8);
  *reinterpret_cast<uint4*>(Wp + (sk * 4 + g) * REC_UNITS + half * 8) = v;
}

__global__ void __launch_bounds__(REC_THREADS, 1)
lstm_rec_resident_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW,
                             const RecArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int H = args.H;
  const int kbs = H / BK;                                   // k blocks per step
  uint8_t* w_smem = smem_raw;                               // kbs x 8 KB, resident
  uint8_t* a_smem = smem_raw + kbs * REC_W_KB_BYTES;        // ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(a_smem + REC_STAGES * REC_A_BYTES);
  uint64_t* empty_bar = full_bar + REC_STAGES;
  uint64_t* tfull_bar = empty_bar + REC_STAGES;             // [REC_MAX_TILES]
  uint64_t* tempty_bar = tfull_bar + REC_MAX_TILES;         // [REC_MAX_TILES]
  uint64_t* w_bar = tempty_bar + REC_MAX_TILES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_slices = H / REC_UNITS;
  const int slice = blockIdx.x % n_slices;
  const int group = blockIdx.x / n_slices;
  const int tile0 = group * args.tiles_per_cta;
  const int ntiles = max(0, min(args.tiles_per_cta, args.tiles_m - tile0));

  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < REC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < REC_MAX_TILES; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  const long long RH = static_cast<long long>(args.rows) * H;
  const long long t0_clock = clock64();
  constexpr long long kSpinLimit = 8000000000LL;            // ~4 s at 2 GHz: a lost peer must not hang the device

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && ntiles > 0) {
      // the resident slice, once
      mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(kbs * REC_W_KB_BYTES));
      for (int kb = 0; kb < kbs; ++kb)
        tma_load_2d(w_smem + kb * REC_W_KB_BYTES, &tmW, w_bar, 0, slice * H + kb * BK);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 1; t < args.T; ++t) {                    // step 0 has no recurrent term (h_{-1} = 0)
        for (int i = 0; i < ntiles; ++i) {
          const int m = tile0 + i;
          const unsigned int* flag = args.ready + static_cast<long long>(t - 1) * args.tiles_m + m;
          while (ld_acquire_u32(flag) < static_cast<unsigned int>(n_slices)) {
            __nanosleep(20);
            if (clock64() - t0_clock > kSpinLimit) __trap();
          }
          fence_proxy_async_global();                       // peers' generic-proxy stores of h_{t-1} -> TMA reads
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], REC_A_BYTES);
            tma_load_3d(a_smem + stage * REC_A_BYTES, &tmH, &full_bar[stage], kb * BK, m * BM, t);
            if (++stage == REC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, REC_BN, 0, 1);
      mbar_wait(w_bar, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 1; t < args.T; ++t) {
        const uint32_t aphase = static_cast<uint32_t>(t - 1) & 1;   // accumulator i is used once per step
        for (int i = 0; i < ntiles; ++i) {
          mbar_wait(&tempty_bar[i], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + i * REC_BN;
          for (int kb = 0; kb < kbs; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(a_smem + stage * REC_A_BYTES);
            const uint32_t sb = smem_u32(w_smem + kb * REC_W_KB_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 2048, 8192, 1024), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == REC_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[i]);
        }
      }
    }
  } else if (ntiles > 0) {
    // ===================================================== epilogue warps: thread = row of the tile, 16 units
    const int q = warp & 3;
    const int u0 = slice * REC_UNITS;
    for (int t = 0; t < args.T; ++t) {
      const uint32_t aphase = static_cast<uint32_t>(t - 1) & 1;
      for (int i = 0; i < ntiles; ++i) {
        const int m = tile0 + i;
        const int r = m * BM + q * 32 + lane;
        const bool ok = r < args.rows;
        const int rr = ok ? r : args.rows - 1;
        const long long off = static_cast<long long>(rr) * H + u0;
        // independent global loads first: the hoisted input projection, the previous cell state, the length
        float z[4][REC_UNITS];
        {
          const float* zp = args.zx + (static_cast<long long>(t) * args.rows + rr) * 4 * H + u0;
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int v = 0; v < REC_UNITS / 4; ++v) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(zp + g * H) + v);
              z[g][4 * v] = a.x; z[g][4 * v + 1] = a.y; z[g][4 * v + 2] = a.z; z[g][4 * v + 3] = a.w;
            }
        }
        float cp[REC_UNITS];
        if (t > 0) {
          const float4* c4 = reinterpret_cast<const float4*>(args.c_all + t * RH + off);
#pragma unroll
          for (int v = 0; v < REC_UNITS / 4; ++v) {
            const float4 a = __ldcg(c4 + v);
            cp[4 * v] = a.x; cp[4 * v + 1] = a.y; cp[4 * v + 2] = a.z; cp[4 * v + 3] = a.w;
          }
        } else {
#pragma unroll
          for (int v = 0; v < REC_UNITS; ++v) cp[v] = 0.f;
        }
        const bool live = ok && (t < __ldg(args.seq_len + rr));
        if (t > 0) {
          mbar_wait(&tfull_bar[i], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + i * REC_BN;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t acc[16];
            tmem_ld16(taddr + g * REC_UNITS, acc);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < REC_UNITS; ++v) z[g][v] += __uint_as_float(acc[v]);
          }
          // the accumulator is in registers: the MMA warp may start this tile's next step
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[i]);
        }
        float* c_out = args.c_all + (t + 1) * RH + off;
        __nv_bfloat16* h_out = args.h_all + (t + 1) * RH + off;
        if (ok) {
          if (!live) {
            // dynamic_rnn: rows past their sequence_length keep their state
            uint4 hp[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
            if (t > 0) {
              const uint4* h4 = reinterpret_cast<const uint4*>(args.h_all + t * RH + off);
              hp[0] = __ldcg(h4);
              hp[1] = __ldcg(h4 + 1);
            }
#pragma unroll
            for (int v = 0; v < REC_UNITS / 4; ++v)
              reinterpret_cast<float4*>(c_out)[v] = make_float4(cp[4 * v], cp[4 * v + 1], cp[4 * v + 2], cp[4 * v + 3]);
            reinterpret_cast<uint4*>(h_out)[0] = hp[0];
            reinterpret_cast<uint4*>(h_out)[1] = hp[1];
          } else {
            uint32_t hq[REC_UNITS / 2], gq[4][REC_UNITS / 2];
            float cn[REC_UNITS];
#pragma unroll
            for (int v = 0; v < REC_UNITS; v += 2) {
              float hn[2], gi[2], gj[2], gf[2], go[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                gi[e] = sigmoid_f(z[0][v + e]);
                gj[e] = tanh_f(z[1][v + e]);
                gf[e] = sigmoid_f(z[2][v + e] + 1.0f);      // forget_bias = 1.0 added at use
                go[e] = sigmoid_f(z[3][v + e]);
                cn[v + e] = cp[v + e] * gf[e] + gi[e] * gj[e];
                hn[e] = tanh_f(cn[v + e]) * go[e];
              }
              __nv_bfloat162 b;
              b = __floats2bfloat162_rn(hn[0], hn[1]); hq[v / 2] = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(gi[0], gi[1]); gq[0][v / 2] = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(gj[0], gj[1]); gq[1][v / 2] = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(gf[0], gf[1]); gq[2][v / 2] = *reinterpret_cast<uint32_t*>(&b);
              b = __floats2bfloat162_rn(go[0], go[1]); gq[3][v / 2] = *reinterpret_cast<uint32_t*>(&b);
            }
#pragma unroll
            for (int v = 0; v < REC_UNITS / 4; ++v)
              reinterpret_cast<float4*>(c_out)[v] = make_float4(cn[4 * v], cn[4 * v + 1], cn[4 * v + 2], cn[4 * v + 3]);
            reinterpret_cast<uint4*>(h_out)[0] = make_uint4(hq[0], hq[1], hq[2], hq[3]);
            reinterpret_cast<uint4*>(h_out)[1] = make_uint4(hq[4], hq[5], hq[6], hq[7]);
            if (args.gates_all != nullptr) {
              __nv_bfloat16* gp = args.gates_all + (static_cast<long long>(t) * args.rows + r) * 4 * H + u0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                reinterpret_cast<uint4*>(gp + g * H)[0] = make_uint4(gq[g][0], gq[g][1], gq[g][2], gq[g][3]);
                reinterpret_cast<uint4*>(gp + g * H)[1] = make_uint4(gq[g][4], gq[g][5], gq[g][6], gq[g][7]);
              }
            }
          }
        }
        // publish h_t of this tile: the stores of all 128 epilogue threads happen-before the barrier, one thread
        // then makes them visible device-wide (cumulative fence) and increments the tile's counter (release)
        asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 epilogue warps
        if (warp == 2 && lane == 0 && t + 1 < args.T) {
          fence_proxy_async_global();
          __threadfence();
          atomicAdd(args.ready + static_cast<long long>(t) * args.tiles_m + m, 1u);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace evc
