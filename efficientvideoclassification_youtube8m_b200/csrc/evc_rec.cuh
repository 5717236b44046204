// Persistent multi-step kernel for the small-row BasicLSTM recurrences (RNN_L2 of both models and the
// student's RNN_L1: <= 1024 rows).  One launch runs all T steps of one cell:
//
//   for t in 0..T-1:
//     GEMM phase   every CTA owns one (128 x 256 tile, K split) work item of [x_t | h_{t-1}] * W:
//                  TMA (3-D maps over the step index) -> tcgen05.mma -> f32 partial slab in HBM/L2
//     grid barrier
//     cell phase   all threads of all CTAs: z = sum of slabs + bias, gates, c/h update, length mask;
//                  h_t (bf16) is the A operand of step t+1, so the writers fence generic -> async proxy
//     grid barrier
//
// Per step this replaces two launches (split-K GEMM + cell kernel, ~27 us) by two grid barriers
// (~1.5 us each); the CTAs stay resident, barriers/TMEM are set up once.  All CTAs must be
// co-resident: grid <= #SMs with one CTA per SM (host side checks).
#pragma once
#include "evc_gemm.cuh"

namespace evc {

struct RecArgs {
  int rows, H, T;
  int tiles_m, tiles_n, S;       // work items = tiles_m * tiles_n * S = gridDim.x
  int kb_x, kb_h;                // 64-deep k blocks of the input part and of the recurrent part
  const float* bias;             // [4H]
  const int* seq_len;            // [rows]
  float* c_all;                  // [(T+1), rows, H]
  __nv_bfloat16* h_all;          // [(T+1), rows, H]
  __nv_bfloat16* gates_all;      // [T, rows, 4H] or null
  float* slabs;                  // [S][rows][4H] f32 partial pre-activations
  unsigned int* barrier;         // zero-initialised arrival counter
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
lstm_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH,
                    const __grid_constant__ CUtensorMap tmW, const RecArgs args) {
  constexpr int BN = 256;
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  float* epi_stage_base = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + EPI_WARPS * Cfg::EPI_STAGE_WORDS * 4);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  const int H = args.H, rows = args.rows;
  const int w = blockIdx.x;
  const int m_blk = w % args.tiles_m;
  const int n_blk = (w / args.tiles_m) % args.tiles_n;
  const int ks = w / (args.tiles_m * args.tiles_n);
  const long long RH = static_cast<long long>(rows) * H;
  const long long slab_stride = static_cast<long long>(rows) * 4 * H;
  const unsigned int G = gridDim.x;
  unsigned int bar_target = 0;

  int stage = 0;          // smem pipeline position (the producer and the MMA thread advance their own copies in step)
  uint32_t phase = 0;
  int it = 0;             // accumulator tiles this CTA has produced so far (TMEM stage / phase bookkeeping)

  for (int t = 0; t < args.T; ++t) {
    const int kb_total = args.kb_x + (t == 0 ? 0 : args.kb_h);   // h_{-1} = 0: skip the recurrent half at t = 0
    const int kb_per = (kb_total + args.S - 1) / args.S;
    const int splits = (kb_total + kb_per - 1) / kb_per;
    const int kb0 = ks * kb_per;
    const int kb1 = min(kb_total, kb0 + kb_per);
    const int nkb = max(kb1 - kb0, 0);
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;

    // ------------------------------------------------------------ GEMM phase
    if (warp == 0) {
      if (lane == 0) {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (kb < args.kb_x) tma_load_3d(sa, &tmX, &full_bar[stage], kb * BK, m_blk * BM, t);
          else tma_load_3d(sa, &tmH, &full_bar[stage], (kb - args.kb_x) * BK, m_blk * BM, t);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            tma_load_2d(sb + i * 8192, &tmW, &full_bar[stage], i * H + n_blk * 64, kb * BK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && nkb > 0) {
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 2048, 8192, 1024), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    } else if (nkb > 0) {
      // epilogue warps: accumulator [g*64 + u] -> slab[ks][row][g*H + n_blk*64 + u] (f32, coalesced through staging)
      const int q = warp & 3;
      float* st_f = epi_stage_base + (warp - 2) * Cfg::EPI_STAGE_WORDS;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const int row0 = m_blk * BM + q * 32;
      const int nrows = rows - row0;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (nrows > 0) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          const int g = c0 >> 6, u = c0 & 63;
          float* dst = args.slabs + ks * slab_stride + static_cast<long long>(row0) * 4 * H + g * H + n_blk * 64 + u;
          stage_put_f32(st_f, lane, v);
          __syncwarp();
          flush_f32(st_f, dst, 4LL * H, nrows, lane);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (nkb > 0) ++it;

    bar_target += G;
    grid_barrier(args.barrier, bar_target);

    // ------------------------------------------------------------ cell phase (all threads of all CTAs)
    {
      const int hq = H >> 2;
      const long long nquad = static_cast<long long>(rows) * hq;
      const float* c_prev = (t == 0) ? nullptr : args.c_all + t * RH;
      const __nv_bfloat16* h_prev = (t == 0) ? nullptr : args.h_all + t * RH;
      float* c_out = args.c_all + (t + 1) * RH;
      __nv_bfloat16* h_out = args.h_all + (t + 1) * RH;
      __nv_bfloat16* gates = args.gates_all ? args.gates_all + t * RH * 4 : nullptr;
      for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < nquad;
           idx += static_cast<long long>(G) * blockDim.x) {
        const int r = static_cast<int>(idx / hq);
        const int u = static_cast<int>(idx % hq) * 4;
        const long long off = static_cast<long long>(r) * H + u;
        float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c_prev != nullptr) cp = __ldcg(reinterpret_cast<const float4*>(c_prev + off));
        if (t >= __ldg(args.seq_len + r)) {               // dynamic_rnn: keep the state past sequence_length
          *reinterpret_cast<float4*>(c_out + off) = cp;
          uint2 hp = make_uint2(0u, 0u);
          if (h_prev != nullptr) hp = __ldcg(reinterpret_cast<const uint2*>(h_prev + off));
          *reinterpret_cast<uint2*>(h_out + off) = hp;
          continue;
        }
        float z[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + g * H + u));
          z[g][0] = b.x; z[g][1] = b.y; z[g][2] = b.z; z[g][3] = b.w;
        }
        for (int s = 0; s < splits; ++s) {
          const float* zp = args.slabs + s * slab_stride + static_cast<long long>(r) * 4 * H + u;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(zp + g * H));
            z[g][0] += a.x; z[g][1] += a.y; z[g][2] += a.z; z[g][3] += a.w;
          }
        }
        const float cpa[4] = {cp.x, cp.y, cp.z, cp.w};
        float gi[4], gj[4], gf[4], go[4], cn[4], hn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          gi[k] = sigmoid_f(z[0][k]);
          gj[k] = tanh_f(z[1][k]);
          gf[k] = sigmoid_f(z[2][k] + 1.0f);
          go[k] = sigmoid_f(z[3][k]);
          cn[k] = cpa[k] * gf[k] + gi[k] * gj[k];
          hn[k] = tanh_f(cn[k]) * go[k];
        }
        *reinterpret_cast<float4*>(c_out + off) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        {
          __nv_bfloat162 lo = __floats2bfloat162_rn(hn[0], hn[1]), hi = __floats2bfloat162_rn(hn[2], hn[3]);
          *reinterpret_cast<uint2*>(h_out + off) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
        if (gates != nullptr) {
          __nv_bfloat16* gp = gates + static_cast<long long>(r) * 4 * H + u;
          const float* gs[4] = {gi, gj, gf, go};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(gs[g][0], gs[g][1]), hi = __floats2bfloat162_rn(gs[g][2], gs[g][3]);
            *reinterpret_cast<uint2*>(gp + g * H) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
          }
        }
      }
      fence_proxy_async_global();   // h_t is read by the next step's TMA loads (async proxy)
    }
    bar_target += G;
    grid_barrier(args.barrier, bar_target);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace evc
