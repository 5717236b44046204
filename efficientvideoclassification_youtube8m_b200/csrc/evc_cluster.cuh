// One BasicLSTM step for the small-row regime as ONE kernel: split-K over a thread-block cluster with the partial
// sums exchanged through distributed shared memory, then the cell update in the same kernel.
//
// The per-step path of round 1 for <= 1024 rows was two launches: a split-K GEMM that writes S f32 slabs of the gate
// pre-activations to memory (rows x 4H x 4 B each) and a cell kernel that reads them back -- ~15 + 8 us at 256 rows,
// both latency-bound.  Here the S CTAs that share an output tile form a cluster:
//   1. each CTA runs the TMA -> tcgen05.mma pipeline over ITS share of the k blocks into a [128 x 256] f32 accumulator
//      in TMEM (tile = 128 rows x 4 gates x 64 units);
//   2. reduce-scatter through DSMEM: CTA j of the cluster owns units [j*64/S, (j+1)*64/S) of the tile; every CTA
//      copies the accumulator columns of the other CTAs' units from TMEM straight into THEIR shared memory
//      (st.shared::cluster, 16-byte stores, XOR-swizzled so that a warp's 32 rows spread over the banks);
//   3. after a cluster barrier every CTA sums the S partials of its units in rank order (bit-reproducible), adds the
//      bias, applies the gate non-linearities, the c/h update and the dynamic_rnn length mask, and stores h / c /
//      gates for 128 rows x 64/S units (thread = row, as in the resident-weights kernel).
// No slab goes through memory, one launch per step, and only intra-cluster barriers: any number of these grids may be
// in flight (unlike the persistent kernel of evc_rec.cuh there is no co-residency requirement across clusters).
// The receive buffers alias the operand pipeline stages, which are dead once the cluster has passed barrier #1.
#pragma once
#include "evc_gemm.cuh"

namespace evc {

constexpr int CL_BN = 256;                       // 4 gates x 64 units
constexpr int CL_STAGES = 4;
constexpr int CL_STAGE_BYTES = BM * BK * 2 + CL_BN * BK * 2;   // 16 + 32 KB
constexpr int CL_THREADS = 192;
constexpr int CL_SMEM_BYTES = CL_STAGES * CL_STAGE_BYTES + 256;

struct ClusterStepArgs {
  int rows, H, t;
  int tiles_m, tiles_n;          // tiles_n = H / 64
  int kb_a1, kb_total;           // k blocks from map A1 (x_t), total (x_t | h_{t-1})
  const float* bias;             // [4H]
  const int* seq_len;            // [rows]
  const float* c_prev;           // [rows, H] or null (t = 0)
  const __nv_bfloat16* h_prev;   // [rows, H] or null
  float* c_out;                  // [rows, H]
  __nv_bfloat16* h_out;          // [rows, H]
  __nv_bfloat16* gates;          // [rows, 4H] or null
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// KS = cluster size = number of K splits (2, 4 or 8); UP = 64 / KS hidden units per CTA after the reduce-scatter
template <int KS>
__global__ void __launch_bounds__(CL_THREADS, 1)
lstm_cluster_step_fwd_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                             const __grid_constant__ CUtensorMap tmB, const ClusterStepArgs args) {
  constexpr int UP = 64 / KS;                      // units owned by a CTA
  constexpr int COLS = 4 * UP;                     // accumulator columns owned by a CTA (4 gates)
  constexpr int CHUNKS = COLS / 4;                 // 16-byte chunks per row of a receive slot
  constexpr int SLOT_BYTES = BM * COLS * 4;        // one peer's partial for my columns
  static_assert((KS - 1) * SLOT_BYTES <= CL_STAGES * CL_STAGE_BYTES, "receive buffers alias the pipeline stages");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CL_STAGES * CL_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + CL_STAGES;
  uint64_t* tfull_bar = empty_bar + CL_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ks = static_cast<int>(cluster_ctarank());
  const int tile = blockIdx.x / KS;
  const int m_blk = tile % args.tiles_m;
  const int n_blk = tile / args.tiles_m;
  const int H = args.H;

  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < CL_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();                              // also: every CTA of the cluster is running before any DSMEM access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  // this CTA's share of the k blocks (balanced; may be empty when there are fewer k blocks than CTAs)
  const int kb0 = args.kb_total * ks / KS;
  const int kb1 = args.kb_total * (ks + 1) / KS;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * CL_STAGE_BYTES;
        uint8_t* sb = sa + BM * BK * 2;
        mbar_arrive_expect_tx(&full_bar[stage], CL_STAGE_BYTES);
        if (kb < args.kb_a1) tma_load_2d(sa, &tmA1, &full_bar[stage], kb * BK, m_blk * BM);
        else tma_load_2d(sa, &tmA2, &full_bar[stage], (kb - args.kb_a1) * BK, m_blk * BM);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_2d(sb + g * 8192, &tmB, &full_bar[stage], g * H + n_blk * 64, kb * BK);
        if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, CL_BN, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * CL_STAGE_BYTES);
        const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 2048, 8192, 1024), idesc,
                    (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == CL_STAGES) { stage = 0; phase ^= 1; }
      }
      if (kb1 > kb0) umma_commit(tfull_bar);       // arrives when every MMA of this CTA has completed
      else mbar_arrive(tfull_bar);
    }
  }
  // ---- every thread: the MMAs of this CTA are complete (its operand stages are dead) ...
  __syncwarp();
  mbar_wait(tfull_bar, 0);
  tc_fence_after();
  // ... and so are those of the whole cluster: peers may now write into this CTA's (aliased) receive buffers
  cluster_sync_all();
  const bool has_acc = kb1 > kb0;

  if (warp >= 2) {
    // ===================================================== reduce-scatter through DSMEM: thread = row of the tile
    const int q = warp & 3;
    const int rl = q * 32 + lane;                                   // row inside the tile = TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t recv_local = smem_u32(smem);
#pragma unroll 1
    for (int d = 1; d < KS; ++d) {
      const int dst = (ks + d) % KS;                                // spread the traffic: everybody starts elsewhere
      const int slot = ks < dst ? ks : ks - 1;                      // my slot in dst's buffer (sources in rank order)
      const uint32_t base = mapa_u32(recv_local, dst) + slot * SLOT_BYTES + rl * (COLS * 4);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t r[UP];
        if (has_acc) {
          if constexpr (UP == 32) tmem_ld32(taddr + g * 64 + dst * UP, r);
          else if constexpr (UP == 16) tmem_ld16(taddr + g * 64 + dst * UP, r);
          else tmem_ld8(taddr + g * 64 + dst * UP, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int v = 0; v < UP; ++v) r[v] = 0u;
        }
#pragma unroll
        for (int v = 0; v < UP / 4; ++v) {
          const int chunk = (g * (UP / 4) + v) ^ (rl & (CHUNKS - 1));
          st_cluster_v4(base + chunk * 16, __uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                        __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        }
      }
    }
  }
  cluster_sync_all();                               // barrier.cluster arrive.release / wait.acquire: partials visible

  if (warp >= 2) {
    // ===================================================== cell update for 128 rows x UP units
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int r = m_blk * BM + rl;
    const bool ok = r < args.rows;
    const int rr = ok ? r : args.rows - 1;
    const int u0 = n_blk * 64 + ks * UP;
    const long long off = static_cast<long long>(rr) * H + u0;
    float z[4][UP];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int v = 0; v < UP / 4; ++v) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + g * H + u0) + v);
        z[g][4 * v] = b.x; z[g][4 * v + 1] = b.y; z[g][4 * v + 2] = b.z; z[g][4 * v + 3] = b.w;
      }
    }
    float cp[UP];
    if (args.c_prev != nullptr) {
#pragma unroll
      for (int v = 0; v < UP / 4; ++v) {
        const float4 a = *reinterpret_cast<const float4*>(args.c_prev + off + 4 * v);
        cp[4 * v] = a.x; cp[4 * v + 1] = a.y; cp[4 * v + 2] = a.z; cp[4 * v + 3] = a.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < UP; ++v) cp[v] = 0.f;
    }
    const bool live = ok && (args.t < __ldg(args.seq_len + rr));
    // partial sums in rank order 0 .. KS-1 (the same order in every CTA: the result does not depend on the split)
    const float* recv = reinterpret_cast<const float*>(smem);
#pragma unroll
    for (int s = 0; s < KS; ++s) {
      if (s == ks) {
        if (has_acc) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t acc[UP];
            if constexpr (UP == 32) tmem_ld32(taddr + g * 64 + ks * UP, acc);
            else if constexpr (UP == 16) tmem_ld16(taddr + g * 64 + ks * UP, acc);
            else tmem_ld8(taddr + g * 64 + ks * UP, acc);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < UP; ++v) z[g][v] += __uint_as_float(acc[v]);
          }
        }
      } else {
        const int slot = s < ks ? s : s - 1;
        const float* p = recv + slot * (SLOT_BYTES / 4) + rl * COLS;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int v = 0; v < UP / 4; ++v) {
            const int chunk = (g * (UP / 4) + v) ^ (rl & (CHUNKS - 1));
            const float4 a = *reinterpret_cast<const float4*>(p + chunk * 4);
            z[g][4 * v] += a.x; z[g][4 * v + 1] += a.y; z[g][4 * v + 2] += a.z; z[g][4 * v + 3] += a.w;
          }
      }
    }
    if (ok) {
      float* c_out = args.c_out + off;
      __nv_bfloat16* h_out = args.h_out + off;
      if (!live) {
        // dynamic_rnn: rows past their sequence_length keep their state
#pragma unroll
        for (int v = 0; v < UP / 4; ++v)
          reinterpret_cast<float4*>(c_out)[v] = make_float4(cp[4 * v], cp[4 * v + 1], cp[4 * v + 2], cp[4 * v + 3]);
#pragma unroll
        for (int v = 0; v < UP / 8; ++v) {
          uint4 hp = make_uint4(0u, 0u, 0u, 0u);
          if (args.h_prev != nullptr) hp = reinterpret_cast<const uint4*>(args.h_prev + off)[v];
          reinterpret_cast<uint4*>(h_out)[v] = hp;
        }
      } else {
        uint32_t hq[UP / 2], gq[4][UP / 2];
        float cn[UP];
#pragma unroll
        for (int v = 0; v < UP; v += 2) {
          float hn[2], gi[2], gj[2], gf[2], go[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            gi[e] = sigmoid_f(z[0][v + e]);
            gj[e] = tanh_f(z[1][v + e]);
            gf[e] = sigmoid_f(z[2][v + e] + 1.0f);          // forget_bias = 1.0 added at use
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
        for (int v = 0; v < UP / 4; ++v)
          reinterpret_cast<float4*>(c_out)[v] = make_float4(cn[4 * v], cn[4 * v + 1], cn[4 * v + 2], cn[4 * v + 3]);
#pragma unroll
        for (int v = 0; v < UP / 8; ++v)
          reinterpret_cast<uint4*>(h_out)[v] = make_uint4(hq[4 * v], hq[4 * v + 1], hq[4 * v + 2], hq[4 * v + 3]);
        if (args.gates != nullptr) {
          __nv_bfloat16* gp = args.gates + static_cast<long long>(r) * 4 * H + u0;
#pragma unroll
          for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int v = 0; v < UP / 8; ++v)
              reinterpret_cast<uint4*>(gp + g * H)[v] = make_uint4(gq[g][4 * v], gq[g][4 * v + 1], gq[g][4 * v + 2], gq[g][4 * v + 3]);
        }
      }
    }
  }

  // (all remote stores into this CTA completed before barrier #2; nothing of a peer is touched after it)
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace evc
