// Host side of the tcgen05 GEMM family: tensor-map construction, launch, and the C-ABI entry
// points evc_gemm_bf16 / evc_lstm_seq_fwd / evc_lstm_seq_bwd declared in include/evc.h.
#include "evc_gemm.cuh"
#include "evc_rec.cuh"
#include "evc_cluster.cuh"
#include "evc_host.h"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

namespace evc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor maps are pure functions of (pointer, extents, pitch, box, element size).  The step replays
// the same few hundred operand descriptors every iteration, so they are cached: encoding them anew
// cost ~40 % of the host-side launch time of a training step.
struct TmapKey {
  const void* ptr;
  uint64_t inner, outer, pitch;
  uint32_t box_inner, box_outer, elem;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && pitch == o.pitch && box_inner == o.box_inner &&
           box_outer == o.box_outer && elem == o.elem;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner * 0xC2B2AE3D27D4EB4Full) ^ (k.outer << 21) ^ (k.pitch << 7) ^ (static_cast<uint64_t>(k.box_inner) << 50) ^
         (static_cast<uint64_t>(k.box_outer) << 40) ^ k.elem;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mutex;

// 2-D tensor [outer][inner] (bf16 or f32) with row pitch `pitch_elems`; box = box_inner x box_outer, 128B swizzle.
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                     uint32_t box_inner, uint32_t box_outer, int elem_bytes = 2) {
  const TmapKey key{ptr, inner, outer, pitch_elems, box_inner, box_outer, static_cast<uint32_t>(elem_bytes)};
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *m = it->second;
      return EVC_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (pitch_elems * elem_bytes) % 16 != 0)
    return set_error(EVC_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();   // callers that churn buffers: bound the memory
    g_tmap_cache.emplace(key, *m);
  }
  return EVC_OK;
}

// logical A[M,K]: a_mn = 0 -> stored [M][K]; a_mn = 1 -> stored [K][M]
static int make_tmap_a(CUtensorMap* m, const void* p, int a_mn, long long ld, int M, int K) {
  return a_mn ? make_tmap(m, p, M, K, ld, 64, 64) : make_tmap(m, p, K, M, ld, 64, BM);
}
// logical B[K,N]: b_mn = 0 -> stored [N][K]; b_mn = 1 -> stored [K][N]
// (cs = cluster size: a K-major B tile is loaded as cs row slices, one per CTA of the cluster)
static int make_tmap_b(CUtensorMap* m, const void* p, int b_mn, long long ld, int N, int K, int bn, int cs) {
  return b_mn ? make_tmap(m, p, N, K, ld, 64, 64) : make_tmap(m, p, K, N, ld, 64, bn / cs);
}

// 3-D bf16 tensor [steps][rows][inner] (row pitch = inner, step stride in elements); box 64 x 128 x 1
static int make_tmap_steps(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t steps,
                           uint64_t step_stride_elems) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (inner * 2) % 16 != 0 || (step_stride_elems * 2) % 16 != 0)
    return set_error(EVC_ERR_ARG, "TMA operand must be 16-byte aligned with 16-byte multiple pitches");
  cuuint64_t dims[3] = {inner, rows, steps};
  cuuint64_t strides[2] = {inner * 2, step_stride_elems * 2};
  cuuint32_t box[3] = {64, BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EVC_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed");
  return EVC_OK;
}

constexpr int kCluster = 2;   // CTAs per cluster sharing a multicast B tile
static int g_debug = 0;       // profiling experiments only (evc_debug_set)

// EVC_PAIR=0: multicast pairs (two 128-row MMAs sharing a multicast B tile) instead of cta_group::2 (A/B experiment)
static bool pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVC_PAIR");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

template <int A_MN, int B_MN, int BN, int EPI, int CS, int PAIR = 0>
static int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b, const CUtensorMap& c,
                  const GemmArgs& args_in, cudaStream_t stream, const CUtensorMap* a1lo = nullptr,
                  const CUtensorMap* a2lo = nullptr, const CUtensorMap* blo = nullptr) {
  GemmArgs args = args_in;
  args.segments = (a1lo != nullptr && blo != nullptr) ? 3 : 1;
  const CUtensorMap& xa1 = a1lo ? *a1lo : a1;
  const CUtensorMap& xa2 = a2lo ? *a2lo : (a1lo ? *a1lo : a2);
  const CUtensorMap& xb = blo ? *blo : b;
  using Cfg = GemmCfg<BN, PAIR>;
  auto kern = gemm_kernel<A_MN, B_MN, BN, EPI, CS, PAIR>;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES)) return rc;
  const int tiles_mc = (args.tiles_m + CS - 1) / CS;
  const int work = tiles_mc * args.tiles_n * args.split_k;
  if (work <= 0) return EVC_OK;
  const int max_clusters = num_sms() / CS;
  const int clusters = work < max_clusters ? work : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CS);
  cfg.blockDim = dim3(EpiCfg<EPI>::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_debug & 256) ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a1, a2, b, c, xa1, xa2, xb, args);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(gemm_kernel)");
  return check_launch("gemm_kernel");
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ generic GEMM
// force_bn: 0 = choose, else 128/256.  split_stride != 0: split-K partial slabs instead of atomics
// (*splits_out receives the number of slabs actually written).
static int gemm_store(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, int M, int N,
                      int K, void* C, int c_bf16, long long ldc, const float* bias, int split_k, int accumulate,
                      cudaStream_t stream, int force_bn = 0, long long split_stride = 0, int* splits_out = nullptr,
                      const void* A2 = nullptr, int K1 = 0, const void* A_lo = nullptr, const void* B_lo = nullptr,
                      const void* A2_lo = nullptr, bool stream_k = false, float* sumsq_out = nullptr,
                      float alpha = 1.f) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(EVC_ERR_ARG, "gemm: empty problem");
  // split-bf16 "precise" product: both residual planes or none (same shapes / pitches as the hi planes)
  const bool x2 = A_lo != nullptr || B_lo != nullptr;
  if (x2 && (A_lo == nullptr || B_lo == nullptr || (A2 != nullptr) != (A2_lo != nullptr)))
    return set_error(EVC_ERR_ARG, "gemm: split-bf16 mode needs the lo plane of every operand");
  if (split_k < 1) split_k = 1;
  if ((split_k > 1 || accumulate) && c_bf16) return set_error(EVC_ERR_ARG, "gemm: split-K/accumulate needs f32 C");
  int bn = force_bn ? force_bn : ((N > 128) ? 256 : 128);
  // weight-streaming regime (few rows, wide N: the MoE logits GEMMs, 256 x 9432 / 14148 x 4096): 128 x 256 tiles give
  // fewer work items than SMs (experts: 74 CTAs of 148) and stream the weights at a fraction of the HBM rate; 128-wide
  // tiles fill the chip -- as long as they still fit ONE round over the clusters (experts: 74 pair tiles, 33 vs 37 us;
  // gates: 111 pair tiles = two rounds, the second half empty: 52 vs 41 us with 256-wide tiles, measured on B200,
  // profiles/r02_exp_moe_gemm.txt)
  static int ws128 = -1;   // EVC_WS128=0: keep 256-wide tiles in the weight-streaming regime; 2: the round-2 rule
  if (ws128 < 0) {         // without the one-round condition (A/B experiments)
    const char* e = getenv("EVC_WS128");
    ws128 = e ? atoi(e) : 1;
  }
  if (ws128 && !force_bn && bn == 256 && split_k <= 1 && M <= 2 * BM && K <= 8192) {
    const long long tm = ceil_div(M, BM), cl = tm >= 2 ? num_sms() / kCluster : num_sms();
    const long long w256 = ceil_div(static_cast<int>(tm), tm >= 2 ? kCluster : 1) * static_cast<long long>(ceil_div(N, 256));
    const long long w128 = ceil_div(static_cast<int>(tm), tm >= 2 ? kCluster : 1) * static_cast<long long>(ceil_div(N, 128));
    if (w256 < cl && w128 > w256 && (ws128 == 2 || (w128 + cl - 1) / cl <= (w256 + cl - 1) / cl)) bn = 128;
  }
  GemmArgs g = {};
  g.M = M; g.N = N;
  g.tiles_m = ceil_div(M, BM);
  g.tiles_n = ceil_div(N, bn);
  g.kb_total = ceil_div(K, BK);
  g.kb_a1 = A2 ? K1 / BK : g.kb_total;          // A = [A1 (K1 columns) | A2] concatenated along K
  g.kb_per_split = ceil_div(g.kb_total, split_k);
  g.split_k = ceil_div(g.kb_total, g.kb_per_split);
  g.C = C; g.ldc = ldc; g.c_bf16 = c_bf16; g.bias = bias;
  g.split_stride = split_stride;
  g.debug = g_debug;
  // A too large for the L2 and several N tiles: walk N first so that each A tile is fetched from HBM once
  g.n_fastest = (static_cast<long long>(M) * K * 2 > (64LL << 20) && g.tiles_n > 1) ? 1 : 0;
  g.atomic_add = (split_stride == 0 && (g.split_k > 1 || accumulate)) ? 1 : 0;
  if (splits_out) *splits_out = g.split_k;
  CUtensorMap ta, ta2, tb;
  int rc = make_tmap_a(&ta, A, a_mn, A2 ? K1 : lda, M, A2 ? K1 : K);
  if (rc) return rc;
  ta2 = ta;
  if (A2) {
    rc = make_tmap_a(&ta2, A2, a_mn, K - K1, M, K - K1);
    if (rc) return rc;
  }
  CUtensorMap talo, ta2lo, tblo;
  if (x2) {
    rc = make_tmap_a(&talo, A_lo, a_mn, A2 ? K1 : lda, M, A2 ? K1 : K);
    if (rc) return rc;
    ta2lo = talo;
    if (A2) {
      rc = make_tmap_a(&ta2lo, A2_lo, a_mn, K - K1, M, K - K1);
      if (rc) return rc;
    }
  }
  int cs = (g.tiles_m >= 2) ? kCluster : 1;
  if (cs > 1) {   // pairing must not add a round over the SMs (odd tile counts pad to a dummy tile)
    const long long w1 = static_cast<long long>(g.tiles_m) * g.tiles_n * g.split_k;
    const long long w2 = static_cast<long long>(ceil_div(g.tiles_m, cs)) * g.tiles_n * g.split_k;
    const long long r1 = (w1 + num_sms() - 1) / num_sms();
    const long long r2 = (w2 + num_sms() / cs - 1) / (num_sms() / cs);
    if (r2 > r1) cs = 1;
  }
  rc = make_tmap_b(&tb, B, b_mn, ldb, N, K, bn, cs);
  if (rc) return rc;
  if (x2) {
    rc = make_tmap_b(&tblo, B_lo, b_mn, ldb, N, K, bn, cs);
    if (rc) return rc;
  }
  if (stream_k) {
    // stream-K over two slabs (see WorkIter): every cluster gets an equal share of the (tile, k-block) space
    const long long ctiles = static_cast<long long>(ceil_div(g.tiles_m, cs)) * g.tiles_n;
    const int clusters = num_sms() / cs;
    // (not eligible -- fewer tiles than clusters, no slab output -- : the static split-K schedule as requested)
    if (split_stride != 0 && ctiles >= clusters && g.kb_total >= 2) {
      g.stream_k = 1;
      g.split_k = 2;
      g.kb_per_split = g.kb_total;
      g.atomic_add = 0;
      if (splits_out) *splits_out = 2;
    }
  }
  const CUtensorMap* pa = x2 ? &talo : nullptr;
  const CUtensorMap* pa2 = x2 ? &ta2lo : nullptr;
  const CUtensorMap* pb = x2 ? &tblo : nullptr;
  // C through TMA bulk stores when its layout allows (16-byte aligned rows) and no atomics are needed
  CUtensorMap tc = ta;
  const int eb = c_bf16 ? 2 : 4;
  // (partial slabs are stacked along the row coordinate of one tensor map: a box must not cross a slab)
  const bool slab_rows_ok = split_stride == 0 || (split_stride % ldc == 0 && M % 32 == 0);
  if (!g.atomic_add && !(g_debug & 128) && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc * eb) % 16 == 0 &&
      slab_rows_ok) {
    const long long slab_rows = split_stride ? split_stride / ldc : 0;
    const long long total_rows = split_stride ? slab_rows * (g.split_k - 1) + M : M;
    rc = make_tmap(&tc, C, N, total_rows, ldc, c_bf16 ? 64 : 32, 32, eb);
    if (rc) return rc;
    g.tma_store = 1;
    g.split_rows = static_cast<int>(slab_rows);
  }
  if (sumsq_out != nullptr || alpha != 1.f) {
    // the scale and the sum of squares are applied where the epilogue holds the final f32 values: plain TMA stores only
    if (!g.tma_store || c_bf16 || g.split_k > 1 || split_stride != 0)
      return set_error(EVC_ERR_ARG, "gemm: alpha / sumsq_out need a plain f32 store into a 16-byte aligned C (no split-K)");
    g.sumsq_out = sumsq_out;
    g.alpha_m1 = alpha - 1.f;
  }
#define EVC_DISPATCH(AM, BMN)                                                                    \
  if (a_mn == AM && b_mn == BMN) {                                                                \
    if (cs == 1)                                                                                  \
      return bn == 256 ? launch<AM, BMN, 256, EPI_STORE, 1>(ta, ta2, tb, tc, g, stream, pa, pa2, pb)           \
                       : launch<AM, BMN, 128, EPI_STORE, 1>(ta, ta2, tb, tc, g, stream, pa, pa2, pb);          \
    if (pair_mode())                                                                              \
      return bn == 256 ? launch<AM, BMN, 256, EPI_STORE, kCluster, 1>(ta, ta2, tb, tc, g, stream, pa, pa2, pb) \
                       : launch<AM, BMN, 128, EPI_STORE, kCluster, 1>(ta, ta2, tb, tc, g, stream, pa, pa2, pb);\
    return bn == 256 ? launch<AM, BMN, 256, EPI_STORE, kCluster>(ta, ta2, tb, tc, g, stream, pa, pa2, pb)      \
                     : launch<AM, BMN, 128, EPI_STORE, kCluster>(ta, ta2, tb, tc, g, stream, pa, pa2, pb);     \
  }
  EVC_DISPATCH(0, 0)
  EVC_DISPATCH(0, 1)
  EVC_DISPATCH(1, 0)
  EVC_DISPATCH(1, 1)
#undef EVC_DISPATCH
  return set_error(EVC_ERR_ARG, "gemm: bad major flags");
}

}  // namespace evc

using namespace evc;

extern "C" int evc_debug_set(int flags) { g_debug = flags; return 0; }

extern "C" int evc_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major,
                             long long ldb, int M, int N, int K, void* C, int c_is_bf16, long long ldc,
                             const float* bias, int split_k, int accumulate, void* stream) {
  return gemm_store(A, a_mn_major, lda, B, b_mn_major, ldb, M, N, K, C, c_is_bf16, ldc, bias, split_k, accumulate,
                    static_cast<cudaStream_t>(stream));
}

extern "C" int evc_gemm_bf16x2(const void* A, const void* A_lo, int a_mn_major, long long lda, const void* B,
                               const void* B_lo, int b_mn_major, long long ldb, int M, int N, int K, void* C,
                               int c_is_bf16, long long ldc, const float* bias, int split_k, int accumulate,
                               void* stream) {
  if (A_lo == nullptr || B_lo == nullptr) return set_error(EVC_ERR_ARG, "gemm_bf16x2: lo planes required");
  return gemm_store(A, a_mn_major, lda, B, b_mn_major, ldb, M, N, K, C, c_is_bf16, ldc, bias, split_k, accumulate,
                    static_cast<cudaStream_t>(stream), 0, 0, nullptr, nullptr, 0, A_lo, B_lo);
}

extern "C" int evc_gemm_bf16_wgrad(const void* A, const void* A_lo, int a_mn_major, long long lda, const void* B,
                                   const void* B_lo, int b_mn_major, long long ldb, int M, int N, int K, float* C,
                                   long long ldc, float alpha, float* sumsq_out, void* stream) {
  return gemm_store(A, a_mn_major, lda, B, b_mn_major, ldb, M, N, K, C, 0, ldc, nullptr, 1, 0,
                    static_cast<cudaStream_t>(stream), 0, 0, nullptr, nullptr, 0, A_lo, B_lo, nullptr, false, sumsq_out,
                    alpha);
}

// ------------------------------------------------------------------ BasicLSTM layer, forward over T steps
// split-K factor for a GEMM with `tiles` output tiles and kb_total 64-deep k blocks: minimise
// (rounds over the SMs) x (k blocks per split + a per-round pipeline fill / epilogue / slab traffic
// cost worth ~16 k blocks; calibrated on the 5120x1024x4096 recurrent dgrad, where 2 splits win).
static int pick_split(int tiles, int kb_total) {
  static int forced = -1;   // experiments: EVC_FORCE_SPLIT=<n> overrides the model for multi-round problems
  if (forced < 0) {
    const char* e = getenv("EVC_FORCE_SPLIT");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0 && tiles > num_sms()) return forced;
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 16; ++s) {
    if (s > 1 && kb_total / s < 4) break;
    const long long rounds = (static_cast<long long>(tiles) * s + num_sms() - 1) / num_sms();
    const long long cost = rounds * ((kb_total + s - 1) / s + 16);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

extern "C" long long evc_lstm_workspace_bytes(int rows, int H, int Kx, int precise) {
  // forward small-row path: S x rows x 4H f32 ; backward: S x rows x H f32
  // (precise = split-bf16 mode: every forward step goes through the slab path, whatever the row count)
  int sf = (rows <= 1024 || precise) ? pick_split(ceil_div(rows, BM) * (4 * H / 256), (Kx + H) / BK) : 0;
  int sb = pick_split(ceil_div(rows, BM) * ceil_div(H, 256), 4 * H / BK);
  if (sb < 2) sb = 2;   // the stream-K schedule of the recurrent dgrad writes two slabs
  const long long f = static_cast<long long>(sf) * rows * 4 * H * 4;
  const long long b = static_cast<long long>(sb) * rows * H * 4;
  return (f > b ? f : b) + 256;
}

// ------------------------------------------------------------------ cluster split-K step (csrc/evc_cluster.cuh)
// Number of K splits (= cluster size) for a small-row step, 0 = not eligible: the tiles x KS CTAs should fit one wave.
// OFF by default (EVC_CLUSTER_STEP=1 or evc_debug_set bit 16384 enables it): measured on B200 at 256 rows it takes 19.7 /
// 15.7 us per step (Kx = 4096 / 1024) against 18.4 / 14.5 us for the slab path -- both stream the whole weight matrix
// from the L2 once per 128-row tile and step (84 MB at Kx = 4096), which is what bounds them; the DSMEM exchange
// (96 KB per CTA + two cluster barriers) costs as much as the slab round trip it removes
// (profiles/r02_resident_recurrence.md).  Only keeping the weights resident (evc_rec.cuh: 9.3 us per step) helps.
static int cluster_step_splits(int rows, int H, int Kx) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("EVC_CLUSTER_STEP");
    enabled = e ? atoi(e) : 0;
  }
  if (!(enabled || (g_debug & 16384)) || (g_debug & 8192) || H % 64 != 0 || Kx % BK != 0) return 0;
  const int tiles = ceil_div(rows, BM) * (H / 64);
  int ks = 8;
  while (ks >= 2 && tiles * ks > num_sms()) ks >>= 1;
  while (ks >= 2 && Kx / BK < 2 * ks) ks >>= 1;      // at least two k blocks per CTA at t = 0
  return ks >= 2 ? ks : 0;
}

template <int KS>
static int launch_cluster_step(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b,
                               const ClusterStepArgs& args, cudaStream_t stream) {
  auto kern = lstm_cluster_step_fwd_kernel<KS>;
  if (int rc = opt_in_smem(reinterpret_cast<const void*>(kern), CL_SMEM_BYTES)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(args.tiles_m * args.tiles_n * KS));
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = CL_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a1, a2, b, args);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(lstm_cluster_step_fwd_kernel)");
  return check_launch("lstm_cluster_step_fwd_kernel");
}

// Steps [t_begin, t_end) of the same layer: lets the host interleave the two cells of a MultiRNNCell on two
// streams (cell 1 step t only needs cell 0 step t), so that the SMs one kernel leaves idle in its last round
// over the tiles are taken by the other cell's kernel.
extern "C" int evc_lstm_seq_fwd_steps(const void* x, long long x_step_stride, int Kx, const void* W,
                                      const float* bias, int rows, int H, int T, int t_begin, int t_end,
                                      const int* seq_len, void* h_all, float* c_all, void* gates_all,
                                      void* workspace, long long workspace_bytes, const void* x_lo, const void* W_lo,
                                      void* h_lo_all, void* gates_lo_all, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || T <= 0) return set_error(EVC_ERR_ARG, "lstm_seq_fwd: empty problem");
  // split-bf16 "precise" mode: the residual planes of x, W and h (and of the saved gates when gates are saved)
  const bool x2 = x_lo != nullptr || W_lo != nullptr || h_lo_all != nullptr;
  if (x2 && (x_lo == nullptr || W_lo == nullptr || h_lo_all == nullptr || (gates_all != nullptr) != (gates_lo_all != nullptr)))
    return set_error(EVC_ERR_ARG, "lstm_seq_fwd: split-bf16 mode needs the lo planes of x, W, h (and gates)");
  if (x2 && workspace == nullptr) return set_error(EVC_ERR_ARG, "lstm_seq_fwd: split-bf16 mode needs a workspace");
  const __nv_bfloat16* xl = static_cast<const __nv_bfloat16*>(x_lo);
  __nv_bfloat16* hl = static_cast<__nv_bfloat16*>(h_lo_all);
  __nv_bfloat16* gl = static_cast<__nv_bfloat16*>(gates_lo_all);
  if (t_begin < 0 || t_end > T || t_begin >= t_end) return set_error(EVC_ERR_ARG, "lstm_seq_fwd: bad step range");
  if (H % 64 != 0 || Kx % 64 != 0) return set_error(EVC_ERR_ARG, "lstm_seq_fwd: H and Kx must be multiples of 64");
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* hb = static_cast<__nv_bfloat16*>(h_all);
  __nv_bfloat16* gb = static_cast<__nv_bfloat16*>(gates_all);
  const long long RH = static_cast<long long>(rows) * H;
  int rc;
  if (const int ksplit = (!x2 && rows <= 1024) ? cluster_step_splits(rows, H, Kx) : 0) {
    // Small-row steps as ONE kernel each: split-K over a cluster, partial sums through DSMEM, cell update fused
    CUtensorMap tbw;
    rc = make_tmap_b(&tbw, W, 1, 4LL * H, 4 * H, Kx + H, 256, 1);
    if (rc) return rc;
    for (int t = t_begin; t < t_end; ++t) {
      CUtensorMap ta1, ta2;
      rc = make_tmap_a(&ta1, xb + t * x_step_stride, 0, Kx, rows, Kx);
      if (rc) return rc;
      rc = make_tmap_a(&ta2, hb + t * RH, 0, H, rows, H);
      if (rc) return rc;
      ClusterStepArgs a = {};
      a.rows = rows; a.H = H; a.t = t;
      a.tiles_m = ceil_div(rows, BM); a.tiles_n = H / 64;
      a.kb_a1 = Kx / BK;
      a.kb_total = a.kb_a1 + (t == 0 ? 0 : H / BK);
      a.bias = bias; a.seq_len = seq_len;
      a.c_prev = (t == 0) ? nullptr : c_all + t * RH;
      a.h_prev = (t == 0) ? nullptr : hb + t * RH;
      a.c_out = c_all + (t + 1) * RH;
      a.h_out = hb + (t + 1) * RH;
      a.gates = gb ? gb + t * RH * 4 : nullptr;
      rc = ksplit == 8 ? launch_cluster_step<8>(ta1, ta2, tbw, a, stream)
         : ksplit == 4 ? launch_cluster_step<4>(ta1, ta2, tbw, a, stream)
                       : launch_cluster_step<2>(ta1, ta2, tbw, a, stream);
      if (rc) return rc;
    }
    return EVC_OK;
  }
  if ((rows <= 1024 || x2) && workspace != nullptr) {
    // Small-row steps (RNN_L2, student): one 128x256 tile per CTA would serialise the whole K on a few
    // SMs.  Split K over the SMs into f32 partial slabs, then one full-occupancy cell kernel sums
    // the slabs and applies bias / gates / state update / length mask.
    float* part = static_cast<float*>(workspace);
    const long long slab = static_cast<long long>(rows) * 4 * H;
    for (int t = t_begin; t < t_end; ++t) {
      const int K = Kx + (t == 0 ? 0 : H);          // h_{-1} = 0: skip the recurrent half at t = 0
      const int want = pick_split(ceil_div(rows, BM) * (4 * H / 256), K / BK);
      if (static_cast<long long>(want) * slab * 4 > workspace_bytes)
        return set_error(EVC_ERR_ARG, "lstm_seq_fwd: workspace too small (evc_lstm_workspace_bytes)");
      int splits = 1;
      rc = gemm_store(xb + t * x_step_stride, 0, Kx, W, 1, 4LL * H, rows, 4 * H, K, part, 0, 4LL * H, nullptr, want, 0,
                      stream, 256, slab, &splits, (t == 0) ? nullptr : hb + t * RH, Kx,
                      x2 ? xl + t * x_step_stride : nullptr, x2 ? W_lo : nullptr,
                      (x2 && t > 0) ? hl + t * RH : nullptr);
      if (rc) return rc;
      rc = launch_lstm_cell_fwd(part, splits, slab, bias, (t == 0) ? nullptr : c_all + t * RH,
                                (t == 0) ? nullptr : hb + t * RH, seq_len, t, rows, H, c_all + (t + 1) * RH,
                                hb + (t + 1) * RH, gb ? gb + t * RH * 4 : nullptr, stream,
                                (x2 && t > 0) ? hl + t * RH : nullptr, x2 ? hl + (t + 1) * RH : nullptr,
                                (x2 && gl) ? gl + t * RH * 4 : nullptr);
      if (rc) return rc;
    }
    return EVC_OK;
  }
  const int cs = (rows > BM) ? kCluster : 1;
  CUtensorMap tb;
  rc = make_tmap_b(&tb, W, 1, 4LL * H, 4 * H, Kx + H, 256, cs);
  if (rc) return rc;
  for (int t = t_begin; t < t_end; ++t) {
    CUtensorMap ta1, ta2;
    rc = make_tmap_a(&ta1, xb + t * x_step_stride, 0, Kx, rows, Kx);
    if (rc) return rc;
    rc = make_tmap_a(&ta2, hb + t * RH, 0, H, rows, H);
    if (rc) return rc;
    GemmArgs g = {};
    g.M = rows; g.N = 4 * H; g.H = H;
    g.tiles_m = ceil_div(rows, BM);
    g.tiles_n = H / 64;
    g.split_k = 1;
    g.kb_a1 = Kx / BK;
    g.kb_total = g.kb_a1 + (t == 0 ? 0 : H / BK);  // h_{-1} = 0: skip the recurrent half at t = 0
    g.kb_per_split = g.kb_total;
    g.bias = bias;
    g.debug = g_debug;
    g.t = t; g.seq_len = seq_len;
    g.c_prev = (t == 0) ? nullptr : c_all + t * RH;
    g.h_prev = (t == 0) ? nullptr : hb + t * RH;
    g.c_out = c_all + (t + 1) * RH;
    g.h_out = hb + (t + 1) * RH;
    g.gates = gb ? gb + t * RH * 4 : nullptr;
    rc = (cs == 1) ? launch<0, 1, 256, EPI_LSTM_FWD, 1>(ta1, ta2, tb, tb, g, stream)
         : pair_mode() ? launch<0, 1, 256, EPI_LSTM_FWD, kCluster, 1>(ta1, ta2, tb, tb, g, stream)
                       : launch<0, 1, 256, EPI_LSTM_FWD, kCluster>(ta1, ta2, tb, tb, g, stream);
    if (rc) return rc;
  }
  return EVC_OK;
}

// ------------------------------------------------------------------ resident-weights persistent recurrence
// (csrc/evc_rec.cuh).  Plan: G row groups x H/16 unit slices of co-resident CTAs.
struct RecPlan {
  int tiles_m, n_slices, groups, tiles_per_cta;
  long long zx_bytes, wp_bytes, flag_bytes;
  bool ok;
};
static RecPlan rec_plan(int rows, int H, int T) {
  RecPlan p = {};
  p.ok = false;
  if (rows <= 0 || T <= 0 || H % BK != 0) return p;
  p.tiles_m = ceil_div(rows, BM);
  p.n_slices = H / REC_UNITS;
  if (p.n_slices > num_sms() || rec_smem_bytes(H) > 227 * 1024) return p;
  p.groups = num_sms() / p.n_slices;
  if (p.groups > p.tiles_m) p.groups = p.tiles_m;
  p.tiles_per_cta = ceil_div(p.tiles_m, p.groups);
  if (p.tiles_per_cta > REC_MAX_TILES) return p;
  p.groups = ceil_div(p.tiles_m, p.tiles_per_cta);          // no empty groups
  p.zx_bytes = ((static_cast<long long>(T) * rows * 4 * H * 4 + 1023) / 1024) * 1024;
  p.wp_bytes = static_cast<long long>(H) * 4 * H * 2;
  p.flag_bytes = ((static_cast<long long>(T) * p.tiles_m * 4 + 255) / 256) * 256;
  p.ok = true;
  return p;
}

extern "C" long long evc_lstm_rec_workspace_bytes(int rows, int H, int T) {
  const RecPlan p = rec_plan(rows, H, T);
  return p.ok ? p.zx_bytes + p.wp_bytes + p.flag_bytes : 0;
}

extern "C" int evc_lstm_seq_fwd_resident(const void* x, long long x_step_stride, int Kx, const void* W,
                                         const float* bias, int rows, int H, int T, const int* seq_len, void* h_all,
                                         float* c_all, void* gates_all, void* workspace, long long workspace_bytes,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const RecPlan p = rec_plan(rows, H, T);
  if (!p.ok) return set_error(EVC_ERR_UNSUPPORTED, "lstm_seq_fwd_resident: shape not eligible (evc_lstm_rec_workspace_bytes == 0)");
  if (Kx % BK != 0) return set_error(EVC_ERR_ARG, "lstm_seq_fwd_resident: Kx must be a multiple of 64");
  if (x_step_stride != static_cast<long long>(rows) * Kx)
    return set_error(EVC_ERR_ARG, "lstm_seq_fwd_resident: the input steps must be contiguous ([T*rows, Kx])");
  if (workspace == nullptr || workspace_bytes < p.zx_bytes + p.wp_bytes + p.flag_bytes)
    return set_error(EVC_ERR_ARG, "lstm_seq_fwd_resident: workspace too small (evc_lstm_rec_workspace_bytes)");
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0)
    return set_error(EVC_ERR_ARG, "lstm_seq_fwd_resident: workspace must be 1024-byte aligned");
  char* ws = static_cast<char*>(workspace);
  float* zx = reinterpret_cast<float*>(ws);
  __nv_bfloat16* wp = reinterpret_cast<__nv_bfloat16*>(ws + p.zx_bytes);
  unsigned int* ready = reinterpret_cast<unsigned int*>(ws + p.zx_bytes + p.wp_bytes);
  const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(W);
  __nv_bfloat16* hb = static_cast<__nv_bfloat16*>(h_all);
  // 1. the slice layout of the recurrent rows of the kernel
  {
    const long long n = static_cast<long long>(p.n_slices) * H * 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>((n + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_pack_wh_kernel, wb + static_cast<long long>(Kx) * 4 * H, H, wp);
    count_launch();
    if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(lstm_pack_wh_kernel)");
  }
  // 2. the input projection of all steps: Zx = X Wx + bias  ([T*rows, Kx] x [Kx, 4H], f32)
  int rc = gemm_store(x, 0, Kx, W, 1, 4LL * H, T * rows, 4 * H, Kx, zx, 0, 4LL * H, bias, 1, 0, stream);
  if (rc) return rc;
  // 3. the recurrence
  cudaError_t e = cudaMemsetAsync(ready, 0, static_cast<size_t>(p.flag_bytes), stream);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(ready flags)");
  CUtensorMap th, tw;
  rc = make_tmap_steps(&th, hb, H, rows, T + 1, static_cast<long long>(rows) * H);
  if (rc) return rc;
  rc = make_tmap(&tw, wp, REC_BN, static_cast<uint64_t>(p.n_slices) * H, REC_BN, REC_BN, BK);
  if (rc) return rc;
  RecArgs a = {};
  a.rows = rows; a.H = H; a.T = T;
  a.tiles_m = p.tiles_m; a.groups = p.groups; a.tiles_per_cta = p.tiles_per_cta;
  a.zx = zx; a.seq_len = seq_len; a.c_all = c_all; a.h_all = hb;
  a.gates_all = static_cast<__nv_bfloat16*>(gates_all);
  a.ready = ready;
  const size_t smem = rec_smem_bytes(H);
  // (the opt-in is cached per kernel and device: ask for the maximum once, launches use what their H needs)
  rc = opt_in_smem(reinterpret_cast<const void*>(lstm_rec_resident_fwd_kernel), 227 * 1024);
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.groups * p.n_slices));
  cfg.blockDim = dim3(REC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, lstm_rec_resident_fwd_kernel, th, tw, a);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(lstm_rec_resident_fwd_kernel)");
  return check_launch("lstm_rec_resident_fwd_kernel");
}

extern "C" int evc_lstm_seq_fwd(const void* x, long long x_step_stride, int Kx, const void* W, const float* bias,
                                int rows, int H, int T, const int* seq_len, void* h_all, float* c_all,
                                void* gates_all, void* workspace, long long workspace_bytes, void* stream_) {
  return evc_lstm_seq_fwd_steps(x, x_step_stride, Kx, W, bias, rows, H, T, 0, T, seq_len, h_all, c_all, gates_all,
                                workspace, workspace_bytes, nullptr, nullptr, nullptr, nullptr, stream_);
}

// ------------------------------------------------------------------ BasicLSTM layer, backward over T steps
extern "C" int evc_lstm_seq_bwd(const void* W, int Kx, int rows, int H, int T, const int* seq_len,
                                const void* gates_all, const float* c_all, const float* dh_ext_all,
                                const float* dh_final, long long ld_dh_final, const float* dc_final,
                                long long ld_dc_final, float* dh_pass, float* dc, void* dz_all, float* dbias,
                                void* workspace, long long workspace_bytes, const void* W_lo, const void* gates_lo_all,
                                void* dz_lo_all, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows <= 0 || T <= 0) return set_error(EVC_ERR_ARG, "lstm_seq_bwd: empty problem");
  const bool x2 = W_lo != nullptr || gates_lo_all != nullptr || dz_lo_all != nullptr;
  if (x2 && (W_lo == nullptr || gates_lo_all == nullptr || dz_lo_all == nullptr || workspace == nullptr))
    return set_error(EVC_ERR_ARG, "lstm_seq_bwd: split-bf16 mode needs the lo planes of W, gates, dz and a workspace");
  const __nv_bfloat16* gl = static_cast<const __nv_bfloat16*>(gates_lo_all);
  __nv_bfloat16* zl = static_cast<__nv_bfloat16*>(dz_lo_all);
  const __nv_bfloat16* whl = x2 ? static_cast<const __nv_bfloat16*>(W_lo) + static_cast<long long>(Kx) * 4 * H : nullptr;
  // Which form of the backward step: (a) split-K / stream-K GEMM into f32 slabs + cell kernel (default with a
  // workspace), (b) ONE kernel per step whose epilogue is the cell backward (EPI_LSTM_BWD: no slab round trip, no
  // second launch; the only form without a workspace).  Measured on B200 with cta_group::2 pairs, 8 epilogue warps
  // and the bias sums fused (round 2, same-call A/B, profiles/r02_pair_ab.txt): (b) makes the joint step 5 % SLOWER
  // (12.93 vs 12.31 ms) -- 128-wide tiles pull 1.33x the operand bytes per FLOP and the 32 B/element epilogue does
  // not hide behind a 64-k-block main loop -- so (a) stays the default; EVC_FUSED_BWD=1 selects (b) above 1024 rows.
  static int fused_default = -1;
  if (fused_default < 0) {
    const char* e = getenv("EVC_FUSED_BWD");
    fused_default = e ? atoi(e) : 0;
  }
  // (evc_debug_set bit 2048 forces the slab path, bit 4096 the fused path: tests exercise both in one process)
  const bool want_fused = workspace == nullptr || (g_debug & 4096) ||
                          (!(g_debug & 2048) && fused_default != 0 && rows > 1024);
  if (H % 128 != 0) return set_error(EVC_ERR_ARG, "lstm_seq_bwd: H must be a multiple of 128");
  const __nv_bfloat16* gb = static_cast<const __nv_bfloat16*>(gates_all);
  __nv_bfloat16* zb = static_cast<__nv_bfloat16*>(dz_all);
  const __nv_bfloat16* wh = static_cast<const __nv_bfloat16*>(W) + static_cast<long long>(Kx) * 4 * H;
  const long long RH = static_cast<long long>(rows) * H;
  if (workspace != nullptr && (x2 || !want_fused)) {
    // Recurrent dgrad dh = dz_{t+1} Wh^T as a (split-K) GEMM into f32 slabs + a full-occupancy cell
    // kernel.  Measured faster than the fused epilogue at every row count: the cell backward moves
    // 36 B per element, which 4 epilogue warps per SM cannot keep in flight behind a 4096-deep GEMM.
    float* part = static_cast<float*>(workspace);
    // (128x128 tiles without split-K -- twice the tiles, half the slab traffic -- measured 2 % slower per step)
    const int want = pick_split(ceil_div(rows, BM) * ceil_div(H, 256), 4 * H / BK);
    if (static_cast<long long>(want < 2 ? 2 : want) * RH * 4 > workspace_bytes)
      return set_error(EVC_ERR_ARG, "lstm_seq_bwd: workspace too small (evc_lstm_workspace_bytes)");
    // stream-K (two slabs, every cluster the same number of k blocks) when there are more tiles than clusters:
    // the 160 tiles of the 5120-row dgrad otherwise quantise into 3 rounds for 2.16 rounds of work
    static int use_sk = -1;
    if (use_sk < 0) {
      const char* e = getenv("EVC_STREAMK");
      use_sk = e ? atoi(e) : 1;
    }
    if (dbias != nullptr) {
      cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * 4 * H, stream);
      if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(dbias)");
    }
    for (int t = T - 1; t >= 0; --t) {
      const bool last = (t == T - 1);
      int splits = 0;
      if (!last) {
        int rc = gemm_store(zb + (t + 1) * RH * 4, 0, 4LL * H, wh, 0, 4LL * H, rows, H, 4 * H, part, 0, H, nullptr,
                            want, 0, stream, 256, RH, &splits, nullptr, 0, x2 ? zl + (t + 1) * RH * 4 : nullptr, whl,
                            nullptr, use_sk != 0);
        if (rc) return rc;
      }
      int rc = launch_lstm_cell_bwd(part, splits, RH, gb + t * RH * 4, (t == 0) ? nullptr : c_all + t * RH,
                                    dh_ext_all ? dh_ext_all + t * RH : nullptr, H, last ? dh_final : dh_pass,
                                    last ? ld_dh_final : H, last ? dc_final : dc, last ? ld_dc_final : H, seq_len, t,
                                    rows, H, zb + t * RH * 4, dc, dh_pass, dbias, stream,
                                    x2 ? gl + t * RH * 4 : nullptr, x2 ? zl + t * RH * 4 : nullptr);
      if (rc) return rc;
    }
    return EVC_OK;
  }
  const int cs = (rows > BM) ? kCluster : 1;
  CUtensorMap tb;  // B[k = gate column, n = unit] = Wh[unit][gate column] : stored [N][K] = K-major
  int rc = make_tmap_b(&tb, wh, 0, 4LL * H, H, 4 * H, 128, cs);
  if (rc) return rc;
  if (dbias != nullptr) {
    cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * 4 * H, stream);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(dbias)");
  }
  for (int t = T - 1; t >= 0; --t) {
    const bool last = (t == T - 1);
    if (last && workspace != nullptr) {
      // no recurrent term at the last step: the stand-alone cell kernel (HBM-bound, no tensor work to fuse with)
      rc = launch_lstm_cell_bwd(static_cast<const float*>(workspace), 0, RH, gb + t * RH * 4,
                                (t == 0) ? nullptr : c_all + t * RH, dh_ext_all ? dh_ext_all + t * RH : nullptr, H,
                                dh_final, ld_dh_final, dc_final, ld_dc_final, seq_len, t, rows, H, zb + t * RH * 4, dc,
                                dh_pass, dbias, stream);
      if (rc) return rc;
      continue;
    }
    CUtensorMap ta;
    // A = dz_{t+1} [rows, 4H]; at the last step there is no recurrent term (the map is unused)
    rc = make_tmap_a(&ta, zb + (last ? t : t + 1) * RH * 4, 0, 4LL * H, rows, 4 * H);
    if (rc) return rc;
    GemmArgs g = {};
    g.M = rows; g.N = H; g.H = H;
    g.tiles_m = ceil_div(rows, BM);
    g.tiles_n = H / 128;
    g.split_k = 1;
    g.kb_total = last ? 0 : (4 * H) / BK;
    g.kb_a1 = g.kb_total;
    g.kb_per_split = g.kb_total > 0 ? g.kb_total : 1;
    g.debug = g_debug;
    g.t = t; g.seq_len = seq_len;
    g.gates = const_cast<__nv_bfloat16*>(gb + t * RH * 4);
    g.c_prev = (t == 0) ? nullptr : c_all + t * RH;
    g.dh_ext = dh_ext_all ? dh_ext_all + t * RH : nullptr;
    g.ld_dh_ext = H;
    g.dh_pass_in = last ? dh_final : dh_pass;
    g.ld_dh_pass_in = last ? ld_dh_final : H;
    g.dc_in = last ? dc_final : dc;
    g.ld_dc_in = last ? ld_dc_final : H;
    g.dh_pass_out = dh_pass;
    g.dc_out = dc;
    g.dz_out = zb + t * RH * 4;
    g.dbias = dbias;
    rc = (cs == 1) ? launch<0, 0, 128, EPI_LSTM_BWD, 1>(ta, ta, tb, tb, g, stream)
         : pair_mode() ? launch<0, 0, 128, EPI_LSTM_BWD, kCluster, 1>(ta, ta, tb, tb, g, stream)
                       : launch<0, 0, 128, EPI_LSTM_BWD, kCluster>(ta, ta, tb, tb, g, stream);
    if (rc) return rc;
  }
  return EVC_OK;
}
