/*
 * libevc -- C ABI of the B200-native (sm_100a) Hierarchical-LSTM teacher-student hot path
 * of shwetabhardwaj44/EfficientVideoClassification_Youtube8M.
 *
 * The reference is a pure-Python TensorFlow-1.x graph: it has no FFI of its own (SURVEY.md
 * 2.1), so each entry point below cites the reference op chain (file:line, relative to
 * /root/reference/code_student_uniform) whose arithmetic it replaces.  The Python plugin
 * classes in efficientvideoclassification_youtube8m_b200/ bind these with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, performs
 *     no allocation and no host synchronisation (so sequences of calls are CUDA-graph capturable)
 *   - return value: 0 = ok, <0 = EVC_ERR_*; text via evc_last_error() (thread local)
 *   - "bf16" buffers are __nv_bfloat16 (passed as void*); TMA operands must be 16-byte aligned
 *     with a row pitch that is a multiple of 8 elements
 *   - lstm weight `W` is the BasicLSTMCell kernel [in+H, 4H] (rows = [x ; h], gate column order
 *     i, j, f, o) exactly as TF lays it out, converted to bf16
 *   - SPLIT-BF16 ("precise") MODE.  Arguments named `*_lo` are nullable residual planes: a value v is held as
 *     hi = bf16(v) in the main buffer and lo = bf16(v - hi) in the `_lo` buffer of the same shape and pitch, and
 *     every contraction forms A_hi*B_hi + A_hi*B_lo + A_lo*B_hi in the f32 accumulator (~16 mantissa bits per
 *     operand; 3x the tensor work).  With all `_lo` arguments NULL the library computes with plain bf16 operands.
 *     The north_star's "loss within 1 % over the first 200 steps" at the reference's learning rate 1e-3 needs
 *     it: tests/noise_floor.py shows that rounding ANY product's operands to 8 mantissa bits moves single
 *     steps of the curve by > 10 %, while f32 and split-bf16 stay below 0.03 %.
 */
#ifndef EVC_H_
#define EVC_H_

#ifdef __cplusplus
extern "C" {
#endif

#define EVC_OK 0
#define EVC_ERR_ARG (-1)
#define EVC_ERR_CUDA (-2)
#define EVC_ERR_UNSUPPORTED (-3)

int evc_version(void);
const char* evc_last_error(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
long long evc_launch_count(void);

/* profiling experiments and tests only (bit flags): 128 = plain GEMMs store through the LSU path instead of TMA bulk
 * stores, 256 = no programmatic dependent launch for the GEMM kernels, 1024 = release the dependent grid at
 * kernel start instead of at the last tile, 2048 / 4096 = evc_lstm_seq_bwd always takes the slab path / the fused
 * dgrad + cell-backward kernel, 16384 = small-row forward steps run as the cluster split-K kernel of csrc/evc_cluster.cuh
 * (measured slower than the slab path, off by default), 8192 = never.  Also settable with the EVC_DEBUG environment variable. */
int evc_debug_set(int flags);

/* ---- input: tf.nn.l2_normalize (train.py:256) + uniform gather (train.py:265-272) or
 * gather_nd of sampled frames (model_utils.py:34-36,55-58), fused with the split into
 * num_chunks sub-sequences (frame_level_models.py:237,307).
 * src f32 [B,T,D]; frame_idx int32 [K] (idx_per_batch=0), [B,K] (=1) or NULL (identity, K<=T).
 * out_bf16 (nullable): [(tt*num_chunks + chunk)*B + b][D], frame k = chunk*(K/num_chunks)+tt.
 * out_f32 (nullable): [B,K,D] (the tensor the reference feeds to create_model*). */
int evc_frames_pack(const float* src, int B, int T, int D, const int* frame_idx, int idx_per_batch, int K,
                    int num_chunks, int normalize, void* out_bf16, float* out_f32, void* out_lo, void* stream);

/* Same from the uint8 features of the tfrecords: readers.py:160-172 (decode_raw -> float32) +
 * utils.Dequantize (utils.py:9-25: q*(4/255) + (4/512 - 2), float32) + zero padding of frames
 * >= num_frames (readers.py:173 resize_axis), then as evc_frames_pack. */
int evc_frames_pack_u8(const unsigned char* src, const int* num_frames, int B, int T, int D, const int* frame_idx,
                       int idx_per_batch, int K, int num_chunks, int normalize, void* out_bf16, float* out_f32,
                       void* out_lo, void* stream);

/* train.py:263-264: int64((num_frames / 300) * int(300/every_n)), evaluated in float64. */
int evc_num_frames_student(const int* num_frames, int B, int max_frames, int every_n, long long* out,
                           void* stream);

/* frame_level_models.py:240,256 / :309,327: len_l1[c*B+b] = min(l, max(0, n - l*c)),
 * len_l2[b] = int32(ceil(float32(n)/l)).  num_frames is int32 (is_int64=0) or int64 (=1). */
int evc_lstm_lengths(const void* num_frames, int is_int64, int B, int num_chunks, int chunk_len, int* len_l1,
                     int* len_l2, void* stream);

/* model_utils.py:49-53 SampleRandomFrames index rule: idx[b,k] = int32(u[b,k]*float32(n_b)). */
int evc_random_frame_index(const float* u, const int* num_frames, int B, int K, int* idx, void* stream);
/* model_utils.py:23-33 SampleRandomSequence: start=int32(u[b]*float32(max(n-K,0)+1)); min(start+k,n-1). */
int evc_random_sequence_index(const float* u, const int* num_frames, int B, int K, int* idx, void* stream);

/* Sequence lengths of a randomly sampled student input (BASELINE config #5): both samplers index inside
 * [0, num_frames), so all K sampled frames are real frames: out[b] = num_frames[b] > 0 ? K : 0 (int64, the
 * dtype create_model_inference takes, frame_level_models.py:308-309). */
int evc_sampled_lengths(const int* num_frames, int B, int K, long long* out, void* stream);

/* tf.random_uniform([n], float32) of the two samplers (model_utils.py:25-27,50-51): Philox4x32-10, element i =
 * lane i%4 of counter offset + i/4 under key `seed`, 23 random mantissa bits -> [0,1) as TF's Uint32ToFloat. */
int evc_random_uniform(unsigned long long seed, unsigned long long offset, float* out, long long n, void* stream);

/* ---- dense contraction on tcgen05 tensor cores: C[M,N] (=|+=) A[M,K] * B[K,N] (+ bias[N]).
 * a_mn_major=0: A stored [M][lda] (K contiguous); 1: stored [K][lda] (M contiguous).
 * b_mn_major=0: B stored [N][ldb] (K contiguous); 1: stored [K][ldb] (N contiguous).
 * C f32 or bf16 with row pitch ldc.  split_k>1 or accumulate!=0 adds into C with f32 atomics
 * (C must then be f32 and pre-initialised).  Replaces TF MatMul inside BasicLSTMCell._linear
 * and slim.fully_connected (video_level_models.py:423-435) and their gradients. */
int evc_gemm_bf16(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb,
                  int M, int N, int K, void* C, int c_is_bf16, long long ldc, const float* bias, int split_k,
                  int accumulate, void* stream);
/* The same contraction in split-bf16 mode (A_lo / B_lo: residual planes, same layout as A / B). */
int evc_gemm_bf16x2(const void* A, const void* A_lo, int a_mn_major, long long lda, const void* B, const void* B_lo,
                    int b_mn_major, long long ldb, int M, int N, int K, void* C, int c_is_bf16, long long ldc,
                    const float* bias, int split_k, int accumulate, void* stream);
/* Weight-gradient form: C f32 = alpha * (A * B), stored plainly (no bias, no split-K; A_lo / B_lo nullable residual
 * planes), and sumsq_out[0] += sum of C^2 (nullable), taken in the epilogue from the accumulators -- the per-variable
 * gradient norm of slim's clip_gradient_norms (train.py:329-334) without a second pass over the gradient.  alpha: 1, or
 * 1/world when the contraction runs over the gathered batch of all data-parallel ranks (the averaged gradient).
 * C must be 16-byte aligned. */
int evc_gemm_bf16_wgrad(const void* A, const void* A_lo, int a_mn_major, long long lda, const void* B, const void* B_lo,
                        int b_mn_major, long long ldb, int M, int N, int K, float* C, long long ldc, float alpha,
                        float* sumsq_out, void* stream);

/* ---- tf.nn.dynamic_rnn(BasicLSTMCell(H, forget_bias=1.0), x, sequence_length) for ONE cell of
 * the MultiRNNCell stack (frame_level_models.py:221-257, 291-328), all `rows` sequences at once.
 * Per step one fused kernel: [x_t | h_{t-1}] * W on tcgen05, then in the epilogue bias, gate
 * non-linearities, c/h update and the `t >= sequence_length` state copy-through.
 * x bf16: step t at x + t*x_step_stride, [rows,Kx].  h_all bf16 [(T+1),rows,H] and c_all f32
 * [(T+1),rows,H] receive the state after every step (slot 0 = initial zero state, not read).
 * gates_all bf16 [T,rows,4H] (nullable) keeps the post-activation gates for the backward pass. */
int evc_lstm_seq_fwd(const void* x, long long x_step_stride, int Kx, const void* W, const float* bias, int rows,
                     int H, int T, const int* seq_len, void* h_all, float* c_all, void* gates_all, void* workspace,
                     long long workspace_bytes, void* stream);
/* Steps [t_begin, t_end) of evc_lstm_seq_fwd (same buffers, same arithmetic).  cell 1 of a MultiRNNCell
 * only needs cell 0's output of the same step (tf.nn.dynamic_rnn evaluates the stack step by step,
 * frame_level_models.py:221-249), so the host interleaves the two cells' steps on two streams. */
int evc_lstm_seq_fwd_steps(const void* x, long long x_step_stride, int Kx, const void* W, const float* bias,
                           int rows, int H, int T, int t_begin, int t_end, const int* seq_len, void* h_all,
                           float* c_all, void* gates_all, void* workspace, long long workspace_bytes,
                           const void* x_lo, const void* W_lo, void* h_lo_all, void* gates_lo_all, void* stream);
/* Scratch needed by evc_lstm_seq_fwd / evc_lstm_seq_bwd for one cell (split-K partial slabs).
 * Forward steps with <= 1024 rows (RNN_L2, the student): with a workspace a split-K GEMM into f32 slabs + a
 * full-occupancy cell kernel; without one the fused-epilogue kernel.  (An alternative -- ONE kernel per step that splits
 * K over a thread-block cluster and exchanges the partial sums through distributed shared memory, csrc/evc_cluster.cuh --
 * is kept behind EVC_CLUSTER_STEP=1: parity-tested, measured 7 % slower.) */
long long evc_lstm_workspace_bytes(int rows, int H, int Kx, int precise);

/* The same layer with the recurrence as ONE persistent launch whose CTAs keep their slice of the recurrent
 * weights (4 gate columns x 16 units x H rows of `W`, 128 KB at H = 1024) resident in shared memory for all T
 * steps (csrc/evc_rec.cuh; small-row regime: RNN_L2, the student's RNN_L1).  The input half of the matmul is
 * hoisted into one GEMM over all steps (x must be contiguous over the steps: x_step_stride == rows*Kx).
 * evc_lstm_rec_workspace_bytes returns 0 when the shape is not eligible (too many rows for one wave of
 * co-resident CTAs, H too large for the resident slice); workspace must be 1024-byte aligned.  The caller must
 * not run two of these launches concurrently on different streams (all CTAs of a launch must be co-resident). */
long long evc_lstm_rec_workspace_bytes(int rows, int H, int T);
int evc_lstm_seq_fwd_resident(const void* x, long long x_step_stride, int Kx, const void* W, const float* bias,
                              int rows, int H, int T, const int* seq_len, void* h_all, float* c_all,
                              void* gates_all, void* workspace, long long workspace_bytes, void* stream);

/* Backward twin: for t = T-1..0 one fused kernel computing dz_{t+1} * Wh^T on tcgen05 and, in the
 * epilogue, the gate gradients dz_t (bf16 [T,rows,4H]) with the sequence_length mask.
 * dh_ext_all f32 [T,rows,H] (nullable): gradient w.r.t. the cell output at each step (from the cell
 * above).  dh_final/dc_final (nullable, row pitches ld_*): gradient w.r.t. the final state.
 * dh_pass, dc: f32 [rows,H] scratch.  Two forms of a step: ONE kernel whose epilogue is the cell backward (no
 * workspace needed; the default above 1024 rows), or a split-K / stream-K GEMM into f32 slabs in `workspace` + a
 * full-occupancy cell kernel (small row counts, and always in split-bf16 mode).
 * dbias f32 [4H] (nullable): the bias gradient = column sums of dz over all steps and rows, accumulated by the
 * kernel that writes dz (zeroed by this call). */
int evc_lstm_seq_bwd(const void* W, int Kx, int rows, int H, int T, const int* seq_len, const void* gates_all,
                     const float* c_all, const float* dh_ext_all, const float* dh_final, long long ld_dh_final,
                     const float* dc_final, long long ld_dc_final, float* dh_pass, float* dc, void* dz_all,
                     float* dbias, void* workspace, long long workspace_bytes, const void* W_lo,
                     const void* gates_lo_all, void* dz_lo_all, void* stream);

/* final MultiRNNCell state [c0|h0|c1|h1] (state_is_tuple=False; frame_level_models.py:252,257). */
int evc_state_pack(const float* c0, const void* h0, const float* c1, const void* h1, int rows, int H,
                   void* out_bf16, float* out_f32, const void* h0_lo, const void* h1_lo, void* out_lo, void* stream);

/* f32 [rows,cols] -> bf16 [rows,ld] operand copy (pad columns zeroed). */
int evc_cast_bf16(const float* src, long long rows, int cols, int ld, void* dst, void* dst_lo, void* stream);
int evc_fill_f32(float* p, long long n, float value, void* stream);

/* ---- MoeModel mixture (video_level_models.py:437-447).
 * G f32 [B,ldg] gate logits (column c*(M+1)+m), E f32 [B,lde] expert logits (c*M+m, bias added).
 * p_out f32 [B,V] = sum_{m<M} softmax(G[b,c,:])[m] * sigmoid(E[b,c,m]). */
int evc_moe_mix_fwd(const float* G, long long ldg, const float* E, long long lde, int B, int V, int M,
                    float* p_out, void* stream);
/* its backward: dP f32 [B,V] -> dG, dE (bf16 GEMM operands with row pitches lddg / ldde). */
int evc_moe_mix_bwd(const float* G, long long ldg, const float* E, long long lde, const float* dP, int B, int V,
                    int M, void* dG, long long lddg, void* dE, long long ldde, void* dG_lo, void* dE_lo,
                    void* stream);

/* CrossEntropyLoss rows (losses.py:90-97, eps=1e-5; labels u8 [B,V], nullable) and L_PRED rows
 * KL(Categorical(probs=PT) || Categorical(probs=P)) (train.py:398-402; PT nullable), and
 * dP (nullable) = d(ce_scale*CE_row + kl_scale*KL_row)/dP. */
int evc_ce_kl_loss(const float* P, const float* PT, const unsigned char* labels, int B, int V, float ce_scale,
                   float kl_scale, float* ce_rows, float* kl_rows, float* dP, void* stream);

/* The three calls above in one launch for the training step: mixture forward (P), CE rows, KL rows
 * (PT nullable), and dG/dE = d(ce_scale*CE_row + kl_scale*KL_row)/d logits. */
int evc_moe_mix_loss(const float* G, long long ldg, const float* E, long long lde, const float* PT,
                     const unsigned char* labels, int B, int V, int M, float ce_scale, float kl_scale, float* P,
                     float* ce_rows, float* kl_rows, void* dG, long long lddg, void* dE, long long ldde,
                     void* dG_lo, void* dE_lo, void* stream);

/* out[0] = scale * sum rows[0..n)  (tf.reduce_mean / reduce_sum over the batch). */
int evc_reduce_rows(const float* rows, int n, float scale, float* out, void* stream);

/* L_REP (train.py:359-362): rows[b] = sum_j (t-s)^2; d_student (nullable) = grad_scale*(s-t). */
int evc_rep_loss(const float* teacher_state, const float* student_state, int B, int S, float grad_scale,
                 float* rows, float* d_student, void* stream);

/* out[n] += sum_r X[r,n] (bias gradients; out pre-zeroed). */
int evc_colsum_bf16(const void* X, long long rows, int N, long long ld, float* out, void* stream);

/* ---- slim.learning.create_train_op (train.py:329-334,413-418): per-variable clip_by_norm + Adam.
 * evc_sumsq: out[0] += sum (g + weight_decay*w)^2 and out_wsq[0] += sum w^2 (w, out_wsq nullable).
 * evc_clip_adam: g' = (g + wd*w) * c*min(rsqrt(normsq),1/c) (clip_norm<=0: no clip); TF ApplyAdam
 * with lr_t read from device memory; refreshes the bf16 operand copy (nullable). */
int evc_sumsq(const float* g, const float* w, float weight_decay, long long n, float* out, float* out_wsq,
              void* stream);
/* [TF adam.py] step[0] += 1; lr_t[0] = lr*sqrt(1-beta2^t)/(1-beta1^t) (device-resident step counter). */
int evc_adam_lr(long long* step, float lr, float beta1, float beta2, float* lr_t, void* stream);
int evc_clip_adam(float* w, const float* g, float* m, float* v, long long n, const float* normsq,
                  float clip_norm, float weight_decay, const float* lr_t, float beta1, float beta2, float eps,
                  void* shadow_bf16, int cols, long long ld_shadow, void* shadow_lo, void* stream);
/* The same update with the squared norm of the regularised gradient assembled from parts taken where they are cheap
 * instead of by an evc_sumsq pass over g and w (slim clip_gradient_norms on g + wd*w, train.py:329-334):
 *   |g + wd*w|^2 = *normsq + *normsq_fused + 2*wd * *reg_cross + wd^2 * *reg_wsq      (the last three nullable)
 * normsq_fused = sum g^2 from the weight-gradient GEMM's epilogue (evc_gemm_bf16_wgrad), reg_cross = <g, w>
 * (evc_reg_cross), reg_wsq = sum w^2.  wsq_out (nullable): += sum of the squares of the UPDATED weights, i.e. the
 * next step's reg_wsq and regulariser value (video_level_models.py:428,434). */
int evc_clip_adam_fused(float* w, const float* g, float* m, float* v, long long n, const float* normsq,
                        float clip_norm, float weight_decay, const float* lr_t, float beta1, float beta2, float eps,
                        void* shadow_bf16, int cols, long long ld_shadow, void* shadow_lo, const float* normsq_fused,
                        const float* reg_cross, const float* reg_wsq, float* wsq_out, void* stream);
/* <g, w> of a fully connected layer (slim.fully_connected, video_level_models.py:423-435) from its logits and the
 * gradient w.r.t. them, without reading g or w:  g = X^T dL, logits = X w + bias  =>  <g, w> = sum dL * (logits - bias).
 * logits f32 [B, ld_logits], dlogits bf16 [B, ld_dlogits] (+ residual plane, nullable), bias [N] nullable; out[0] += sum. */
int evc_reg_cross(const float* logits, long long ld_logits, const void* dlogits, const void* dlogits_lo,
                  long long ld_dlogits, const float* bias, int B, int N, float* out, void* stream);

/* ---- eval_util.py:118-124 top_k_triplets: per video the k largest predictions (value desc,
 * lower class index first among equals), their values and (nullable) labels. */
int evc_topk(const float* P, int B, int V, int k, const unsigned char* labels, int* idx_out, float* val_out,
             unsigned char* lab_out, void* stream);

/* ---- the per-batch training / evaluation metrics of eval_util.py on the device, given evc_topk's output for the
 * same batch (idx/val/lab [B,k]): out[0] = hit@1 (eval_util.py:17-31), out[1] = PERR (:34-59), out[2] = GAP over
 * the pooled top-k triplets (:61-79, ties ordered by (class, video)), out[3] = mean of loss_rows (nullable).
 * Scratch: perr_rows f32 [B], npos_rows int32 [B], acc double[1] (zero before the first call; reset by the call).
 * Epoch accumulators (nullable): class_pos int32 [V] += positives per class (:114), sums double[4] += (videos,
 * hit sum, perr sum, loss sum) -- what EvaluationMetrics.accumulate keeps (:139-167). */
int evc_batch_metrics(const float* P, const unsigned char* labels, int B, int V, int k, const int* idx,
                      const float* val, const unsigned char* lab, const float* loss_rows, float* perr_rows,
                      int* npos_rows, int* class_pos, double* acc, float* out, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVC_H_ */
