"""Thin torch-tensor wrappers over the C ABI (include/evc.h).  Every function launches the
hand-written sm_100a kernels on the current CUDA stream; nothing here computes on the host
or through torch operators."""
from __future__ import annotations

from typing import Optional  # noqa: F401

import torch

from ._lib import check, lib, ptr, stream

BF16 = torch.bfloat16


def _cuda(*ts):
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("libevc operands must be contiguous CUDA tensors")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise ValueError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: the "
                             "library launches on the current device (call torch.cuda.set_device first)")


def pad8(n: int, mult: int = 64) -> int:
    return (n + mult - 1) // mult * mult


def gemm(A, B, M, N, K, out, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, bias=None,
         split_k=1, accumulate=False, A_lo=None, B_lo=None):
    """out[M,N] (=|+=) A[M,K] @ B[K,N]; see evc_gemm_bf16 for the storage conventions.  A_lo / B_lo: residual
    planes of the split-bf16 mode (same layout as A / B; both or none)."""
    lda = lda if lda is not None else A.stride(0)
    ldb = ldb if ldb is not None else B.stride(0)
    ldc = ldc if ldc is not None else out.stride(0)
    if A_lo is not None or B_lo is not None:
        check(lib.evc_gemm_bf16x2(ptr(A), ptr(A_lo), int(a_mn), lda, ptr(B), ptr(B_lo), int(b_mn), ldb, M, N, K,
                                  ptr(out), int(out.dtype == BF16), ldc, ptr(bias), split_k, int(accumulate), stream()),
              "evc_gemm_bf16x2")
        return out
    check(lib.evc_gemm_bf16(ptr(A), int(a_mn), lda, ptr(B), int(b_mn), ldb, M, N, K, ptr(out),
                            int(out.dtype == BF16), ldc, ptr(bias), split_k, int(accumulate), stream()),
          "evc_gemm_bf16")
    return out


def gemm_wgrad(A, B, M, N, K, out, sumsq_out=None, a_mn=False, b_mn=False, lda=None, ldb=None, ldc=None, A_lo=None,
               B_lo=None, alpha=1.0):
    """Weight-gradient GEMM: out[M,N] (f32) = alpha * A @ B and, if given, sumsq_out[0] += sum(out^2) from the epilogue."""
    lda = lda if lda is not None else A.stride(0)
    ldb = ldb if ldb is not None else B.stride(0)
    ldc = ldc if ldc is not None else out.stride(0)
    assert out.dtype == torch.float32 and (sumsq_out is None or sumsq_out.dtype == torch.float32)
    check(lib.evc_gemm_bf16_wgrad(ptr(A), ptr(A_lo), int(a_mn), lda, ptr(B), ptr(B_lo), int(b_mn), ldb, M, N, K,
                                  ptr(out), ldc, float(alpha), ptr(sumsq_out), stream()), "evc_gemm_bf16_wgrad")
    return out


def reg_cross(logits, ld_logits, dlogits, ld_dlogits, bias, B, N, out, dlogits_lo=None):
    """out[0] += <g, w> of a fully connected layer = sum dlogits * (logits - bias)."""
    _cuda(logits, dlogits, bias, out, dlogits_lo)
    check(lib.evc_reg_cross(ptr(logits), ld_logits, ptr(dlogits), ptr(dlogits_lo), ld_dlogits, ptr(bias), B, N,
                            ptr(out), stream()), "evc_reg_cross")


def frames_pack(src, frame_idx, K, num_chunks, normalize, out_bf16=None, out_f32=None, out_lo=None):
    _cuda(src, frame_idx, out_bf16, out_f32, out_lo)
    B, T, D = src.shape
    per_batch = int(frame_idx is not None and frame_idx.dim() == 2)
    check(lib.evc_frames_pack(ptr(src), B, T, D, ptr(frame_idx), per_batch, K, num_chunks, int(normalize),
                              ptr(out_bf16), ptr(out_f32), ptr(out_lo), stream()), "evc_frames_pack")


def frames_pack_u8(src_u8, num_frames, frame_idx, K, num_chunks, normalize, out_bf16=None, out_f32=None, out_lo=None):
    """frames_pack on the quantised (uint8) features: Dequantize + zero padding fused in."""
    _cuda(src_u8, num_frames, frame_idx, out_bf16, out_f32, out_lo)
    assert src_u8.dtype == torch.uint8 and num_frames.dtype == torch.int32
    B, T, D = src_u8.shape
    per_batch = int(frame_idx is not None and frame_idx.dim() == 2)
    check(lib.evc_frames_pack_u8(ptr(src_u8), ptr(num_frames), B, T, D, ptr(frame_idx), per_batch, K, num_chunks,
                                 int(normalize), ptr(out_bf16), ptr(out_f32), ptr(out_lo), stream()), "evc_frames_pack_u8")


def num_frames_student(num_frames, every_n, max_frames=300, out=None):
    _cuda(num_frames)
    assert num_frames.dtype == torch.int32
    out = out if out is not None else torch.empty(num_frames.shape[0], dtype=torch.int64, device=num_frames.device)
    check(lib.evc_num_frames_student(ptr(num_frames), num_frames.shape[0], max_frames, every_n, ptr(out), stream()),
          "evc_num_frames_student")
    return out


def lstm_lengths(num_frames, num_chunks, chunk_len, len_l1=None, len_l2=None):
    _cuda(num_frames)
    B = num_frames.shape[0]
    assert num_frames.dtype in (torch.int32, torch.int64)
    dev = num_frames.device
    len_l1 = len_l1 if len_l1 is not None else torch.empty(num_chunks * B, dtype=torch.int32, device=dev)
    len_l2 = len_l2 if len_l2 is not None else torch.empty(B, dtype=torch.int32, device=dev)
    check(lib.evc_lstm_lengths(ptr(num_frames), int(num_frames.dtype == torch.int64), B, num_chunks, chunk_len,
                               ptr(len_l1), ptr(len_l2), stream()), "evc_lstm_lengths")
    return len_l1, len_l2


def random_frame_index(u, num_frames, out=None):
    _cuda(u, num_frames, out)
    B, K = u.shape
    idx = out if out is not None else torch.empty(B, K, dtype=torch.int32, device=u.device)
    check(lib.evc_random_frame_index(ptr(u), ptr(num_frames), B, K, ptr(idx), stream()), "evc_random_frame_index")
    return idx


def sampled_lengths(num_frames, K, out):
    """Sequence lengths of the randomly sampled student input: K for a video with frames, 0 for an empty one."""
    _cuda(num_frames, out)
    assert num_frames.dtype == torch.int32 and out.dtype == torch.int64
    check(lib.evc_sampled_lengths(ptr(num_frames), num_frames.shape[0], K, ptr(out), stream()), "evc_sampled_lengths")
    return out


def random_sequence_index(u, num_frames, K, out=None):
    _cuda(u, num_frames, out)
    B = u.shape[0]
    idx = out if out is not None else torch.empty(B, K, dtype=torch.int32, device=u.device)
    check(lib.evc_random_sequence_index(ptr(u), ptr(num_frames), B, K, ptr(idx), stream()),
          "evc_random_sequence_index")
    return idx


def random_uniform(out, seed: int, offset: int = 0):
    """Fill the f32 tensor `out` with U[0,1) draws of the counter-based Philox stream (seed, offset)."""
    _cuda(out)
    assert out.dtype == torch.float32
    check(lib.evc_random_uniform(int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), ptr(out), out.numel(),
                                 stream()), "evc_random_uniform")
    return out


def lstm_workspace_bytes(rows, H, Kx, precise=False):
    return int(lib.evc_lstm_workspace_bytes(rows, H, Kx, int(precise)))


def lstm_seq_fwd(x, x_step_stride, Kx, W, bias, rows, H, T, seq_len, h_all, c_all, gates_all, workspace=None,
                 t_begin=0, t_end=None, cuda_stream=None, x_lo=None, W_lo=None, h_lo_all=None, gates_lo_all=None):
    """All T steps of one BasicLSTMCell layer, or only steps [t_begin, t_end); on the current stream or on
    the raw `cuda_stream` handle.  *_lo: residual planes of the split-bf16 mode."""
    t_end = T if t_end is None else t_end
    check(lib.evc_lstm_seq_fwd_steps(ptr(x), x_step_stride, Kx, ptr(W), ptr(bias), rows, H, T, t_begin, t_end,
                                     ptr(seq_len), ptr(h_all), ptr(c_all), ptr(gates_all), ptr(workspace),
                                     workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                     ptr(x_lo), ptr(W_lo), ptr(h_lo_all), ptr(gates_lo_all),
                                     stream() if cuda_stream is None else cuda_stream), "evc_lstm_seq_fwd_steps")


def lstm_rec_workspace_bytes(rows, H, T):
    """Scratch of the resident-weights persistent recurrence for this shape; 0 = not eligible."""
    return int(lib.evc_lstm_rec_workspace_bytes(rows, H, T))


def lstm_seq_fwd_resident(x, x_step_stride, Kx, W, bias, rows, H, T, seq_len, h_all, c_all, gates_all, workspace):
    check(lib.evc_lstm_seq_fwd_resident(ptr(x), x_step_stride, Kx, ptr(W), ptr(bias), rows, H, T, ptr(seq_len),
                                        ptr(h_all), ptr(c_all), ptr(gates_all), ptr(workspace),
                                        workspace.numel() * workspace.element_size(), stream()),
          "evc_lstm_seq_fwd_resident")


def lstm_seq_bwd(W, Kx, rows, H, T, seq_len, gates_all, c_all, dh_ext_all, dh_final, ld_dh_final, dc_final,
                 ld_dc_final, dh_pass, dc, dz_all, workspace=None, dbias=None, W_lo=None, gates_lo_all=None,
                 dz_lo_all=None):
    check(lib.evc_lstm_seq_bwd(ptr(W), Kx, rows, H, T, ptr(seq_len), ptr(gates_all), ptr(c_all), ptr(dh_ext_all),
                               ptr(dh_final), ld_dh_final, ptr(dc_final), ld_dc_final, ptr(dh_pass), ptr(dc),
                               ptr(dz_all), ptr(dbias), ptr(workspace),
                               workspace.numel() * workspace.element_size() if workspace is not None else 0,
                               ptr(W_lo), ptr(gates_lo_all), ptr(dz_lo_all), stream()), "evc_lstm_seq_bwd")


def state_pack(c0, h0, c1, h1, rows, H, out_bf16=None, out_f32=None, h0_lo=None, h1_lo=None, out_lo=None):
    check(lib.evc_state_pack(ptr(c0), ptr(h0), ptr(c1), ptr(h1), rows, H, ptr(out_bf16), ptr(out_f32), ptr(h0_lo),
                             ptr(h1_lo), ptr(out_lo), stream()), "evc_state_pack")


def cast_bf16(src, dst, rows, cols, ld, dst_lo=None):
    check(lib.evc_cast_bf16(ptr(src), rows, cols, ld, ptr(dst), ptr(dst_lo), stream()), "evc_cast_bf16")


def fill_f32(t, value=0.0):
    check(lib.evc_fill_f32(ptr(t), t.numel(), float(value), stream()), "evc_fill_f32")


def moe_mix_fwd(G, ldg, E, lde, B, V, M, p_out):
    check(lib.evc_moe_mix_fwd(ptr(G), ldg, ptr(E), lde, B, V, M, ptr(p_out), stream()), "evc_moe_mix_fwd")


def moe_mix_bwd(G, ldg, E, lde, dP, B, V, M, dG, lddg, dE, ldde, dG_lo=None, dE_lo=None):
    check(lib.evc_moe_mix_bwd(ptr(G), ldg, ptr(E), lde, ptr(dP), B, V, M, ptr(dG), lddg, ptr(dE), ldde, ptr(dG_lo),
                              ptr(dE_lo), stream()), "evc_moe_mix_bwd")


def ce_kl_loss(P, PT, labels, ce_scale, kl_scale, ce_rows, kl_rows, dP):
    B, V = P.shape
    check(lib.evc_ce_kl_loss(ptr(P), ptr(PT), ptr(labels), B, V, ce_scale, kl_scale, ptr(ce_rows), ptr(kl_rows),
                             ptr(dP), stream()), "evc_ce_kl_loss")


def moe_mix_loss(G, ldg, E, lde, PT, labels, B, V, M, ce_scale, kl_scale, P, ce_rows, kl_rows, dG, lddg, dE, ldde,
                 dG_lo=None, dE_lo=None):
    check(lib.evc_moe_mix_loss(ptr(G), ldg, ptr(E), lde, ptr(PT), ptr(labels), B, V, M, ce_scale, kl_scale, ptr(P),
                               ptr(ce_rows), ptr(kl_rows), ptr(dG), lddg, ptr(dE), ldde, ptr(dG_lo), ptr(dE_lo),
                               stream()), "evc_moe_mix_loss")


def reduce_rows(rows, scale, out):
    check(lib.evc_reduce_rows(ptr(rows), rows.numel(), scale, ptr(out), stream()), "evc_reduce_rows")


def adam_lr(step, lr, beta1, beta2, lr_t):
    check(lib.evc_adam_lr(ptr(step), lr, beta1, beta2, ptr(lr_t), stream()), "evc_adam_lr")


def rep_loss(t_state, s_state, grad_scale, rows, d_student):
    B, S = s_state.shape
    check(lib.evc_rep_loss(ptr(t_state), ptr(s_state), B, S, grad_scale, ptr(rows), ptr(d_student), stream()),
          "evc_rep_loss")


def colsum_bf16(X, rows, N, ld, out):
    check(lib.evc_colsum_bf16(ptr(X), rows, N, ld, ptr(out), stream()), "evc_colsum_bf16")


def sumsq(g, w, weight_decay, out, out_wsq=None):
    check(lib.evc_sumsq(ptr(g), ptr(w), weight_decay, g.numel(), ptr(out), ptr(out_wsq), stream()), "evc_sumsq")


def clip_adam(w, g, m, v, normsq, clip_norm, weight_decay, lr_t, beta1, beta2, eps, shadow=None, cols=0,
              ld_shadow=0, shadow_lo=None, normsq_fused=None, reg_cross=None, reg_wsq=None, wsq_out=None):
    """normsq_fused / reg_cross / reg_wsq / wsq_out: the norm assembled from its parts (evc_clip_adam_fused)."""
    if normsq_fused is not None or reg_cross is not None or reg_wsq is not None or wsq_out is not None:
        check(lib.evc_clip_adam_fused(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), ptr(normsq), clip_norm, weight_decay,
                                      ptr(lr_t), beta1, beta2, eps, ptr(shadow), cols, ld_shadow, ptr(shadow_lo),
                                      ptr(normsq_fused), ptr(reg_cross), ptr(reg_wsq), ptr(wsq_out), stream()),
              "evc_clip_adam_fused")
        return
    check(lib.evc_clip_adam(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), ptr(normsq), clip_norm, weight_decay,
                            ptr(lr_t), beta1, beta2, eps, ptr(shadow), cols, ld_shadow, ptr(shadow_lo), stream()),
          "evc_clip_adam")


def topk(P, k, labels=None):
    """eval_util.top_k_triplets on the device: (idx int32 [B,k], val f32 [B,k], lab u8 [B,k] | None)."""
    _cuda(P, labels)
    if k <= 0:
        raise ValueError("k must be a positive integer.")  # eval_util.py:103-104
    B, V = P.shape
    k = min(k, V)
    idx = torch.empty(B, k, dtype=torch.int32, device=P.device)
    val = torch.empty(B, k, dtype=torch.float32, device=P.device)
    lab = torch.empty(B, k, dtype=torch.uint8, device=P.device) if labels is not None else None
    if B > 0:
        check(lib.evc_topk(ptr(P), B, V, k, ptr(labels), ptr(idx), ptr(val), ptr(lab), stream()), "evc_topk")
    return idx, val, lab


class BatchMetrics:
    """Device-side hit@1 / PERR / GAP of one batch (eval_util.py:17-79) with reusable scratch; optionally the
    epoch accumulators of EvaluationMetrics (class positives, sums).  `run` returns the device vector
    [hit@1, PERR, GAP, mean loss] (no host synchronisation) and the top-k triplets."""

    def __init__(self, batch, vocab, k, device, accumulate=False):
        self.B, self.V, self.k = batch, vocab, min(k, vocab)
        self.perr_rows = torch.zeros(batch, dtype=torch.float32, device=device)
        self.npos_rows = torch.zeros(batch, dtype=torch.int32, device=device)
        self.acc = torch.zeros(1, dtype=torch.float64, device=device)
        self.out = torch.zeros(4, dtype=torch.float32, device=device)
        self.class_pos = torch.zeros(vocab, dtype=torch.int32, device=device) if accumulate else None
        self.sums = torch.zeros(4, dtype=torch.float64, device=device) if accumulate else None

    def run(self, P, labels_u8, loss_rows=None, n=None):
        """P f32 [>=n, V], labels u8 [>=n, V]; the first n rows (default: all) are the batch."""
        _cuda(P, labels_u8, loss_rows)
        n = P.shape[0] if n is None else n
        if n > self.B or P.shape[1] != self.V:
            raise ValueError("batch larger than the scratch buffers / wrong number of classes")
        idx, val, lab = topk(P[:n], self.k, labels_u8[:n])
        if n > 0:
            check(lib.evc_batch_metrics(ptr(P), ptr(labels_u8), n, self.V, self.k, ptr(idx), ptr(val), ptr(lab),
                                        ptr(loss_rows), ptr(self.perr_rows), ptr(self.npos_rows), ptr(self.class_pos),
                                        ptr(self.acc), ptr(self.out), ptr(self.sums), stream()), "evc_batch_metrics")
        return self.out, idx, val, lab
