"""Execution plan of one Hierarchical-LSTM model instance on one B200.

Restates frame_level_models.py:200-338 (HierarchicalLstmModel.create_model /
create_model_inference) + video_level_models.py:397-448 (MoeModel) as a fixed sequence of
libevc kernel launches over pre-allocated HBM buffers:

  frames_pack -> [RNN_L1 cell0, cell1 over all chunks as one batch] -> state_pack
              -> [RNN_L2 cell0, cell1] -> state_pack -> gates/experts GEMMs -> mixture

The `num_chunks` lower-level dynamic_rnn calls of the reference share weights and start
from a zero state, so they run as one batch of num_chunks*B rows.  The two
cells of a MultiRNNCell are evaluated layer by layer (cell 1 reads the whole h-sequence
of cell 0), which lets every time step be one fused GEMM+cell kernel.

Memory layout (all row-major, per model instance, B = batch, C = chunks, l = frames/chunk,
R1 = C*B rows at the lower level):
  x      bf16 [l][R1][D]          row = chunk*B + b      (time-major inside a chunk)
  h_all  bf16 [T+1][rows][H]      c_all f32 [T+1][rows][H]   gates bf16 [T][rows][4H]
  l2_in  bf16 [C][B][4H]          = final [c0|h0|c1|h1] of every chunk (RNN_L2 input)
  state  f32  [B][4H]             final MultiRNNCell state of RNN_L2 (MoE input, L_REP operand)
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from .params import HLstmParams

BF16 = torch.bfloat16

# Stream-level overlap (bit mask, EVC_OVERLAP): 1 = the student model on its own stream (steps.py),
# 2 = the two cells of RNN_L1 interleaved step by step on two streams, 4 = weight-gradient GEMMs on a
# side stream next to the backward recurrence of the layer below, 8 = (one GPU) clip+Adam of the classifier
# variables on an optimizer stream next to the LSTM backward.
OVERLAP_STUDENT, OVERLAP_CELLS, OVERLAP_WGRAD, OVERLAP_OPTIMIZER, OVERLAP_CELLS_L2 = 1, 2, 4, 8, 16


# The resident-weights persistent recurrence (csrc/evc_rec.cuh) needs all its CTAs co-resident; two such grids
# launched on different streams could each hold part of the SMs and wait for the rest forever.  Every launch
# therefore waits for the previous one of the process (any stream) through this event.
_REC_TOKEN = {}


def resident_mode() -> bool:
    """EVC_RESIDENT=0 switches the persistent recurrence off (per-step split-K GEMM + cell kernel instead)."""
    return os.environ.get("EVC_RESIDENT", "1") != "0"


def overlap_mode() -> int:
    """Default 7.  Bit 8 measured no gain on the joint step and -4 % on the cfg #4 fine-tune step (the
    optimizer's HBM traffic slows the power-capped GEMMs it runs next to); bit 16 = the same interleaving
    for RNN_L2's small-row steps (second split-K scratch buffer)."""
    return int(os.environ.get("EVC_OVERLAP", "7"))


class _Layer:
    """Buffers of one BasicLSTMCell unrolled over T steps for `rows` sequences."""

    def __init__(self, rows, H, T, training, dev, precise=False):
        self.rows, self.H, self.T = rows, H, T
        self.h_all = torch.zeros(T + 1, rows, H, dtype=BF16, device=dev)
        self.c_all = torch.zeros(T + 1, rows, H, dtype=torch.float32, device=dev)
        self.gates = torch.empty(T, rows, 4 * H, dtype=BF16, device=dev) if training else None
        self.dz = torch.empty(T, rows, 4 * H, dtype=BF16, device=dev) if training else None
        # residual planes of the split-bf16 mode (None otherwise)
        self.h_lo = torch.zeros(T + 1, rows, H, dtype=BF16, device=dev) if precise else None
        self.gates_lo = torch.empty(T, rows, 4 * H, dtype=BF16, device=dev) if (precise and training) else None
        self.dz_lo = torch.empty(T, rows, 4 * H, dtype=BF16, device=dev) if (precise and training) else None


class HLstmEngine:
    def __init__(self, params: HLstmParams, batch: int, num_frames_in: int, num_chunks: int,
                 training: bool = True):
        cfg = params.cfg
        if num_frames_in % num_chunks != 0:
            raise ValueError("tf.split: number of frames must be divisible by the number of chunks")
        self.p, self.cfg = params, cfg
        self.B, self.K, self.C = batch, num_frames_in, num_chunks
        self.ell = num_frames_in // num_chunks
        self.training = training
        self.precise = bool(getattr(params, "precise", False))
        self.overlap = overlap_mode()
        self._dp_all, self._cls_stream, self._cls_done, self._cls_pending = None, None, None, False
        self._fused_norms = False    # inside lstm_backward: weight-gradient GEMMs also leave |dW|^2 (params.norm_aux)
        self._side: Optional[torch.cuda.Stream] = None
        self._events = []
        dev = params.device
        H, D, V, M, S = cfg.lstm_cells, cfg.feature_size, cfg.vocab_size, cfg.num_mixtures, cfg.state_size
        self.R1 = self.C * self.B
        B, R1, ell, C = self.B, self.R1, self.ell, self.C
        px = self.precise
        self.x = torch.empty(ell, R1, D, dtype=BF16, device=dev)
        self.x_lo = torch.empty(ell, R1, D, dtype=BF16, device=dev) if px else None
        self.len_l1 = torch.zeros(R1, dtype=torch.int32, device=dev)
        self.len_l2 = torch.zeros(B, dtype=torch.int32, device=dev)
        self.l1 = [_Layer(R1, H, ell, training, dev, px) for _ in range(2)]
        self.l2_in = torch.empty(C, B, S, dtype=BF16, device=dev)
        self.l2_in_lo = torch.empty(C, B, S, dtype=BF16, device=dev) if px else None
        self.l2 = [_Layer(B, H, C, training, dev, px) for _ in range(2)]
        self.state = torch.empty(B, S, dtype=torch.float32, device=dev)
        self.state_bf16 = torch.empty(B, S, dtype=BF16, device=dev)
        self.state_lo = torch.empty(B, S, dtype=BF16, device=dev) if px else None
        self.ldg, self.lde = V * (M + 1), V * M
        self.G = torch.empty(B, self.ldg, dtype=torch.float32, device=dev)
        self.E = torch.empty(B, self.lde, dtype=torch.float32, device=dev)
        self.pred = torch.empty(B, V, dtype=torch.float32, device=dev)
        # split-K partial slabs of the recurrence steps (one scratch buffer shared by all cells)
        ws = max(ops.lstm_workspace_bytes(R1, H, D, px), ops.lstm_workspace_bytes(R1, H, H, px),
                 ops.lstm_workspace_bytes(B, H, S, px), ops.lstm_workspace_bytes(B, H, H, px))
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=dev)
        self.workspace2 = (torch.empty(ops.lstm_workspace_bytes(B, H, H), dtype=torch.uint8, device=dev)
                           if self.overlap & OVERLAP_CELLS_L2 else None)
        # scratch of the resident-weights persistent recurrence (hoisted input projection Zx, slice layout of Wh,
        # step flags) for the cells whose row count is eligible: RNN_L2 always, RNN_L1 up to ~2048 rows
        rec = (max(ops.lstm_rec_workspace_bytes(R1, H, ell), ops.lstm_rec_workspace_bytes(B, H, C))
               if resident_mode() and not px else 0)
        self.rec_ws = None
        if rec:
            raw = torch.empty(rec + 1024, dtype=torch.uint8, device=dev)
            skip = (-raw.data_ptr()) % 1024
            self.rec_ws = raw[skip:skip + rec]
        if training:
            self.lddg, self.ldde = ops.pad8(self.ldg, 64), ops.pad8(self.lde, 64)
            self.dG = torch.zeros(B, self.lddg, dtype=BF16, device=dev)
            self.dE = torch.zeros(B, self.ldde, dtype=BF16, device=dev)
            self.dG_lo = torch.zeros(B, self.lddg, dtype=BF16, device=dev) if px else None
            self.dE_lo = torch.zeros(B, self.ldde, dtype=BF16, device=dev) if px else None
            self.dP = torch.empty(B, V, dtype=torch.float32, device=dev)
            self.dstate = torch.zeros(B, S, dtype=torch.float32, device=dev)
            self.dx_l2 = torch.empty(C * B, H, dtype=torch.float32, device=dev)       # d h0 sequence of RNN_L2
            self.dl2_in = torch.empty(R1, S, dtype=torch.float32, device=dev)         # d final state of every chunk
            self.dx_l1 = torch.empty(ell * R1, H, dtype=torch.float32, device=dev)    # d h0 sequence of RNN_L1
            self.scr_l1 = (torch.empty(R1, H, dtype=torch.float32, device=dev),
                           torch.empty(R1, H, dtype=torch.float32, device=dev))
            self.scr_l2 = (torch.empty(B, H, dtype=torch.float32, device=dev),
                           torch.empty(B, H, dtype=torch.float32, device=dev))

    # ------------------------------------------------------------------ forward
    def _side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.p.device, priority=int(os.environ.get("EVC_PRIO_SIDE", "0")))
        return self._side

    def _event(self, i: int) -> torch.cuda.Event:
        while len(self._events) <= i:
            self._events.append(torch.cuda.Event())
        return self._events[i]

    def _cell_fwd(self, layer: _Layer, x, x_stride, Kx, level, cell, seq_len, t_begin=0, t_end=None,
                  cuda_stream=None, workspace=None, x_lo=None):
        p = self.p
        if self.precise:
            # split-bf16 mode: every step = 3-segment split-K GEMM into f32 slabs + the cell kernel
            k = p.kernel(level, cell)
            ops.lstm_seq_fwd(x, x_stride, Kx, p.shadow[k], p.w[p.bias(level, cell)], layer.rows, layer.H, layer.T,
                             seq_len, layer.h_all, layer.c_all, layer.gates, self.workspace, t_begin, t_end,
                             cuda_stream, x_lo=x_lo, W_lo=p.shadow_lo[k], h_lo_all=layer.h_lo,
                             gates_lo_all=layer.gates_lo)
            return
        if (t_begin == 0 and t_end is None and cuda_stream is None and self._resident_ok(layer)
                and x_stride == layer.rows * Kx):
            dev, cur = p.device, torch.cuda.current_stream()
            prev = _REC_TOKEN.get(dev)
            if prev is not None:
                cur.wait_event(prev)
            ops.lstm_seq_fwd_resident(x, x_stride, Kx, p.shadow[p.kernel(level, cell)], p.w[p.bias(level, cell)],
                                      layer.rows, layer.H, layer.T, seq_len, layer.h_all, layer.c_all, layer.gates,
                                      self.rec_ws)
            ev = prev if prev is not None else torch.cuda.Event()
            ev.record(cur)
            _REC_TOKEN[dev] = ev
            return
        ops.lstm_seq_fwd(x, x_stride, Kx, p.shadow[p.kernel(level, cell)], p.w[p.bias(level, cell)],
                         layer.rows, layer.H, layer.T, seq_len, layer.h_all, layer.c_all, layer.gates,
                         self.workspace if workspace is None else workspace, t_begin, t_end, cuda_stream)

    def _resident_ok(self, layer: _Layer) -> bool:
        """Use the persistent resident-weights kernel where it is measured faster than the per-step path
        (profiles/r02_resident_recurrence.md): one 128-row tile per CTA (rows <= 256 at H = 1024) and enough steps
        to amortise the one-off slice load + hoisted projection; EVC_RESIDENT=2 forces it wherever it is eligible."""
        if self.rec_ws is None:
            return False
        need = ops.lstm_rec_workspace_bytes(layer.rows, layer.H, layer.T)
        if not 0 < need <= self.rec_ws.numel():
            return False
        if os.environ.get("EVC_RESIDENT", "1") == "2":
            return True
        slices = layer.H // 16
        groups = max(1, min(148 // slices, (layer.rows + 127) // 128))
        return (layer.rows + 127) // 128 <= groups and layer.T >= 8

    def _cells_interleaved(self, a: _Layer, b: _Layer, x, x_stride, Kx, level, seq_len, workspace_b=None):
        """MultiRNNCell wavefront: cell 1 step t (side stream) next to cell 0 step t+1 (current stream)."""
        main, side = torch.cuda.current_stream(), self._side_stream()
        h0_seq = a.h_all[1:]
        H = a.H
        for t in range(a.T):
            self._cell_fwd(a, x, x_stride, Kx, level, 0, seq_len, t, t + 1, main.cuda_stream)
            ev = self._event(t)
            ev.record(main)
            side.wait_event(ev)
            self._cell_fwd(b, h0_seq, a.rows * H, H, level, 1, seq_len, t, t + 1, side.cuda_stream, workspace_b)
        main.wait_stream(side)

    def forward(self, src: torch.Tensor, frame_idx: Optional[torch.Tensor], normalize: bool,
                num_frames: torch.Tensor, raw_num_frames: Optional[torch.Tensor] = None, mix: bool = True) -> None:
        """src f32 (or uint8, quantised) [B, T_src, D]; frame_idx int32 [K] / [B,K] / None; num_frames
        int32|int64 [B] (already the *sampled* count for the student); raw_num_frames int32 [B] is the
        unsampled count (uint8 sources only).  Fills self.state and self.pred."""
        self.forward_lstm(src, frame_idx, normalize, num_frames, raw_num_frames)
        self.classifier_forward(mix)

    def forward_lstm(self, src: torch.Tensor, frame_idx: Optional[torch.Tensor], normalize: bool,
                     num_frames: torch.Tensor, raw_num_frames: Optional[torch.Tensor] = None) -> None:
        """Both LSTM levels: fills self.state (f32) and self.state_bf16."""
        cfg = self.cfg
        H, D, S = cfg.lstm_cells, cfg.feature_size, cfg.state_size
        B, R1, ell, C = self.B, self.R1, self.ell, self.C
        if src.shape[0] != B or src.shape[2] != D:
            raise ValueError(f"model_input shape {tuple(src.shape)} does not match the plan [{B},*,{D}]")
        if src.dtype == torch.uint8:
            # quantised tfrecord features: Dequantize + zero padding fused into the pack kernel
            if raw_num_frames is None:
                raise ValueError("uint8 features need the unsampled num_frames (int32) for zero padding")
            ops.frames_pack_u8(src, raw_num_frames, frame_idx, self.K, C, normalize, out_bf16=self.x, out_lo=self.x_lo)
        else:
            ops.frames_pack(src, frame_idx, self.K, C, normalize, out_bf16=self.x, out_lo=self.x_lo)
        ops.lstm_lengths(num_frames, C, ell, self.len_l1, self.len_l2)
        a, b = self.l1
        if self.precise:
            self._cell_fwd(a, self.x, R1 * D, D, 0, 0, self.len_l1, x_lo=self.x_lo)
            self._cell_fwd(b, a.h_all[1:], R1 * H, H, 0, 1, self.len_l1, x_lo=a.h_lo[1:])
            ops.state_pack(a.c_all[ell], a.h_all[ell], b.c_all[ell], b.h_all[ell], R1, H, out_bf16=self.l2_in,
                           h0_lo=a.h_lo[ell], h1_lo=b.h_lo[ell], out_lo=self.l2_in_lo)
            a2, b2 = self.l2
            self._cell_fwd(a2, self.l2_in, B * S, S, 1, 0, self.len_l2, x_lo=self.l2_in_lo)
            self._cell_fwd(b2, a2.h_all[1:], B * H, H, 1, 1, self.len_l2, x_lo=a2.h_lo[1:])
            ops.state_pack(a2.c_all[C], a2.h_all[C], b2.c_all[C], b2.h_all[C], B, H, out_bf16=self.state_bf16,
                           out_f32=self.state, h0_lo=a2.h_lo[C], h1_lo=b2.h_lo[C], out_lo=self.state_lo)
            return
        if (self.overlap & OVERLAP_CELLS) and R1 > 1024 and not self._resident_ok(a):
            # MultiRNNCell wavefront: cell 1 step t next to cell 0 step t+1 (fused-epilogue steps only: the
            # split-K path of the small-row steps shares one scratch buffer)
            self._cells_interleaved(a, b, self.x, R1 * D, D, 0, self.len_l1)
        else:
            self._cell_fwd(a, self.x, R1 * D, D, 0, 0, self.len_l1)
            self._cell_fwd(b, a.h_all[1:], R1 * H, H, 0, 1, self.len_l1)
        ops.state_pack(a.c_all[ell], a.h_all[ell], b.c_all[ell], b.h_all[ell], R1, H, out_bf16=self.l2_in)
        a2, b2 = self.l2
        if self.workspace2 is not None and B <= 1024 and not self._resident_ok(a2):
            self._cells_interleaved(a2, b2, self.l2_in, B * S, S, 1, self.len_l2, self.workspace2)
        else:
            self._cell_fwd(a2, self.l2_in, B * S, S, 1, 0, self.len_l2)
            self._cell_fwd(b2, a2.h_all[1:], B * H, H, 1, 1, self.len_l2)
        ops.state_pack(a2.c_all[C], a2.h_all[C], b2.c_all[C], b2.h_all[C], B, H,
                       out_bf16=self.state_bf16, out_f32=self.state)

    def classifier_forward(self, mix: bool = True) -> None:
        """MoeModel on self.state_bf16 -> self.pred (video_level_models.py:423-447).  mix=False stops after
        the two logit GEMMs (the training step applies the mixture inside its fused loss kernel)."""
        p, cfg = self.p, self.cfg
        S, V, M = cfg.state_size, cfg.vocab_size, cfg.num_mixtures
        lo = p.shadow_lo.get if self.precise else (lambda n: None)
        ops.gemm(self.state_bf16, p.shadow[p.gates_w], self.B, self.ldg, S, self.G, b_mn=True,
                 A_lo=self.state_lo, B_lo=lo(p.gates_w))
        ops.gemm(self.state_bf16, p.shadow[p.experts_w], self.B, self.lde, S, self.E, b_mn=True,
                 bias=p.w[p.experts_b], A_lo=self.state_lo, B_lo=lo(p.experts_w))
        if mix:
            ops.moe_mix_fwd(self.G, self.ldg, self.E, self.lde, self.B, V, M, self.pred)

    def classifier_loss_fused(self, labels_u8, teacher_pred, ce_scale, kl_scale, ce_rows, kl_rows) -> None:
        """Mixture + CrossEntropyLoss rows + KL(teacher||student) rows + gradients w.r.t. the logits (self.dG,
        self.dE) in one launch; fills self.pred.  Follow with classifier_backward(None, ..., logits_done=True)."""
        cfg = self.cfg
        ops.moe_mix_loss(self.G, self.ldg, self.E, self.lde, teacher_pred, labels_u8, self.B, cfg.vocab_size,
                         cfg.num_mixtures, ce_scale, kl_scale, self.pred, ce_rows, kl_rows, self.dG, self.lddg,
                         self.dE, self.ldde, self.dG_lo, self.dE_lo)

    # ------------------------------------------------------------------ backward
    def _cell_bwd(self, layer: _Layer, level, cell, Kx, seq_len, dh_ext, dfinal, col, scratch):
        """dfinal: f32 [rows, 4H] gradient of the final [c0|h0|c1|h1]; this cell owns columns
        [col, col+H) (c) and [col+H, col+2H) (h)."""
        p = self.p
        H = layer.H
        ld = dfinal.stride(0)
        # (the bias gradient -- column sums of dz -- is accumulated by the cell kernel while it writes dz)
        ops.lstm_seq_bwd(p.shadow[p.kernel(level, cell)], Kx, layer.rows, H, layer.T, seq_len, layer.gates,
                         layer.c_all, dh_ext, dfinal[:, col + H:], ld, dfinal[:, col:], ld,
                         scratch[0], scratch[1], layer.dz, self.workspace, dbias=p.g[p.bias(level, cell)],
                         W_lo=p.shadow_lo.get(p.kernel(level, cell)), gates_lo_all=layer.gates_lo,
                         dz_lo_all=layer.dz_lo)

    def _cell_wgrad(self, layer: _Layer, level, cell, x2d, Kx, x2d_lo=None):
        """dW = [x | h_prev]^T dz over all steps and rows (db comes from the backward recurrence, _cell_bwd)."""
        p = self.p
        H, R = layer.H, layer.T * layer.rows
        dz = layer.dz.view(R, 4 * H)
        dz_lo = layer.dz_lo.view(R, 4 * H) if self.precise else None
        h_lo = layer.h_lo.view(-1, H) if self.precise else None
        gW = p.g[p.kernel(level, cell)]
        if self._fused_norms:
            # |dW|^2 accumulated by the two GEMMs' epilogues (the clip of slim's train op is per variable)
            ss = p.norm_aux[(level * 2 + cell) * 2, 0:1]
            ops.gemm_wgrad(x2d, dz, Kx, 4 * H, R, gW[:Kx], ss, a_mn=True, b_mn=True, lda=Kx, ldc=4 * H,
                           A_lo=x2d_lo, B_lo=dz_lo)
            ops.gemm_wgrad(layer.h_all.view(-1, H), dz, H, 4 * H, R, gW[Kx:], ss, a_mn=True, b_mn=True, lda=H,
                           ldc=4 * H, A_lo=h_lo, B_lo=dz_lo)
            return
        ops.gemm(x2d, dz, Kx, 4 * H, R, gW[:Kx], a_mn=True, b_mn=True, lda=Kx, ldc=4 * H, A_lo=x2d_lo, B_lo=dz_lo)
        ops.gemm(layer.h_all.view(-1, H), dz, H, 4 * H, R, gW[Kx:], a_mn=True, b_mn=True, lda=H, ldc=4 * H,
                 A_lo=h_lo, B_lo=dz_lo)

    def _cell_dx(self, layer: _Layer, level, cell, Kx, out):
        """d input sequence = dz @ Wx^T  (Wx = first Kx rows of the kernel)."""
        p = self.p
        H, R = layer.H, layer.T * layer.rows
        ops.gemm(layer.dz.view(R, 4 * H), p.shadow[p.kernel(level, cell)], R, Kx, 4 * H, out, ldb=4 * H,
                 A_lo=layer.dz_lo.view(R, 4 * H) if self.precise else None,
                 B_lo=p.shadow_lo.get(p.kernel(level, cell)))

    def backward(self, dP: torch.Tensor, dstate_preset: bool = False) -> None:
        """Gradients of every weight into params.g given dP = dLoss/dpredictions [B,V].
        If dstate_preset, self.dstate already holds dLoss/dstate from other consumers of the
        state (L_REP) and the classifier's contribution is added to it."""
        self.classifier_backward(dP, dstate_preset)
        self.lstm_backward()

    def classifier_backward(self, dP: Optional[torch.Tensor], dstate_preset: bool = False,
                            logits_done: bool = False, dp=None) -> None:
        """MoE backward: weight gradients into params.g, d(state) accumulated into self.dstate.
        logits_done: self.dG/self.dE were already produced by classifier_loss_fused.
        dp = (rank, world, group): sharded data parallelism -- see `_classifier_wgrad_gathered`; the two weight
        gradients are then complete (averaged, this rank's row block only) after `wait_classifier_wgrad()`."""
        p, cfg = self.p, self.cfg
        S, V, M, B = cfg.state_size, cfg.vocab_size, cfg.num_mixtures, self.B
        if not logits_done:
            ops.moe_mix_bwd(self.G, self.ldg, self.E, self.lde, dP, B, V, M, self.dG, self.lddg, self.dE, self.ldde,
                            self.dG_lo, self.dE_lo)
        gathers = self._classifier_gather(dp) if dp is not None else None
        if not dstate_preset:
            ops.fill_f32(self.dstate, 0.0)
        lo = p.shadow_lo.get if self.precise else (lambda n: None)
        # split-K so that the work items fill ONE round over the CTA pairs (B = 256, S = 4096: 16 pair tiles x 4 splits
        # on 74 pairs; 8 splits took a second, partly empty round: 47 -> 45 us and 39 -> 35 us, profiles/r02_exp_moe_gemm.txt)
        tm, tn = -(-B // 128), -(-S // 256)
        slots = 74 if tm >= 2 else 148
        sk = max(1, min(8, slots // (-(-tm // 2) * tn if tm >= 2 else tn)))
        ops.gemm(self.dG, p.shadow[p.gates_w], B, S, self.ldg, self.dstate, split_k=sk, accumulate=True,
                 A_lo=self.dG_lo, B_lo=lo(p.gates_w))
        ops.gemm(self.dE, p.shadow[p.experts_w], B, S, self.lde, self.dstate, split_k=sk, accumulate=True,
                 A_lo=self.dE_lo, B_lo=lo(p.experts_w))
        if dp is not None:
            self._classifier_wgrad_gathered(dp, gathers)
        elif p.fused_norms():
            # the two matrices' |g|^2 from the GEMM epilogues and <g, w> from the logits (the regulariser's
            # gradient wd*w enters the clipped norm, train.py:324-334): no sumsq pass over g and w in the optimizer
            p.begin_fused_norms(8, 11)
            ops.gemm_wgrad(self.state_bf16, self.dG, S, self.ldg, B, p.g[p.gates_w], p.norm_aux[8, 0:1], a_mn=True,
                           b_mn=True, A_lo=self.state_lo, B_lo=self.dG_lo)
            ops.gemm_wgrad(self.state_bf16, self.dE, S, self.lde, B, p.g[p.experts_w], p.norm_aux[9, 0:1], a_mn=True,
                           b_mn=True, A_lo=self.state_lo, B_lo=self.dE_lo)
            ops.reg_cross(self.G, self.G.stride(0), self.dG, self.lddg, None, B, self.ldg, p.norm_aux[8, 1:2],
                          self.dG_lo)
            ops.reg_cross(self.E, self.E.stride(0), self.dE, self.ldde, p.w[p.experts_b], B, self.lde,
                          p.norm_aux[9, 1:2], self.dE_lo)
            p.end_fused_norms((8, 9))
        else:
            ops.gemm(self.state_bf16, self.dG, S, self.ldg, B, p.g[p.gates_w], a_mn=True, b_mn=True,
                     A_lo=self.state_lo, B_lo=self.dG_lo)
            ops.gemm(self.state_bf16, self.dE, S, self.lde, B, p.g[p.experts_w], a_mn=True, b_mn=True,
                     A_lo=self.state_lo, B_lo=self.dE_lo)
        gbe = p.g[p.experts_b]
        ops.fill_f32(gbe, 0.0)
        ops.colsum_bf16(self.dE, B, self.lde, self.ldde, gbe)
        if self.precise:
            ops.colsum_bf16(self.dE_lo, B, self.lde, self.ldde, gbe)      # accumulates: hi + lo

    # ---- sharded data parallelism: the classifier's weight gradients without a gradient collective.
    # g = X^T dL is a rank-B product (X = the final state [B, S], dL = gradient w.r.t. the logits [B, N], both bf16 as
    # the weight-gradient GEMM reads them), so the average over ranks (1/world) sum_r X_r^T dL_r is ONE contraction over
    # the gathered batch: every rank all-gathers X and dL (B x (S + 3 V M) bf16 = 9 MB per rank instead of reduce-
    # scattering 386 MB of f32 gradients per model) and computes only the row block of g its optimizer shard owns --
    # the FLOPs of the local weight-gradient GEMM, 1/world of its store traffic, f32 accumulation over all world x B
    # rows inside one GEMM.  Same averaged gradient as reduce_scatter(AVG) of the per-rank products.
    def _classifier_gather(self, dp):
        import torch.distributed as dist
        rank, world, group = dp
        srcs = [self.state_bf16, self.dG, self.dE]
        if self.precise:
            srcs += [self.state_lo, self.dG_lo, self.dE_lo]
        if self._dp_all is None or self._dp_all[0].shape[0] != world * self.B:
            self._dp_all = [torch.empty(world * self.B, t.shape[1], dtype=t.dtype, device=t.device) for t in srcs]
        return [dist.all_gather_into_tensor(dst, src, group=group, async_op=True)
                for dst, src in zip(self._dp_all, srcs)]

    def _classifier_wgrad_gathered(self, dp, gathers) -> None:
        """On a stream of its own (the gathers must not stall the LSTM backward that follows on the caller's stream);
        `wait_classifier_wgrad` joins it."""
        rank, world, group = dp
        p, cfg = self.p, self.cfg
        S, KB = cfg.state_size, world * self.B
        if self._cls_stream is None:
            self._cls_stream = torch.cuda.Stream(device=p.device)
            self._cls_done = torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._cls_stream.wait_stream(main)
        with torch.cuda.stream(self._cls_stream):
            for w in gathers:
                w.wait()
            X, dG, dE = self._dp_all[:3]
            Xl, dGl, dEl = self._dp_all[3:] if self.precise else (None, None, None)
            r0, r1 = p.row_block(p.gates_w, rank, world)
            ops.gemm_wgrad(X[:, r0:r1], dG, r1 - r0, self.ldg, KB, p.g[p.gates_w][r0:r1], None, a_mn=True, b_mn=True,
                           alpha=1.0 / world, A_lo=Xl[:, r0:r1] if Xl is not None else None, B_lo=dGl)
            r0, r1 = p.row_block(p.experts_w, rank, world)
            ops.gemm_wgrad(X[:, r0:r1], dE, r1 - r0, self.lde, KB, p.g[p.experts_w][r0:r1], None, a_mn=True, b_mn=True,
                           alpha=1.0 / world, A_lo=Xl[:, r0:r1] if Xl is not None else None, B_lo=dEl)
            self._cls_done.record(self._cls_stream)
        self._cls_pending = True

    def wait_classifier_wgrad(self) -> None:
        """Make the current stream wait for the gathered-batch weight gradients of the classifier (no-op otherwise)."""
        if self._cls_pending:
            torch.cuda.current_stream().wait_event(self._cls_done)
            self._cls_pending = False

    def lstm_backward(self) -> None:
        """Backward of both LSTM levels given self.dstate = dLoss/d(final state)."""
        p, cfg = self.p, self.cfg
        H, D, S = cfg.lstm_cells, cfg.feature_size, cfg.state_size
        B, R1, ell, C = self.B, self.R1, self.ell, self.C
        # ---- RNN_L2 (cell 1 first: its input gradient feeds cell 0)
        a2, b2 = self.l2
        a, b = self.l1
        self._fused_norms = p.fused_norms()
        if self._fused_norms:
            p.begin_fused_norms(0, 8)
        if self.overlap & OVERLAP_WGRAD:
            # dz of a cell is final once its recurrence is done: its weight-gradient GEMMs (tensor-bound, no
            # dependants until the optimizer) run on the side stream next to the latency-bound recurrence
            # of the cell below
            main, side = torch.cuda.current_stream(), self._side_stream()

            def wgrad_aside(i, *args):
                ev = self._event(i)
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self._cell_wgrad(*args)
        else:
            main = side = None

            def wgrad_aside(i, *args):
                self._cell_wgrad(*args)
        # ---- RNN_L2 (cell 1 first: its input gradient feeds cell 0)
        self._cell_bwd(b2, 1, 1, H, self.len_l2, None, self.dstate, 2 * H, self.scr_l2)
        px = self.precise
        wgrad_aside(0, b2, 1, 1, a2.h_all[1:].view(-1, H), H, a2.h_lo[1:].view(-1, H) if px else None)
        self._cell_dx(b2, 1, 1, H, self.dx_l2)
        self._cell_bwd(a2, 1, 0, S, self.len_l2, self.dx_l2, self.dstate, 0, self.scr_l2)
        wgrad_aside(1, a2, 1, 0, self.l2_in.view(-1, S), S, self.l2_in_lo.view(-1, S) if px else None)
        self._cell_dx(a2, 1, 0, S, self.dl2_in)
        # ---- RNN_L1: final-state gradient of chunk i = gradient of RNN_L2's input at step i
        self._cell_bwd(b, 0, 1, H, self.len_l1, None, self.dl2_in, 2 * H, self.scr_l1)
        wgrad_aside(2, b, 0, 1, a.h_all[1:].view(-1, H), H, a.h_lo[1:].view(-1, H) if px else None)
        self._cell_dx(b, 0, 1, H, self.dx_l1)
        self._cell_bwd(a, 0, 0, D, self.len_l1, self.dx_l1, self.dl2_in, 0, self.scr_l1)
        self._cell_wgrad(a, 0, 0, self.x.view(-1, D), D, self.x_lo.view(-1, D) if px else None)
        if side is not None:
            main.wait_stream(side)
        if self._fused_norms:
            p.end_fused_norms((0, 2, 4, 6))
            self._fused_norms = False
