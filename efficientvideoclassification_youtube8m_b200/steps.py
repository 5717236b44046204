"""Step functions: what the reference's ``build_graph`` + one ``sess.run`` do, as fixed launch
sequences over two :class:`HLstmEngine` plans.

  TeacherStudentTrainer.step   <- train.py:185-427 build_graph + train.py:516 sess.run
  StudentFinetuneTrainer.step  <- train_finetune.py:185-331 + :404
  TeacherStudentEvaluator      <- validate.py:109-189 (student predictions inside the T+S graph)
  StudentEvaluator             <- eval_finetune.py:108-175

Inputs are the raw (dequantised, un-normalised) frame features ``model_input_raw`` f32
[B,300,D], ``num_frames`` int32 [B] and ``labels`` bool/uint8 [B,V] on the device — what the
reference's reader queue hands to the graph.  Data parallelism: every rank runs the
same step on its own batch; gradients are averaged with one NCCL allreduce per model over the
flat gradient buffer before the per-variable clip + Adam.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import ops
from .engine import OVERLAP_OPTIMIZER, OVERLAP_STUDENT, HLstmEngine, overlap_mode
from .params import HLstmParams, ModelConfig

MAX_FRAMES = 300  # train.py:262
SAMPLERS = ("uniform", "random_frames", "random_sequence")


def uniform_frame_indices(every_n: int):
    """train.py:265-269."""
    out, k = [], 0
    while every_n * k <= MAX_FRAMES - 1:
        out.append(every_n * k)
        k += 1
    return out


def _as_u8(labels: torch.Tensor) -> torch.Tensor:
    if labels.dtype == torch.uint8:
        return labels
    if labels.dtype == torch.bool:
        return labels.view(torch.uint8)
    raise TypeError("labels must be a bool or uint8 tensor [B, vocab_size]")


class _Base:
    def __init__(self, cfg: ModelConfig, batch_size: int, device, every_n: int, num_inputs_L1: int,
                 base_learning_rate: float, clip_gradient_norm: float, regularization_penalty: float,
                 shard_optimizer: Optional[bool] = None, learning_rate_decay: float = 1.0,
                 learning_rate_decay_examples: float = 4000000.0, sampling: str = "uniform",
                 sampling_seed: int = 0):
        self.cfg, self.B, self.device = cfg, batch_size, torch.device(device)
        self.every_n, self.num_inputs_L1 = every_n, num_inputs_L1
        self.base_lr, self.clip, self.penalty = base_learning_rate, clip_gradient_norm, regularization_penalty
        self.lr_decay, self.lr_decay_examples = learning_rate_decay, learning_rate_decay_examples
        idx = uniform_frame_indices(every_n)
        if len(idx) % num_inputs_L1 != 0:
            raise ValueError(f"every_n={every_n}: {len(idx)} sampled frames do not split into "
                             f"{num_inputs_L1} chunks (tf.split would fail, SURVEY F13)")
        self.student_frames = len(idx)
        self.frame_idx = torch.tensor(idx, dtype=torch.int32, device=self.device)
        self.nf_student = torch.zeros(batch_size, dtype=torch.int64, device=self.device)
        # BASELINE config #5 "random vs uniform": the student's frames drawn by model_utils.SampleRandomFrames
        # ("random_frames": K independent frames per video, model_utils.py:39-58) or SampleRandomSequence
        # ("random_sequence": K consecutive frames from a random start, :11-36) instead of every n-th frame.
        if sampling not in SAMPLERS:
            raise ValueError(f"sampling must be one of {SAMPLERS}")
        self.sampling, self.sampling_seed, self._draws = sampling, int(sampling_seed), 0
        if sampling != "uniform":
            K = self.student_frames
            self.u = torch.empty((batch_size, K) if sampling == "random_frames" else (batch_size,),
                                 dtype=torch.float32, device=self.device)
            self.frame_idx_rand = torch.empty(batch_size, K, dtype=torch.int32, device=self.device)
        self.global_step = 0
        self._pending = []
        self._gathers = {}          # id(params) -> outstanding all-gathers of that model's bf16 operand rows
        self._opt_streams = {}
        # attribution experiments only (profiles/r02_dp_ablation.md): 1 = no gradient collectives, 2 = no operand
        # all-gathers -- the step then computes garbage across ranks but keeps its kernel schedule
        self._dp_ablate = int(os.environ.get("EVC_DP_ABLATE", "0"))
        # optimizer sharding over the data-parallel ranks (on by default when there is more than one)
        self.shard_optimizer = shard_optimizer if shard_optimizer is not None else (self._world() > 1)
        # The sharded optimizer completes its per-variable norms with two 11-float all-reduces.  On the default
        # communicator they would queue behind every gradient reduce-scatter issued so far (one NCCL stream per
        # communicator), i.e. the teacher's optimizer pass would wait for the student's gradients: they get their
        # own communicator.  (Collective: every rank constructs its trainers in the same order.)
        self.norm_group = None
        if self.shard_optimizer and self._world() > 1 and dist.get_backend() == "nccl" \
                and os.environ.get("EVC_NORM_GROUP", "1") != "0":
            self.norm_group = dist.new_group()

    @staticmethod
    def _world():
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _student_sample(self, num_frames, u=None):
        """The student's frame indices and sequence lengths for this batch, on the current stream.
        uniform: every_n-th frame and the float64 length rule (train.py:262-272).
        random_*: indices from the reference's samplers given U[0,1) draws `u` (parity: pass them; otherwise
        the library's Philox stream, counter advancing per step); every sampled frame is a real frame of the
        video (the samplers index inside [0, num_frames)), so the student sees K valid steps -- K * [n > 0]."""
        if self.sampling == "uniform":
            ops.num_frames_student(num_frames, self.every_n, MAX_FRAMES, self.nf_student)
            return self.frame_idx
        K = self.student_frames
        if u is None:
            ops.random_uniform(self.u, self.sampling_seed, self._draws)
            self._draws += (self.u.numel() + 3) // 4
            u = self.u
        elif tuple(u.shape) != tuple(self.u.shape) or u.dtype != torch.float32:
            raise ValueError(f"u must be float32 {tuple(self.u.shape)} for sampling='{self.sampling}'")
        if self.sampling == "random_frames":
            ops.random_frame_index(u, num_frames, out=self.frame_idx_rand)
        else:
            ops.random_sequence_index(u, num_frames, K, out=self.frame_idx_rand)
        ops.sampled_lengths(num_frames, K, self.nf_student)
        return self.frame_idx_rand

    @property
    def lr(self) -> float:
        """tf.train.exponential_decay(base_learning_rate, global_step * batch_size, learning_rate_decay_examples,
        learning_rate_decay, staircase=True) (train.py:222-236; the flag defaults -- decay 1 -- keep it constant).
        global_step counts train ops (two per joint iteration); batch_size is the global batch of
        the data-parallel job."""
        if self.lr_decay == 1.0:
            return self.base_lr
        examples = self.global_step * self.B * self._world()
        return self.base_lr * self.lr_decay ** float(int(examples / self.lr_decay_examples))

    def _allreduce(self, params, lo: int = 0, hi: Optional[int] = None):
        """Average (a slice of) the flat gradient buffer over the data-parallel ranks.  With NCCL the
        collective is asynchronous: it runs on NCCL's stream over NVLink while this rank keeps
        launching backward kernels, and is waited for in `_finish_allreduce` before the optimizer.
        (The gloo backend of the CPU tests has no AVG and uses SUM / world.)"""
        n = self._world()
        if n <= 1:
            return
        buf = params.flat_g[lo:hi] if (lo or hi is not None) else params.flat_g
        if dist.get_backend() == "nccl":
            self._pending.append((params, dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True)))
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            buf.div_(n)

    def _dp(self):
        """(rank, world, group) for `HLstmEngine.classifier_backward` in the sharded data-parallel mode: the
        classifier's weight gradients come from one contraction over the all-gathered batch (9 MB of bf16 operands
        on the wire per rank and model instead of a 386 MB f32 reduce-scatter); EVC_DP_GATHER=0 keeps the
        reduce-scatter for them too (A/B)."""
        if not self._sharded() or (self._dp_ablate & 1) or os.environ.get("EVC_DP_GATHER", "1") == "0":
            return None
        return dist.get_rank(), self._world(), None

    def _reduce_grads(self, params, names):
        """Average the gradients of `names` over the ranks: whole-buffer all-reduce slices in the replicated
        mode, per-matrix reduce-scatter (owner keeps the average of its row block) + bias all-reduce in the
        sharded mode.  Asynchronous on NCCL's stream."""
        n = self._world()
        if n <= 1 or (self._dp_ablate & 1):
            return
        if not (self.shard_optimizer and dist.get_backend() == "nccl"):
            lo = min(params.offsets[x] for x in names)
            hi = max(params.offsets[x] + params.g[x].numel() for x in names)
            self._allreduce(params, lo, hi)
            return
        rank = dist.get_rank()
        gathered = (params.gates_w, params.experts_w) if self._dp() is not None else ()
        for x in names:
            if x in gathered:        # already averaged: computed from the gathered batch (engine._classifier_wgrad_gathered)
                continue
            g = params.g[x]
            if g.dim() == 2:
                r0, r1 = params.row_block(x, rank, n)
                w = dist.reduce_scatter_tensor(g[r0:r1], g, op=dist.ReduceOp.AVG, async_op=True)
            else:
                w = dist.all_reduce(g, op=dist.ReduceOp.AVG, async_op=True)
            self._pending.append((params, w))

    def _wait_classifier_wgrad(self, params):
        for eng in (getattr(self, "t_eng", None), getattr(self, "s_eng", None)):
            if eng is not None and eng.p is params:
                eng.wait_classifier_wgrad()

    def _apply(self, params):
        """clip + Adam for one parameter set (after its gradient collectives have completed)."""
        self._finish_allreduce(params)
        self._wait_classifier_wgrad(params)
        n = self._world()
        if n > 1 and self.shard_optimizer and dist.get_backend() == "nccl":
            self._gathers.setdefault(id(params), []).extend(
                params.apply_gradients_sharded(self.lr, self.clip, self.penalty, dist.get_rank(), n,
                                               gather=not (self._dp_ablate & 2), norm_group=self.norm_group))
        else:
            params.apply_gradients(self.lr, self.clip, self.penalty)

    # ---- single-GPU schedule: the classifier's variables (2/3 of the parameters) have final gradients right
    # after classifier_backward, so their clip+Adam pass (HBM-bound) runs on an optimizer stream next to the
    # LSTM backward (tensor-bound) instead of after it.  With several ranks the optimizer is sharded and
    # waits for the collectives instead.
    def _sharded(self) -> bool:
        return self._world() > 1 and self.shard_optimizer and dist.get_backend() == "nccl"

    def _early_apply_ok(self) -> bool:
        """One GPU: only with EVC_OVERLAP bit 8 (measured no gain: the optimizer's HBM traffic slows the power-capped
        GEMMs it runs next to).  Sharded data parallel: on by default (EVC_EARLY_OPT=0 disables) -- the classifier's
        tensors are 2/3 of the bytes; updating and all-gathering them during the LSTM backward takes their
        all-gather (and a rank's 1/N optimizer pass) off the end of the step."""
        if self.device.type != "cuda":
            return False
        if self._sharded():
            return os.environ.get("EVC_EARLY_OPT", "1") != "0"
        return self._world() == 1 and bool(overlap_mode() & OVERLAP_OPTIMIZER)

    def _opt_stream(self, params):
        st = self._opt_streams.get(id(params))
        if st is None:
            st = self._opt_streams[id(params)] = (torch.cuda.Stream(device=self.device), torch.cuda.Event())
        return st

    def _apply_classifier_early(self, params):
        """Call right after classifier_backward on the stream that ran it."""
        cur = torch.cuda.current_stream()
        opt, ev = self._opt_stream(params)
        params.begin_apply(self.lr)
        ev.record(cur)
        opt.wait_event(ev)
        with torch.cuda.stream(opt):
            self._apply_range(params, 8, None)

    def _apply_range(self, params, first, last):
        """clip + Adam of names[first:last] on the current stream once their gradient collectives are complete."""
        if last is None or last > 8:
            self._wait_classifier_wgrad(params)
        if self._sharded():
            self._finish_allreduce(params)      # (only this range's collectives have been issued so far)
            self._gathers.setdefault(id(params), []).extend(
                params.apply_gradients_sharded(self.lr, self.clip, self.penalty, dist.get_rank(), self._world(),
                                               gather=not (self._dp_ablate & 2), norm_group=self.norm_group,
                                               first=first, last=last, advance=False))
        else:
            self._finish_allreduce(params)
            params.apply_gradients(self.lr, self.clip, self.penalty, first=first, last=last, advance=False)

    def _apply_lstm_late(self, params):
        """The other half of `_apply_classifier_early`, after lstm_backward."""
        self._apply_range(params, 0, 8)
        torch.cuda.current_stream().wait_stream(self._opt_stream(params)[0])

    def _finish_gathers(self, params=None):
        """Make the current stream wait for the outstanding operand all-gathers of one model (or of all).  They are
        issued at the end of a step and only needed by the NEXT forward of that model, so the wait sits there: the
        teacher's forward starts while the student's rows are still on the wire."""
        keys = list(self._gathers) if params is None else [id(params)]
        for k in keys:
            for w in self._gathers.pop(k, []):
                w.wait()

    def _finish_allreduce(self, params=None):
        """Wait for the outstanding collectives (of one parameter set, or all)."""
        keep = []
        for p, w in self._pending:
            if params is None or p is params:
                w.wait()
            else:
                keep.append((p, w))
        self._pending = keep

    @staticmethod
    def _moe_offset(params) -> int:
        """Offset of the classifier tensors in the flat buffers (they follow the 8 LSTM tensors); their
        gradients (2/3 of the bytes) are final right after classifier_backward."""
        return params.offsets[params.gates_w]

    def _check(self, raw, num_frames, labels):
        if raw.dtype not in (torch.float32, torch.uint8) or raw.dim() != 3 or raw.shape[1] != MAX_FRAMES:
            raise ValueError("model_input_raw must be f32 (dequantised) or uint8 (quantised) [B, 300, feature_size]")
        if num_frames.dtype != torch.int32:
            raise TypeError("num_frames must be int32 (readers.py: tf.minimum(..., max_frames))")
        if labels is not None and tuple(labels.shape) != (raw.shape[0], self.cfg.vocab_size):
            raise ValueError("labels must be [B, vocab_size]")


class TeacherStudentTrainer(_Base):
    """Joint teacher+student training step (run_train.sh)."""

    def __init__(self, cfg: ModelConfig = ModelConfig(), batch_size: int = 256, device="cuda",
                 every_n: int = 10, num_inputs_to_lstm: int = 20, num_inputs_L1: int = 5,
                 base_learning_rate: float = 1e-3, clip_gradient_norm: float = 1.0,
                 regularization_penalty: float = 2.0, teacher_seed: Optional[int] = 0,
                 student_seed: Optional[int] = 1, lstm_gain: float = 1.0, shard_optimizer: Optional[bool] = None,
                 learning_rate_decay: float = 1.0, learning_rate_decay_examples: float = 4000000.0,
                 sampling: str = "uniform", sampling_seed: int = 0, precise: bool = False):
        super().__init__(cfg, batch_size, device, every_n, num_inputs_L1, base_learning_rate,
                         clip_gradient_norm, regularization_penalty, shard_optimizer, learning_rate_decay,
                         learning_rate_decay_examples, sampling, sampling_seed)
        # precise: split-bf16 operands (hi + lo planes, 3 tensor-core products per contraction): the arithmetic in
        # which the training curve follows the float32 reference graph to 1 % over hundreds of steps (include/evc.h)
        self.teacher = HLstmParams("model", cfg, device, teacher_seed, lstm_gain, precise)
        self.student = HLstmParams("model_student", cfg, device, student_seed, lstm_gain, precise)
        self.t_eng = HLstmEngine(self.teacher, batch_size, MAX_FRAMES, num_inputs_to_lstm, training=True)
        self.s_eng = HLstmEngine(self.student, batch_size, self.student_frames, num_inputs_L1, training=True)
        dev, B = self.device, batch_size
        self.rows = torch.zeros(4, B, dtype=torch.float32, device=dev)   # CE_T, CE_S, KL, REP rows
        self.losses = torch.zeros(8, dtype=torch.float32, device=dev)    # CE_T, CE_S, L_PRED, L_REP
        # The student only needs the teacher's final state and predictions (constants of its loss, F9): its
        # forward, backward and optimizer pass run on a second stream, so its latency-bound small-row
        # launches fill the SMs the teacher's kernels leave idle and the teacher's HBM-bound clip+Adam
        # runs next to the student's backward GEMMs.
        self.student_stream = (torch.cuda.Stream(device=self.device,
                                                 priority=int(os.environ.get("EVC_PRIO_STUDENT", "0")))
                               if (overlap_mode() & OVERLAP_STUDENT) and self.device.type == "cuda" else None)
        self._teacher_ready = torch.cuda.Event() if self.student_stream is not None else None

    def _forward_backward_two_streams(self, raw, num_frames, labels_u8, fuse_optimizer=False, u=None):
        B = self.B
        t, s = self.t_eng, self.s_eng
        main, side = torch.cuda.current_stream(), self.student_stream
        side.wait_stream(main)                       # the batch (and last step's weights) are ready
        # (the teacher's launches are issued first: when the host starts a step on an idle device, its large
        # kernels should be in the queue before the student's small ones)
        self._finish_gathers(self.teacher)
        t.forward(raw, None, True, num_frames, num_frames, mix=False)
        with torch.cuda.stream(side):
            self._finish_gathers(self.student)
            idx = self._student_sample(num_frames, u)
            s.forward(raw, idx, True, self.nf_student, num_frames, mix=False)
        t.classifier_loss_fused(labels_u8, None, 1.0 / B, 0.0, self.rows[0], None)
        self._teacher_ready.record(main)             # t.state and t.pred are final
        # (the teacher's classifier backward is issued before the student's: collectives of one communicator run in
        # issue order, and the teacher's operands are final first)
        t.classifier_backward(None, logits_done=True, dp=self._dp())
        self._reduce_grads(self.teacher, self.teacher.names[8:])
        if fuse_optimizer:
            self._apply_classifier_early(self.teacher)
        with torch.cuda.stream(side):
            side.wait_event(self._teacher_ready)
            ops.rep_loss(t.state, s.state, 4.0 / B, self.rows[3], s.dstate)
            s.classifier_loss_fused(labels_u8, t.pred, 1.0 / B, 1.0, self.rows[1], self.rows[2])
            s.classifier_backward(None, dstate_preset=True, logits_done=True, dp=self._dp())
            self._reduce_grads(self.student, self.student.names[8:])
            if fuse_optimizer:
                self._apply_classifier_early(self.student)
        t.lstm_backward()
        self._reduce_grads(self.teacher, self.teacher.names[:8])
        ops.reduce_rows(self.rows[0], 1.0 / B, self.losses[0:1])
        if fuse_optimizer:
            self._apply_lstm_late(self.teacher)
        with torch.cuda.stream(side):
            s.lstm_backward()
            self._reduce_grads(self.student, self.student.names[:8])
            if fuse_optimizer:
                self._apply_lstm_late(self.student)
            ops.reduce_rows(self.rows[1], 1.0 / B, self.losses[1:2])
            ops.reduce_rows(self.rows[2], 1.0, self.losses[2:3])
            ops.reduce_rows(self.rows[3], 1.0 / B, self.losses[3:4])

    def forward_backward(self, raw, num_frames, labels_u8, u=None):
        """Losses and gradients of both models (complete on the current stream when this returns)."""
        self._forward_backward(raw, num_frames, labels_u8, u=u)
        if self.student_stream is not None:
            torch.cuda.current_stream().wait_stream(self.student_stream)
        self.t_eng.wait_classifier_wgrad()
        self.s_eng.wait_classifier_wgrad()

    def _forward_backward(self, raw, num_frames, labels_u8, fuse_optimizer=False, u=None):
        if self.student_stream is not None:
            return self._forward_backward_two_streams(raw, num_frames, labels_u8, fuse_optimizer, u)
        B = self.B
        t, s = self.t_eng, self.s_eng
        # teacher: create_model on the normalised 300 frames (train.py:256,281-288)
        self._finish_gathers(self.teacher)
        t.forward(raw, None, True, num_frames, num_frames, mix=False)
        # student: every_n-th frame, float64 length rule (train.py:262-272,349-357) -- or a random sampler
        self._finish_gathers(self.student)
        idx = self._student_sample(num_frames, u)
        s.forward(raw, idx, True, self.nf_student, num_frames, mix=False)
        # teacher loss = penalty*reg + CE (train.py:297-324); reg enters through the optimizer's wd term.
        # One launch: mixture, CE rows and the gradients w.r.t. the logits.
        t.classifier_loss_fused(labels_u8, None, 1.0 / B, 0.0, self.rows[0], None)
        t.classifier_backward(None, logits_done=True, dp=self._dp())
        self._reduce_grads(self.teacher, self.teacher.names[8:])   # classifier gradients travel during the LSTM backward
        t.lstm_backward()
        self._reduce_grads(self.teacher, self.teacher.names[:8])
        # student loss = 2*L_REP + L_PRED + L_CE + penalty*reg (train.py:359-406); teacher tensors are
        # constants for the student's backward (F9)
        ops.rep_loss(t.state, s.state, 4.0 / B, self.rows[3], s.dstate)
        s.classifier_loss_fused(labels_u8, t.pred, 1.0 / B, 1.0, self.rows[1], self.rows[2])
        s.classifier_backward(None, dstate_preset=True, logits_done=True, dp=self._dp())
        self._reduce_grads(self.student, self.student.names[8:])
        s.lstm_backward()
        self._reduce_grads(self.student, self.student.names[:8])
        ops.reduce_rows(self.rows[0], 1.0 / B, self.losses[0:1])
        ops.reduce_rows(self.rows[1], 1.0 / B, self.losses[1:2])
        ops.reduce_rows(self.rows[2], 1.0, self.losses[2:3])
        ops.reduce_rows(self.rows[3], 1.0 / B, self.losses[3:4])

    def apply_gradients(self):
        # the teacher's optimizer pass runs while the student's last gradient slice is still on the wire
        # (one stream) or while the student's backward is still running (two streams)
        self._apply(self.teacher)
        if self.student_stream is not None:
            with torch.cuda.stream(self.student_stream):
                self._apply(self.student)
            torch.cuda.current_stream().wait_stream(self.student_stream)
        else:
            self._apply(self.student)

    def step(self, model_input_raw, num_frames, labels, u=None) -> None:
        """One iteration = both train ops (global_step += 2).  Asynchronous; read
        results with :meth:`fetch`.  u: the random samplers' U[0,1) draws (sampling != "uniform", parity runs)."""
        self._check(model_input_raw, num_frames, labels)
        if self.student_stream is not None and self._early_apply_ok():
            # optimizer passes are issued inside the schedule, each as soon as its gradients are final
            self._forward_backward(model_input_raw, num_frames, _as_u8(labels), fuse_optimizer=True, u=u)
            torch.cuda.current_stream().wait_stream(self.student_stream)
        else:
            self._forward_backward(model_input_raw, num_frames, _as_u8(labels), u=u)
            self.apply_gradients()
        self.global_step += 2

    def fetch(self) -> Dict[str, float]:
        """Device->host read of the step's scalars (the reference fetches them in sess.run and
        checks the loss with check_numerics)."""
        self._finish_gathers()      # (a host sync point: whoever reads the weights next sees complete operand copies)
        v = self.losses.tolist()
        wsq_t = self.teacher.wsq.tolist()
        wsq_s = self.student.wsq.tolist()
        reg_t = self.cfg.l2_penalty * 0.5 * (wsq_t[8] + wsq_t[9])
        reg_s = self.cfg.l2_penalty * 0.5 * (wsq_s[8] + wsq_s[9])
        out = {"teacher_ce": v[0], "teacher_reg": reg_t, "teacher_loss": self.penalty * reg_t + v[0],
               "l_ce": v[1], "l_pred": v[2], "l_rep": v[3], "student_reg": reg_s,
               "student_loss": 2 * v[3] + v[2] + v[1] + self.penalty * reg_s, "global_step": self.global_step}
        for k in ("teacher_loss", "student_loss"):
            if out[k] != out[k] or out[k] in (float("inf"), float("-inf")):
                raise FloatingPointError("LossTensor is inf or nan")   # slim create_train_op check_numerics
        return out


class StudentFinetuneTrainer(_Base):
    """Student-only fine-tuning step (run_finetune.sh; final_loss = penalty*reg + L_CE)."""

    def __init__(self, cfg: ModelConfig = ModelConfig(), batch_size: int = 256, device="cuda",
                 every_n: int = 10, num_inputs_L1: int = 5, base_learning_rate: float = 1e-3,
                 clip_gradient_norm: float = 1.0, regularization_penalty: float = 2.0,
                 student_seed: Optional[int] = 1, lstm_gain: float = 1.0, shard_optimizer: Optional[bool] = None,
                 learning_rate_decay: float = 1.0, learning_rate_decay_examples: float = 4000000.0,
                 sampling: str = "uniform", sampling_seed: int = 0, precise: bool = False):
        super().__init__(cfg, batch_size, device, every_n, num_inputs_L1, base_learning_rate,
                         clip_gradient_norm, regularization_penalty, shard_optimizer, learning_rate_decay,
                         learning_rate_decay_examples, sampling, sampling_seed)
        self.student = HLstmParams("model_student", cfg, device, student_seed, lstm_gain, precise)
        self.s_eng = HLstmEngine(self.student, batch_size, self.student_frames, num_inputs_L1, training=True)
        self.rows = torch.zeros(1, batch_size, dtype=torch.float32, device=self.device)
        self.losses = torch.zeros(4, dtype=torch.float32, device=self.device)

    def step(self, model_input_raw, num_frames, labels, u=None) -> None:
        self._check(model_input_raw, num_frames, labels)
        B, s = self.B, self.s_eng
        self._finish_gathers(self.student)
        idx = self._student_sample(num_frames, u)
        s.forward(model_input_raw, idx, True, self.nf_student, num_frames, mix=False)
        s.classifier_loss_fused(_as_u8(labels), None, 1.0 / B, 0.0, self.rows[0], None)
        s.classifier_backward(None, logits_done=True, dp=self._dp())
        self._reduce_grads(self.student, self.student.names[8:])
        early = self._early_apply_ok()
        if early:
            self._apply_classifier_early(self.student)
        s.lstm_backward()
        self._reduce_grads(self.student, self.student.names[:8])
        ops.reduce_rows(self.rows[0], 1.0 / B, self.losses[0:1])
        if early:
            self._apply_lstm_late(self.student)
        else:
            self._apply(self.student)
        self.global_step += 1

    def fetch(self) -> Dict[str, float]:
        self._finish_gathers()
        v = self.losses.tolist()
        wsq = self.student.wsq.tolist()
        reg = self.cfg.l2_penalty * 0.5 * (wsq[8] + wsq[9])
        out = {"l_ce": v[0], "student_reg": reg, "student_loss": self.penalty * reg + v[0],
               "global_step": self.global_step}
        if out["student_loss"] != out["student_loss"]:
            raise FloatingPointError("LossTensor is inf or nan")
        return out


class StudentEvaluator(_Base):
    """Student-only inference + top-k (eval_finetune.py:108-175, run_eval.sh)."""

    def __init__(self, params: HLstmParams, batch_size: int, every_n: int = 10, num_inputs_L1: int = 5,
                 top_k: int = 20, sampling: str = "uniform", sampling_seed: int = 0):
        super().__init__(params.cfg, batch_size, params.device, every_n, num_inputs_L1, 0.0, 0.0, 0.0,
                         sampling=sampling, sampling_seed=sampling_seed)
        self.student = params
        self.s_eng = HLstmEngine(params, batch_size, self.student_frames, num_inputs_L1, training=False)
        self.top_k = top_k
        self.rows = torch.zeros(batch_size, dtype=torch.float32, device=self.device)

    def step(self, model_input_raw, num_frames, labels=None, u=None):
        """Returns (predictions [B,V], top-k idx, top-k values, top-k labels|None); CE rows in self.rows."""
        self._check(model_input_raw, num_frames, labels)
        s = self.s_eng
        idx = self._student_sample(num_frames, u)
        s.forward(model_input_raw, idx, True, self.nf_student, num_frames)
        lab = _as_u8(labels) if labels is not None else None
        if lab is not None:
            ops.ce_kl_loss(s.pred, None, lab, 1.0, 0.0, self.rows, None, None)
        idx, val, tl = ops.topk(s.pred, self.top_k, lab)
        return s.pred, idx, val, tl


class TeacherEvaluator(_Base):
    """Teacher (all 300 frames) inference + top-k: the other half of validate.py:149-155 and the
    denominator of the student/teacher inference-cost ratio (BASELINE config #2)."""

    def __init__(self, params: HLstmParams, batch_size: int, num_inputs_to_lstm: int = 20, top_k: int = 20):
        super().__init__(params.cfg, batch_size, params.device, 10, 5, 0.0, 0.0, 0.0)
        self.teacher = params
        self.t_eng = HLstmEngine(params, batch_size, MAX_FRAMES, num_inputs_to_lstm, training=False)
        self.top_k = top_k
        self.rows = torch.zeros(batch_size, dtype=torch.float32, device=self.device)

    def step(self, model_input_raw, num_frames, labels=None):
        """Returns (predictions [B,V], top-k idx, top-k values, top-k labels|None); CE rows in self.rows."""
        self._check(model_input_raw, num_frames, labels)
        t = self.t_eng
        t.forward(model_input_raw, None, True, num_frames, num_frames)
        lab = _as_u8(labels) if labels is not None else None
        if lab is not None:
            ops.ce_kl_loss(t.pred, None, lab, 1.0, 0.0, self.rows, None, None)
        idx, val, tl = ops.topk(t.pred, self.top_k, lab)
        return t.pred, idx, val, tl


class TeacherStudentEvaluator(_Base):
    """Teacher + student inference of validate.py:109-189 (run_validate.sh): the teacher on all 300 frames under
    scope "model", the student on the sampled frames under "model_student"; fetched per batch (:173-175,262-263)
    are the STUDENT's predictions, the student's label loss and the state-matching loss
    mean_b sum_j (teacher_state - student_state)^2.  The student runs on a second stream next to the teacher."""

    def __init__(self, teacher: HLstmParams, student: HLstmParams, batch_size: int, every_n: int = 10,
                 num_inputs_to_lstm: int = 20, num_inputs_L1: int = 5, top_k: int = 20, sampling: str = "uniform",
                 sampling_seed: int = 0):
        if teacher.cfg != student.cfg or teacher.device != student.device:
            raise ValueError("teacher and student must share the model configuration and the device")
        super().__init__(student.cfg, batch_size, student.device, every_n, num_inputs_L1, 0.0, 0.0, 0.0,
                         sampling=sampling, sampling_seed=sampling_seed)
        self.teacher, self.student = teacher, student
        self.t_eng = HLstmEngine(teacher, batch_size, MAX_FRAMES, num_inputs_to_lstm, training=False)
        self.s_eng = HLstmEngine(student, batch_size, self.student_frames, num_inputs_L1, training=False)
        self.top_k = top_k
        self.rows = torch.zeros(batch_size, dtype=torch.float32, device=self.device)             # student CE rows
        self.state_loss_rows = torch.zeros(batch_size, dtype=torch.float32, device=self.device)  # L_REP rows
        self.losses = torch.zeros(2, dtype=torch.float32, device=self.device)   # student_label_loss, student_state_loss
        self.student_stream = (torch.cuda.Stream(device=self.device)
                               if (overlap_mode() & OVERLAP_STUDENT) and self.device.type == "cuda" else None)

    def step(self, model_input_raw, num_frames, labels=None, u=None):
        """Returns (student predictions [B,V], top-k idx, top-k values, top-k labels|None).  self.rows holds the
        student's cross-entropy per video, self.state_loss_rows the squared state distance per video and
        self.losses their batch means (validate.py's student_label_loss / student_state_loss)."""
        self._check(model_input_raw, num_frames, labels)
        t, s, B = self.t_eng, self.s_eng, self.B
        main, side = torch.cuda.current_stream(), self.student_stream
        if side is not None:
            side.wait_stream(main)
            t.forward_lstm(model_input_raw, None, True, num_frames, num_frames)      # the teacher's classifier is not fetched
            with torch.cuda.stream(side):
                idx = self._student_sample(num_frames, u)
                s.forward(model_input_raw, idx, True, self.nf_student, num_frames)
            main.wait_stream(side)
        else:
            t.forward_lstm(model_input_raw, None, True, num_frames, num_frames)
            idx = self._student_sample(num_frames, u)
            s.forward(model_input_raw, idx, True, self.nf_student, num_frames)
        ops.rep_loss(t.state, s.state, 0.0, self.state_loss_rows, None)
        ops.reduce_rows(self.state_loss_rows, 1.0 / B, self.losses[1:2])
        lab = _as_u8(labels) if labels is not None else None
        if lab is not None:
            ops.ce_kl_loss(s.pred, None, lab, 1.0, 0.0, self.rows, None, None)
            ops.reduce_rows(self.rows, 1.0 / B, self.losses[0:1])
        idx_k, val, tl = ops.topk(s.pred, self.top_k, lab)
        return s.pred, idx_k, val, tl


def evaluation_loop(evaluator, batches, metrics, log=None) -> Dict[str, object]:
    """The `while not coord.should_stop()` loop of eval_finetune.py:240-275 / validate.py:255-290: run every
    batch of `batches` (tuples of `readers.*.batches`) through `evaluator.step`, fold predictions, labels and the
    per-video cross-entropy into `metrics` (eval_util.EvaluationMetrics, optionally distributed) and return the
    epoch dictionary of `metrics.get()` plus `examples_processed` and the mean `examples_per_second`.

    The execution plan has a fixed batch size: the epoch's last, smaller batch is padded with empty videos
    (num_frames = 0, no labels) whose rows are dropped before they reach the metrics."""
    B, dev = evaluator.B, evaluator.device
    metrics.clear()
    examples, rates, state_losses = 0, [], []
    for ids, x, y, nf in batches:
        t0 = time.time()
        n = int(x.shape[0])
        if n > B:
            raise ValueError(f"batch of {n} videos exceeds the evaluator's plan ({B})")
        if n < B:
            x = torch.cat([x, x.new_zeros((B - n,) + tuple(x.shape[1:]))])
            y = torch.cat([y, y.new_zeros((B - n,) + tuple(y.shape[1:]))])
            nf = torch.cat([nf, nf.new_zeros(B - n)])
        xd, yd, nd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True), nf.to(dev, non_blocking=True)
        pred = evaluator.step(xd, nd, yd)[0]
        info = metrics.accumulate(pred[:n], yd[:n], evaluator.rows[:n])
        state_rows = getattr(evaluator, "state_loss_rows", None)
        if state_rows is not None:       # validate.py:262-269: the state-matching loss is logged per batch
            info = dict(info, student_loss=float(state_rows[:n].mean().item()) if n else 0.0)
            state_losses.append((info["student_loss"], n))
        dt = max(time.time() - t0, 1e-9)
        examples += n
        rates.append(n / dt)
        if log is not None:
            info = dict(info, examples_per_second=n / dt)
            log("examples_processed: %d | %s" % (examples, " | ".join(f"{k}: {v:.4g}" for k, v in info.items())))
    out = dict(metrics.get())
    out["examples_processed"] = examples
    out["examples_per_second"] = float(sum(rates) / len(rates)) if rates else 0.0
    if state_losses:
        out["avg_student_state_loss"] = sum(v * n for v, n in state_losses) / max(sum(n for _, n in state_losses), 1)
    return out
