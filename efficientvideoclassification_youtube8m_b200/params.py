"""The 11-tensor weight layout of one H-LSTM model (README.md:98,105; validate.py:350-374;
train_convert_model.py:501-511) held in HBM:

  * one flat f32 buffer each for the master weights, gradients and the two Adam moments
    (so the data-parallel allreduce and the optimizer stream over contiguous memory);
  * per-tensor views keyed by the exact TF variable names (``state_dict`` round-trips a
    reference checkpoint's name->array map);
  * bf16 operand copies of the five matrices in the SAME [rows, cols] orientation TF uses
    (row pitch padded to 64 elements for TMA), refreshed by the fused clip+Adam kernel.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops


@dataclass(frozen=True)
class ModelConfig:
    feature_size: int = 1152      # rgb 1024 + audio 128 (run_train.sh --feature_sizes "1024, 128")
    lstm_cells: int = 1024        # FLAGS.lstm_cells
    lstm_layers: int = 2          # run_train.sh --lstm_layers 2
    vocab_size: int = 4716        # reader.num_classes
    num_mixtures: int = 2         # FLAGS.moe_num_mixtures
    l2_penalty: float = 1e-8      # MoeModel.create_model default (video_level_models.py:401)

    @property
    def state_size(self) -> int:
        return 2 * self.lstm_layers * self.lstm_cells


def variable_names(scope: str) -> List[str]:
    names = []
    for level in ("RNN_L1", "RNN_L2"):
        for cell in (0, 1):
            base = f"{scope}/{level}/rnn/multi_rnn_cell/cell_{cell}/basic_lstm_cell"
            names += [base + "/kernel", base + "/bias"]
    names += [f"{scope}/classifier/gates/weights", f"{scope}/classifier/experts/weights",
              f"{scope}/classifier/experts/biases"]
    return names


def variable_shapes(scope: str, cfg: ModelConfig) -> Dict[str, tuple]:
    if cfg.lstm_layers != 2:
        raise NotImplementedError("the hot path is the 2-layer stack of run_train.sh (--lstm_layers 2)")
    H, S, D, V, M = cfg.lstm_cells, cfg.state_size, cfg.feature_size, cfg.vocab_size, cfg.num_mixtures
    n = variable_names(scope)
    return {n[0]: (D + H, 4 * H), n[1]: (4 * H,), n[2]: (2 * H, 4 * H), n[3]: (4 * H,),
            n[4]: (S + H, 4 * H), n[5]: (4 * H,), n[6]: (2 * H, 4 * H), n[7]: (4 * H,),
            n[8]: (S, V * (M + 1)), n[9]: (S, V * M), n[10]: (V * M,)}


class HLstmParams:
    """Weights, gradients, Adam state and bf16 operand copies of one variable scope."""

    def __init__(self, scope: str, cfg: ModelConfig, device, seed: Optional[int] = 0,
                 lstm_gain: float = 1.0, precise: bool = False):
        self.scope, self.cfg, self.device = scope, cfg, torch.device(device)
        # split-bf16 "precise" mode (include/evc.h): every operand copy gets a residual plane, contractions run
        # as hi*hi + hi*lo + lo*hi (the mode in which the 200-step loss criterion holds at lr 1e-3)
        self.precise = bool(precise)
        self.names = variable_names(scope)
        self.shapes = variable_shapes(scope, cfg)
        self.offsets, off = {}, 0
        for n in self.names:
            self.offsets[n] = off
            numel = int(np.prod(self.shapes[n]))
            assert numel % 4 == 0, "tensor sizes must keep 16-byte alignment inside the flat buffer"
            off += numel
        self.numel = off
        self.flat_w = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.flat_v = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.w, self.g, self.m, self.v = {}, {}, {}, {}
        for n in self.names:
            o, shp = self.offsets[n], self.shapes[n]
            k = int(np.prod(shp))
            self.w[n] = self.flat_w[o:o + k].view(shp)
            self.g[n] = self.flat_g[o:o + k].view(shp)
            self.m[n] = self.flat_m[o:o + k].view(shp)
            self.v[n] = self.flat_v[o:o + k].view(shp)
        # bf16 operand copies (matrices only), row pitch padded to a multiple of 64
        self.shadow, self.ld, self.shadow_lo = {}, {}, {}
        for n in self.names:
            shp = self.shapes[n]
            if len(shp) == 2:
                ld = ops.pad8(shp[1], 64)
                self.ld[n] = ld
                self.shadow[n] = torch.zeros(shp[0], ld, dtype=torch.bfloat16, device=self.device)
                if self.precise:
                    self.shadow_lo[n] = torch.zeros(shp[0], ld, dtype=torch.bfloat16, device=self.device)
        self.normsq = torch.zeros(len(self.names), dtype=torch.float32, device=self.device)
        self.wsq = torch.zeros(len(self.names), dtype=torch.float32, device=self.device)
        # Norms taken where the gradients are produced (one process, CUDA): the weight-gradient GEMMs leave sum g^2 of
        # their matrix in norm_aux[i, 0] (evc_gemm_bf16_wgrad), evc_reg_cross leaves <g, w> of the two regularised
        # matrices in norm_aux[i, 1], and the clip+Adam kernel accumulates sum w^2 of the weights it writes into
        # wsq_next -- so the optimizer needs no sumsq pass over the matrices (`begin_fused_norms`, `apply_gradients`).
        self.norm_aux = torch.zeros(len(self.names), 2, dtype=torch.float32, device=self.device)
        self.wsq_next = torch.zeros(len(self.names), dtype=torch.float32, device=self.device)
        self._fused_ready = set()     # variables whose norm_aux row is fresh from the latest backward
        self._wsq_valid = False       # wsq_next holds sum w^2 of the CURRENT weights of the regularised matrices
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.lr_t = torch.zeros(1, dtype=torch.float32, device=self.device)
        # autograd handle: the kernels write weight gradients straight into flat_g, the token only makes
        # torch call the backward of the plugin functions
        self.token = torch.zeros(1, dtype=torch.float32, device=self.device, requires_grad=True)
        self.reg_grad_scale = None   # d(total_loss)/d(reg_loss) recorded by losses.regularization_loss
        if seed is not None:
            self.init_glorot(seed, lstm_gain)

    # ---- names of the individual tensors
    def kernel(self, level: int, cell: int) -> str:
        return self.names[(level * 2 + cell) * 2]

    def bias(self, level: int, cell: int) -> str:
        return self.names[(level * 2 + cell) * 2 + 1]

    @property
    def gates_w(self) -> str:
        return self.names[8]

    @property
    def experts_w(self) -> str:
        return self.names[9]

    @property
    def experts_b(self) -> str:
        return self.names[10]

    # ---- initialisation / checkpoint interchange
    def init_glorot(self, seed: int, lstm_gain: float = 1.0) -> None:
        """TF defaults: glorot_uniform kernels / xavier fully_connected weights, zero biases
        (SURVEY Appendix A.6).  Same RandomState stream as oracle.init_params."""
        rng = np.random.RandomState(seed)
        sd = {}
        for n in self.names:
            shp = self.shapes[n]
            if len(shp) == 1:
                sd[n] = np.zeros(shp, dtype=np.float32)
            else:
                limit = math.sqrt(6.0 / (shp[0] + shp[1]))
                w = rng.uniform(-limit, limit, size=shp)
                if "basic_lstm_cell" in n:
                    w = w * lstm_gain
                sd[n] = w.astype(np.float32)
        self.load_state_dict(sd)

    def load_state_dict(self, sd: Dict[str, "np.ndarray | torch.Tensor"], strict: bool = True) -> None:
        missing = [n for n in self.names if n not in sd]
        if missing and strict:
            raise KeyError(f"missing variables: {missing}")
        for n in self.names:
            if n not in sd:
                continue
            t = torch.as_tensor(np.asarray(sd[n]) if not torch.is_tensor(sd[n]) else sd[n])
            if tuple(t.shape) != tuple(self.shapes[n]):
                raise ValueError(f"{n}: shape {tuple(t.shape)} != {self.shapes[n]}")
            self.w[n].copy_(t.to(torch.float32))
        self.refresh_shadows()

    def _sync_if_stale(self) -> None:
        """With the optimizer sharded over the data-parallel ranks (the default for world > 1) a rank only updates
        its own row block of every f32 master matrix; the other rows are refreshed here.  COLLECTIVE: every rank
        must reach state_dict()/save() together (as in a checkpoint hook that runs on all ranks; let rank 0
        alone write the file afterwards)."""
        if not getattr(self, "_master_stale", False):
            return
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("master weights are sharded over ranks that are gone: call sync_master_weights() "
                               "before destroying the process group")
        self.sync_master_weights(dist.get_rank(), dist.get_world_size())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """name -> f32 tensor (clone).  Collective after sharded updates, see `_sync_if_stale`."""
        self._sync_if_stale()
        return {n: self.w[n].detach().clone() for n in self.names}

    def save(self, path: str, write: bool = True) -> None:
        """name -> array map of this scope's variables (`.npz`; keys are the TF checkpoint names).  Collective
        after sharded updates (see `_sync_if_stale`); pass write=(rank == 0) to let one rank write the file."""
        self._sync_if_stale()
        if write:
            np.savez(path, **{n: self.w[n].detach().cpu().numpy() for n in self.names})

    def load(self, path: str, from_scope: Optional[str] = None) -> None:
        """Load a name -> array map.  `from_scope` renames `<from_scope>/...` keys to this scope: with
        from_scope="model_student" into scope "model_student" it is the restore of
        train_convert_model.py:496-517; any other pair re-homes a checkpoint (e.g. student init from
        the teacher)."""
        with np.load(path) as z:
            sd = {}
            for k in z.files:
                name = k
                if from_scope is not None and k.startswith(from_scope + "/"):
                    name = self.scope + k[len(from_scope):]
                sd[name] = z[k]
        self.load_state_dict({n: sd[n] for n in self.names if n in sd}, strict=True)

    # ---- gradient norms from the producers (see __init__)
    def fused_norms(self) -> bool:
        """Whether the engine should take the matrices' gradient norms in the weight-gradient GEMM epilogues: one
        process only (with data parallelism the norm is that of the AVERAGED gradient), EVC_FUSED_NORMS=0 disables."""
        if self.device.type != "cuda" or os.environ.get("EVC_FUSED_NORMS", "1") == "0":
            return False
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)

    def begin_fused_norms(self, first: int, last: int) -> None:
        """Before the weight-gradient GEMMs of names[first:last]: zero their partial-norm slots."""
        ops.fill_f32(self.norm_aux[first:last], 0.0)
        self._fused_ready -= set(range(first, last))

    def end_fused_norms(self, indices) -> None:
        """After them: the slots of `indices` are complete once the launches issued so far have run."""
        self._fused_ready |= set(indices)

    def refresh_shadows(self) -> None:
        self._wsq_valid = False           # (called whenever the masters were changed from outside)
        for n, s in self.shadow.items():
            rows, cols = self.shapes[n]
            ops.cast_bf16(self.w[n], s, rows, cols, self.ld[n], self.shadow_lo.get(n))

    # ---- data-parallel optimizer sharding (ZeRO-1 style): every matrix is split into `world` row blocks; a
    # rank receives the averaged gradient of its block (reduce-scatter), updates only those rows of the f32
    # master / Adam moments, and the refreshed bf16 operand rows are all-gathered.  Biases stay replicated.
    def matrix_names(self) -> List[str]:
        return [n for n in self.names if len(self.shapes[n]) == 2]

    def vector_names(self) -> List[str]:
        return [n for n in self.names if len(self.shapes[n]) == 1]

    def row_block(self, n: str, rank: int, world: int):
        rows = self.shapes[n][0]
        if rows % world != 0:
            raise ValueError(f"{n}: {rows} rows do not split over {world} ranks")
        per = rows // world
        return rank * per, (rank + 1) * per

    def apply_gradients_sharded(self, lr: float, clip_gradient_norm: float, regularization_penalty: float,
                                rank: int, world: int, group=None, beta1: float = 0.9, beta2: float = 0.999,
                                eps: float = 1e-8, gather: bool = True, norm_group=None, first: int = 0,
                                last: Optional[int] = None, advance: bool = True):
        """Same update as apply_gradients for the rows this rank owns.  Expects the matrix gradients
        reduce-scattered (owned rows = average over ranks) and the bias gradients all-reduced.  Returns the
        async all-gather handles of the bf16 operand copies (wait before the next forward).
        Variables names[first:last] only (the clip is per variable, so the classifier's tensors can be updated and
        gathered while the LSTM backward is still running); advance=False after `begin_apply`."""
        import torch.distributed as dist
        wd = float(regularization_penalty) * self.cfg.l2_penalty
        last = len(self.names) if last is None else last
        in_range = set(self.names[first:last])
        ops.fill_f32(self.normsq[first:last], 0.0)
        ops.fill_f32(self.wsq[first:last], 0.0)
        if advance:
            ops.adam_lr(self.adam_step, lr, beta1, beta2, self.lr_t)
        reg = (self.gates_w, self.experts_w)
        idx = {n: i for i, n in enumerate(self.names)}
        blocks = {}
        for n in self.matrix_names():
            if n not in in_range:
                continue
            r0, r1 = self.row_block(n, rank, world)
            blocks[n] = (r0, r1)
            i = idx[n]
            w = self.w[n][r0:r1] if n in reg else None
            ops.sumsq(self.g[n][r0:r1], w, wd, self.normsq[i:i + 1], self.wsq[i:i + 1] if w is not None else None)
        # per-variable norms need the other ranks' row blocks: two tiny SUM all-reduces (11 floats each)
        ng = norm_group if norm_group is not None else group
        dist.all_reduce(self.normsq[first:last], op=dist.ReduceOp.SUM, group=ng)
        dist.all_reduce(self.wsq[first:last], op=dist.ReduceOp.SUM, group=ng)
        for n in self.vector_names():
            if n not in in_range:
                continue
            i = idx[n]
            ops.sumsq(self.g[n], None, 0.0, self.normsq[i:i + 1], None)
        handles = []
        for n in self.names[first:last]:
            i = idx[n]
            if n in blocks:
                r0, r1 = blocks[n]
                cols, ld = self.shapes[n][1], self.ld[n]
                ops.clip_adam(self.w[n][r0:r1], self.g[n][r0:r1], self.m[n][r0:r1], self.v[n][r0:r1],
                              self.normsq[i:i + 1], float(clip_gradient_norm), wd if n in reg else 0.0, self.lr_t,
                              beta1, beta2, eps, self.shadow[n][r0:r1], cols, ld,
                              self.shadow_lo[n][r0:r1] if self.precise else None)
                if not gather:          # attribution experiments only (EVC_DP_ABLATE=2)
                    continue
                handles.append(dist.all_gather_into_tensor(self.shadow[n], self.shadow[n][r0:r1], group=group,
                                                           async_op=True))
                if self.precise:
                    handles.append(dist.all_gather_into_tensor(self.shadow_lo[n], self.shadow_lo[n][r0:r1],
                                                               group=group, async_op=True))
            else:
                ops.clip_adam(self.w[n], self.g[n], self.m[n], self.v[n], self.normsq[i:i + 1],
                              float(clip_gradient_norm), 0.0, self.lr_t, beta1, beta2, eps, None, 0, 0)
        self._master_stale = world > 1
        self._wsq_valid = False
        self._fused_ready -= set(range(first, last))
        return handles

    def sync_master_weights(self, rank: int, world: int, group=None, include_slots: bool = False) -> None:
        """After sharded updates only the owned rows of the f32 masters (and of the Adam moments) are current:
        gather the rest (needed before state_dict()/save()/checkpoints, not on the training path)."""
        import torch.distributed as dist
        if world <= 1 or not getattr(self, "_master_stale", False):
            return
        for n in self.matrix_names():
            r0, r1 = self.row_block(n, rank, world)
            for t in ([self.w[n]] + ([self.m[n], self.v[n]] if include_slots else [])):
                dist.all_gather_into_tensor(t, t[r0:r1].clone(), group=group)
        if include_slots:
            self._master_stale = False
        else:
            self._slots_stale = True
            self._master_stale = False

    def sync_all(self, include_slots: bool = True) -> None:
        """Collective: make the weights (and optimizer slots) of every rank complete; no-op on one rank."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() <= 1:
            return
        if getattr(self, "_slots_stale", False) and include_slots:
            self._master_stale = True          # the weights were synced without the slots: gather both again
        self.sync_master_weights(dist.get_rank(), dist.get_world_size(), include_slots=include_slots)
        if include_slots:
            self._slots_stale = False

    # ---- slim.learning.create_train_op: per-variable clip_by_norm + Adam (train.py:329-334)
    def begin_apply(self, lr: float, beta1: float = 0.9, beta2: float = 0.999) -> None:
        """Advance the Adam step counter and compute lr_t once, for an update that is then applied in
        several `apply_gradients(..., first=.., last=.., advance=False)` passes over slices of the variables."""
        ops.adam_lr(self.adam_step, lr, beta1, beta2, self.lr_t)

    def apply_gradients(self, lr: float, clip_gradient_norm: float, regularization_penalty: float,
                        beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                        first: int = 0, last: Optional[int] = None, advance: bool = True) -> None:
        """g <- g + penalty*l2_penalty*w for the two MoE matrices (gradient of
        penalty * slim.l2_regularizer(l2_penalty)(w)), per-variable clip, TF Adam, bf16 refresh.
        Variables names[first:last] only (the clip is per variable, so a slice can be updated as soon as its
        gradients are final); advance=False after `begin_apply`."""
        wd = float(regularization_penalty) * self.cfg.l2_penalty
        last = len(self.names) if last is None else last
        ops.fill_f32(self.normsq[first:last], 0.0)
        ops.fill_f32(self.wsq[first:last], 0.0)
        if advance:
            ops.adam_lr(self.adam_step, lr, beta1, beta2, self.lr_t)
        reg = (self.gates_w, self.experts_w)
        fused = {i for i in range(first, last) if i in self._fused_ready and len(self.shapes[self.names[i]]) == 2}
        self._fused_ready -= set(range(first, last))
        reg_idx = [i for i in range(first, last) if self.names[i] in reg]
        reg_fused = bool(reg_idx) and all(i in fused for i in reg_idx)
        if reg_idx and not reg_fused:
            fused -= set(reg_idx)
            self._wsq_valid = False       # their clip+Adam launches below do not maintain wsq_next
        if reg_fused:
            # sum w^2 of the regularised matrices: left by the previous step's clip+Adam launches (or one pass over
            # w after the weights were set from outside); it is this step's regulariser value and norm term
            lo, hi = reg_idx[0], reg_idx[-1] + 1
            if not self._wsq_valid:
                ops.fill_f32(self.wsq_next[lo:hi], 0.0)
                for i in reg_idx:
                    ops.sumsq(self.w[self.names[i]], None, 0.0, self.wsq_next[i:i + 1], None)
            self.wsq[lo:hi].copy_(self.wsq_next[lo:hi])
            ops.fill_f32(self.wsq_next[lo:hi], 0.0)
            self._wsq_valid = True
        for i, n in enumerate(self.names):
            if not first <= i < last or i in fused:
                continue
            w = self.w[n] if n in reg else None
            ops.sumsq(self.g[n], w, wd, self.normsq[i:i + 1], self.wsq[i:i + 1] if w is not None else None)
        for i, n in enumerate(self.names):
            if not first <= i < last:
                continue
            shadow = self.shadow.get(n)
            cols = self.shapes[n][1] if shadow is not None else 0
            extra = {}
            if i in fused:
                extra["normsq_fused"] = self.norm_aux[i, 0:1]
                if n in reg:
                    extra.update(reg_cross=self.norm_aux[i, 1:2], reg_wsq=self.wsq[i:i + 1],
                                 wsq_out=self.wsq_next[i:i + 1])
            ops.clip_adam(self.w[n], self.g[n], self.m[n], self.v[n], self.normsq[i:i + 1],
                          float(clip_gradient_norm), wd if n in reg else 0.0, self.lr_t, beta1, beta2, eps,
                          shadow, cols, self.ld.get(n, 0), self.shadow_lo.get(n), **extra)
