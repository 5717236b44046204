"""Graph-level ops the reference's build_graph applies around the models (train.py:253-272),
as differentiable-free device functions (the inputs are data)."""
from __future__ import annotations

import torch

from . import ops


def l2_normalize(model_input_raw: torch.Tensor) -> torch.Tensor:
    """tf.nn.l2_normalize(model_input_raw, feature_dim) (train.py:253-256)."""
    B, T, D = model_input_raw.shape
    out = torch.empty_like(model_input_raw)
    ops.frames_pack(model_input_raw, None, T, 1, True, out_f32=out)
    return out


def gather_frames(model_input: torch.Tensor, frame_index: torch.Tensor) -> torch.Tensor:
    """transpose / tf.gather(list_index_to_retain) / transpose (train.py:270-272) when
    frame_index is int32 [K]; tf.gather_nd with per-video indices (model_utils.py:34-36,55-58)
    when it is int32 [B,K]."""
    B, T, D = model_input.shape
    K = frame_index.shape[-1]
    out = torch.empty(B, K, D, dtype=torch.float32, device=model_input.device)
    ops.frames_pack(model_input, frame_index.contiguous(), K, 1, False, out_f32=out)
    return out


def sample_every_n(model_input: torch.Tensor, every_n: int) -> torch.Tensor:
    """The student's uniform sampler: frames [0, n, 2n, ... <= 299] (train.py:262-272)."""
    from .steps import uniform_frame_indices
    idx = torch.tensor(uniform_frame_indices(every_n), dtype=torch.int32, device=model_input.device)
    return gather_frames(model_input, idx)


def num_frames_student(num_frames: torch.Tensor, every_n: int, max_frames: int = 300) -> torch.Tensor:
    """int64((num_frames / 300) * int(300/every_n)) in float64 (train.py:263-264)."""
    return ops.num_frames_student(num_frames, every_n, max_frames)
