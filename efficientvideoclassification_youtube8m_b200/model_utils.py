"""Frame samplers of code_student_uniform/model_utils.py:11-58 (config #5 'random' sampling).
tf.random_uniform is replaced by torch's device Philox generator; pass ``u`` explicitly to
reproduce a given draw (the index rule itself is bit-exact, tests/test_gpu_sampling.py)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .nn_ops import gather_frames


def _nf(num_frames):
    return num_frames.reshape(-1).to(torch.int32).contiguous()


def SampleRandomSequence(model_input, num_frames, num_samples, u: Optional[torch.Tensor] = None):
    """model_utils.py:11-36: a random contiguous run of num_samples frames."""
    B = model_input.shape[0]
    if u is None:
        u = torch.rand(B, dtype=torch.float32, device=model_input.device)
    idx = ops.random_sequence_index(u.reshape(B).contiguous(), _nf(num_frames), num_samples)
    return gather_frames(model_input, idx)


def SampleRandomFrames(model_input, num_frames, num_samples, u: Optional[torch.Tensor] = None):
    """model_utils.py:39-58: num_samples independent random frames per video."""
    B = model_input.shape[0]
    if u is None:
        u = torch.rand(B, num_samples, dtype=torch.float32, device=model_input.device)
    idx = ops.random_frame_index(u.contiguous(), _nf(num_frames))
    return gather_frames(model_input, idx)
