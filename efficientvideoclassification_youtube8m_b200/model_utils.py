"""Frame samplers of code_student_uniform/model_utils.py:11-58 (config #5 'random' sampling).
tf.random_uniform is replaced by the library's counter-based Philox4x32-10 kernel (evc_random_uniform; seed from
``set_random_seed``, the counter advances with every draw); pass ``u`` explicitly to reproduce a given draw (the
index rule itself is bit-exact, tests/test_gpu_eval.py::test_gather_and_random_samplers_bit_exact)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .nn_ops import gather_frames


_rng = {"seed": 0, "offset": 0}


def set_random_seed(seed: int) -> None:
    """tf.set_random_seed for the samplers' uniform draws."""
    _rng["seed"], _rng["offset"] = int(seed), 0


def random_uniform(shape, device) -> torch.Tensor:
    """tf.random_uniform(shape) on the device; consecutive calls consume consecutive Philox counters."""
    u = torch.empty(shape, dtype=torch.float32, device=device)
    ops.random_uniform(u, _rng["seed"], _rng["offset"])
    _rng["offset"] += (u.numel() + 3) // 4
    return u


def _nf(num_frames):
    return num_frames.reshape(-1).to(torch.int32).contiguous()


def SampleRandomSequence(model_input, num_frames, num_samples, u: Optional[torch.Tensor] = None):
    """model_utils.py:11-36: a random contiguous run of num_samples frames."""
    B = model_input.shape[0]
    if u is None:
        u = random_uniform((B,), model_input.device)
    idx = ops.random_sequence_index(u.reshape(B).contiguous(), _nf(num_frames), num_samples)
    return gather_frames(model_input, idx)


def SampleRandomFrames(model_input, num_frames, num_samples, u: Optional[torch.Tensor] = None):
    """model_utils.py:39-58: num_samples independent random frames per video."""
    B = model_input.shape[0]
    if u is None:
        u = random_uniform((B, num_samples), model_input.device)
    idx = ops.random_frame_index(u.contiguous(), _nf(num_frames))
    return gather_frames(model_input, idx)
