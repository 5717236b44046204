"""slim.learning.create_train_op + tf.train.AdamOptimizer for the plugin path
(train.py:222-242,325-334,408-418): total_loss.backward() through the autograd wrappers of the
CUDA kernels, NCCL average of the flat gradient buffer over data-parallel ranks, then the fused
per-variable clip_by_norm + TF-Adam kernel."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .params import HLstmParams

_global_step = [0]


def get_global_step() -> int:
    return _global_step[0]


def reset_global_step() -> None:
    _global_step[0] = 0


class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate) with TF defaults."""

    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon


def exponential_decay(base_learning_rate, global_examples, decay_examples, decay_rate, staircase=True):
    """tf.train.exponential_decay (train.py:223-236); decay_rate 1 (the default) keeps lr constant."""
    e = global_examples / decay_examples
    if staircase:
        e = float(int(e))
    return base_learning_rate * (decay_rate ** e)


class TrainOp:
    def __init__(self, total_loss, optimizer, variables_to_train: HLstmParams, clip_gradient_norm):
        self.loss, self.opt, self.params, self.clip = total_loss, optimizer, variables_to_train, clip_gradient_norm

    def run(self, retain_graph: bool = False) -> torch.Tensor:
        """Gradients w.r.t. this scope's variables only (F9), clip, Adam, global_step += 1 (F10)."""
        p = self.params
        p.reg_grad_scale = None
        self.loss.backward(inputs=[p.token], retain_graph=retain_graph)
        p.token.grad = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(p.flat_g, op=dist.ReduceOp.AVG)
        penalty = float(p.reg_grad_scale) if p.reg_grad_scale is not None else 0.0
        p.apply_gradients(self.opt.learning_rate, float(self.clip), penalty, self.opt.beta1, self.opt.beta2,
                          self.opt.epsilon)
        _global_step[0] += 1
        loss = self.loss.detach()
        if not bool(torch.isfinite(loss)):
            raise FloatingPointError("LossTensor is inf or nan")      # slim: check_numerics on the loss
        return loss


def create_train_op(total_loss, optimizer, global_step=None, variables_to_train=None, clip_gradient_norm=0):
    if not isinstance(variables_to_train, HLstmParams):
        raise TypeError("variables_to_train must be scope.trainable_variables('<scope>')")
    return TrainOp(total_loss, optimizer, variables_to_train, clip_gradient_norm)
