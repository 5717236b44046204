"""MoeModel behind the reference's plugin surface (code_student_uniform/video_level_models.py:394-448)."""
from __future__ import annotations

import torch

from . import models, ops, scope
from .flags import FLAGS
from .params import ModelConfig


class _MoeFn(torch.autograd.Function):
    """gates/experts GEMMs (tcgen05) + mixture; backward = mixture bwd + dgrad + wgrad GEMMs."""

    @staticmethod
    def forward(ctx, token, engine, state):
        S = engine.cfg.state_size
        ops.cast_bf16(state.contiguous(), engine.state_bf16, state.shape[0], S, S)
        engine.classifier_forward()
        ctx.engine = engine
        return engine.pred.clone()

    @staticmethod
    def backward(ctx, d_pred):
        e = ctx.engine
        e.classifier_backward(d_pred.contiguous(), dstate_preset=False)
        return torch.zeros_like(e.p.token), None, e.dstate.clone()


class MoeModel(models.BaseModel):
    """A softmax over a mixture of logistic models (with L2 regularization)."""

    def create_model(self, model_input, vocab_size, num_mixtures=None, l2_penalty=1e-8, **unused_params):
        """model_input f32 [B, 4*lstm_cells] on the device -> {"predictions": [B, vocab_size]}.
        Variables `<scope>/classifier/{gates/weights, experts/weights, experts/biases}` belong to
        the enclosing model scope (frame_level_models.py:259-265)."""
        num_mixtures = num_mixtures or FLAGS.moe_num_mixtures
        root = scope.root_scope()
        B, S = model_input.shape
        engine = scope.current_engine()
        if engine is None:  # MoeModel used on its own: variables are created under the current scope
            cfg = ModelConfig(lstm_cells=S // 4, vocab_size=vocab_size, num_mixtures=num_mixtures,
                              l2_penalty=l2_penalty)
            p = scope.get_params(root, cfg, model_input.device)
            engine = scope.get_engine(p, B, 1, 1, torch.is_grad_enabled())
        cfg = engine.cfg
        if (cfg.vocab_size, cfg.num_mixtures, cfg.state_size) != (vocab_size, num_mixtures, S):
            raise ValueError("MoeModel arguments do not match the variables of this scope")
        if abs(cfg.l2_penalty - l2_penalty) > 0:
            raise ValueError("l2_penalty differs from the scope's regulariser")
        pred = _MoeFn.apply(engine.p.token, engine, model_input)
        return {"predictions": pred}
