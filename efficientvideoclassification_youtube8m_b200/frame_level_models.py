"""HierarchicalLstmModel behind the reference's plugin surface
(code_student_uniform/frame_level_models.py:198-338)."""
from __future__ import annotations

import torch

from . import models, scope, video_level_models
from .flags import FLAGS
from .params import ModelConfig


class _HLstmStateFn(torch.autograd.Function):
    """Chunked two-level BasicLSTM stack: model_input -> final state [c0|h0|c1|h1] of RNN_L2."""

    @staticmethod
    def forward(ctx, token, engine, model_input, num_frames):
        engine.forward_lstm(model_input.contiguous(), None, False, num_frames.contiguous())
        ctx.engine = engine
        return engine.state.clone()

    @staticmethod
    def backward(ctx, d_state):
        e = ctx.engine
        e.dstate.copy_(d_state)
        e.lstm_backward()
        return torch.zeros_like(e.p.token), None, None, None


def _cfg(model_input, vocab_size):
    return ModelConfig(feature_size=model_input.shape[2], lstm_cells=FLAGS.lstm_cells,
                       lstm_layers=FLAGS.lstm_layers, vocab_size=vocab_size,
                       num_mixtures=FLAGS.moe_num_mixtures)


class HierarchicalLstmModel(models.BaseModel):

    def _run(self, model_input, vocab_size, num_frames, num_chunks, unused_params):
        if not (torch.is_tensor(model_input) and model_input.is_cuda and model_input.dtype == torch.float32):
            raise TypeError("model_input must be a float32 CUDA tensor [batch, frames, features]")
        root = scope.root_scope()
        B, T, _ = model_input.shape
        p = scope.get_params(root, _cfg(model_input, vocab_size), model_input.device)
        engine = scope.get_engine(p, B, T, num_chunks, torch.is_grad_enabled())
        with scope.variable_scope("RNN"):
            state = _HLstmStateFn.apply(p.token, engine, model_input, num_frames)
        with scope.variable_scope("classifier"), scope.use_engine(engine):
            aggregated_model = getattr(video_level_models, FLAGS.video_level_classifier_model)
            final_state_predictions = aggregated_model().create_model(
                model_input=state, vocab_size=vocab_size, **unused_params)
        return state, final_state_predictions

    def create_model(self, model_input, vocab_size, num_frames, **unused_params):
        """Teacher: FLAGS.num_inputs_to_lstm chunks of max_num_frames/num_inputs_to_lstm frames
        (frame_level_models.py:200-267).  Returns (state, {"predictions": ...})."""
        if num_frames.dtype != torch.int32:
            raise TypeError("num_frames must be int32 (frame_level_models.py:240 mixes it with python ints)")
        return self._run(model_input, vocab_size, num_frames, FLAGS.num_inputs_to_lstm, unused_params)

    def create_model_inference(self, model_input, vocab_size, every_n, num_inputs_L1, num_frames, **unused_params):
        """Student: num_inputs_L1 chunks of (max_num_frames/every_n)/num_inputs_L1 frames, int64
        lengths (frame_level_models.py:269-338)."""
        if num_frames.dtype != torch.int64:
            raise TypeError("num_frames must be int64 (frame_level_models.py:308-309 casts to tf.int64)")
        expected = int(FLAGS.max_num_frames / every_n)
        if model_input.shape[1] != expected:
            raise ValueError(f"student input has {model_input.shape[1]} frames, expected max_num_frames/every_n "
                             f"= {expected}")
        return self._run(model_input, vocab_size, num_frames, num_inputs_L1, unused_params)
