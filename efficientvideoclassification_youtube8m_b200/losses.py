"""Losses of the teacher-student step behind the reference's loss plugin surface
(code_student_uniform/losses.py:8-25,86-97; train.py:305-307,359-362,398-406)."""
from __future__ import annotations

import torch

from . import ops, scope as _scope


def _labels_u8(labels):
    if labels.dtype == torch.bool:
        return labels.contiguous().view(torch.uint8)
    if labels.dtype == torch.uint8:
        return labels.contiguous()
    return (labels != 0).contiguous().view(torch.uint8)     # tf.cast(labels, tf.float32) of 0/1 labels


class BaseLoss(object):
    """Inherit from this class when implementing new losses."""

    def calculate_loss(self, unused_predictions, unused_labels, **unused_params):
        raise NotImplementedError()


class _CeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predictions, labels_u8):
        B = predictions.shape[0]
        p = predictions.contiguous()
        rows = torch.empty(B, dtype=torch.float32, device=p.device)
        out = torch.empty(1, dtype=torch.float32, device=p.device)
        ops.ce_kl_loss(p, None, labels_u8, 0.0, 0.0, rows, None, None)
        ops.reduce_rows(rows, 1.0 / B, out)
        ctx.save_for_backward(p, labels_u8)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        p, labels_u8 = ctx.saved_tensors
        dP = torch.empty_like(p)
        ops.ce_kl_loss(p, None, labels_u8, 1.0 / p.shape[0], 0.0, None, None, dP)
        return dP * grad_out, None


class CrossEntropyLoss(BaseLoss):
    """Calculate the cross entropy loss between the predictions and labels (losses.py:86-97):
    mean over the batch of sum_c -[y log(p+1e-5) + (1-y) log(1-p+1e-5)]."""

    def calculate_loss(self, predictions, labels, **unused_params):
        if tuple(predictions.shape) != tuple(labels.shape):
            raise ValueError("predictions and labels must have the same shape")
        return _CeFn.apply(predictions, _labels_u8(labels))


class _KlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher_predictions, student_predictions):
        pt, ps = teacher_predictions.contiguous(), student_predictions.contiguous()
        B = ps.shape[0]
        rows = torch.empty(B, dtype=torch.float32, device=ps.device)
        out = torch.empty(1, dtype=torch.float32, device=ps.device)
        ops.ce_kl_loss(ps, pt, None, 0.0, 0.0, None, rows, None)
        ops.reduce_rows(rows, 1.0, out)
        ctx.save_for_backward(pt, ps)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        pt, ps = ctx.saved_tensors
        dP = torch.empty_like(ps)
        ops.ce_kl_loss(ps, pt, None, 0.0, 1.0, None, None, dP)
        return None, dP * grad_out      # the teacher's predictions are constants for the student (F9)


def prediction_matching_loss(teacher_predictions, student_predictions):
    """L_PRED = tf.reduce_sum(KL_div(predictions, student_predictions)) (train.py:394-398)."""
    return _KlFn.apply(teacher_predictions.detach(), student_predictions)


class _RepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher_state, student_state):
        t, s = teacher_state.contiguous(), student_state.contiguous()
        B = s.shape[0]
        rows = torch.empty(B, dtype=torch.float32, device=s.device)
        out = torch.empty(1, dtype=torch.float32, device=s.device)
        ops.rep_loss(t, s, 0.0, rows, None)
        ops.reduce_rows(rows, 1.0 / B, out)
        ctx.save_for_backward(t, s)
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        t, s = ctx.saved_tensors
        d = torch.empty_like(s)
        rows = torch.empty(s.shape[0], dtype=torch.float32, device=s.device)
        ops.rep_loss(t, s, 2.0 / s.shape[0], rows, d)
        return None, d * grad_out


def representation_matching_loss(teacher_state, student_state):
    """L_REP = reduce_mean(reduce_sum(square(teacher_state - student_state), axis=1)) (train.py:359-362)."""
    return _RepFn.apply(teacher_state.detach(), student_state)


class _RegFn(torch.autograd.Function):
    """sum of slim.l2_regularizer(l2_penalty) over gates/experts weights = l2_penalty * sum(w^2)/2.
    Its gradient (grad_out * l2_penalty * w) is applied inside the fused clip+Adam kernel: the
    backward only records grad_out (= regularization_penalty) on the parameter store."""

    @staticmethod
    def forward(ctx, token, params):
        acc = torch.zeros(2, dtype=torch.float32, device=params.device)
        out = torch.empty(1, dtype=torch.float32, device=params.device)
        for n in (params.gates_w, params.experts_w):
            ops.sumsq(params.w[n], None, 0.0, acc[0:1], None)
        ops.reduce_rows(acc[0:1], 0.5 * params.cfg.l2_penalty, out)
        ctx.params = params
        return out.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        ctx.params.reg_grad_scale = grad_out.detach()
        return torch.zeros_like(ctx.params.token), None


def regularization_loss(scope_name: str):
    """tf.add_n(tf.losses.get_regularization_losses(scope)) (train.py:305-307,377-379)."""
    p = _scope.trainable_variables(scope_name)
    return _RegFn.apply(p.token, p)
