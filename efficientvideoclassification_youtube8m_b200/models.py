"""Plugin root of the model families (interface of code_student_uniform/models.py:4-8).

The step scripts look models up by class name and call ``create_model`` with the batch tensor plus a bag of
keyword arguments they do not all understand (``batch_size``, ``labels``, ``dropout``, ``is_training`` --
train.py:282-288); subclasses pick what they need and ignore the rest.  Here the tensors are CUDA
``torch.Tensor``s and the returned dictionary carries at least ``"predictions"`` [batch, vocab_size].
"""
from __future__ import annotations

from typing import Any, Dict


class BaseModel(object):
    def create_model(self, unused_model_input: Any, **unused_params: Any) -> Dict[str, Any]:
        raise NotImplementedError(f"{type(self).__name__} does not implement create_model()")
