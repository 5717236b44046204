"""tf.variable_scope / variable store replacement.

In the reference, ``with tf.variable_scope("model"): model.create_model(...)`` creates the 11
variables of that scope on first use and reuses them afterwards (train.py:281,349).  Here the
outermost scope name selects one :class:`HLstmParams` (created with TF's default initialisers on
first use) and the execution plans (:class:`HLstmEngine`) built for it."""
from __future__ import annotations

import contextlib
from typing import Dict, Tuple

import torch

from .engine import HLstmEngine
from .params import HLstmParams, ModelConfig

_stack = []
_params: Dict[str, HLstmParams] = {}
_engines: Dict[Tuple, HLstmEngine] = {}
_seeds: Dict[str, int] = {"model": 0, "model_student": 1}


@contextlib.contextmanager
def variable_scope(name: str):
    _stack.append(name)
    try:
        yield "/".join(_stack)
    finally:
        _stack.pop()


def current_scope() -> str:
    return "/".join(_stack)


def root_scope() -> str:
    if not _stack:
        raise RuntimeError("create_model must be called inside variable_scope(...) (train.py:281,349)")
    return _stack[0]


_engine_stack = []


@contextlib.contextmanager
def use_engine(engine):
    """Makes the enclosing model's execution plan visible to the classifier plugin."""
    _engine_stack.append(engine)
    try:
        yield engine
    finally:
        _engine_stack.pop()


def current_engine():
    return _engine_stack[-1] if _engine_stack else None


def set_initializer_seed(scope: str, seed: int) -> None:
    _seeds[scope] = seed


def get_params(scope: str, cfg: ModelConfig, device) -> HLstmParams:
    p = _params.get(scope)
    if p is None:
        p = HLstmParams(scope, cfg, device, seed=_seeds.get(scope, len(_params)))
        _params[scope] = p
    elif p.cfg != cfg:
        raise ValueError(f"variable scope '{scope}' already holds variables of a different shape "
                         f"({p.cfg} vs {cfg}); TF would raise the same reuse error")
    return p


def register_params(p: HLstmParams) -> None:
    _params[p.scope] = p


def get_engine(p: HLstmParams, batch: int, frames: int, chunks: int, training: bool) -> HLstmEngine:
    key = (p.scope, batch, frames, chunks, training)
    e = _engines.get(key)
    if e is None or e.p is not p:
        e = HLstmEngine(p, batch, frames, chunks, training)
        _engines[key] = e
    return e


def trainable_variables(scope: str) -> HLstmParams:
    """tf.get_collection(TRAINABLE_VARIABLES, scope) for an explicit scope."""
    return _params[scope]


def reset_default_graph() -> None:
    _stack.clear()
    _params.clear()
    _engines.clear()
