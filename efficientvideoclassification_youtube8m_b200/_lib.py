"""ctypes binding of libevc.so (include/evc.h).  There is no fallback: if the shared
library is missing the import of any product module fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EVC_LIB_PATH: load another build of the same sources (scripts/exp_epilogue_ablation.py builds one with
# -DEVC_ABLATE); the product always uses the in-tree library
LIB_PATH = os.environ.get("EVC_LIB_PATH") or os.path.join(_HERE, "libevc.so")

P, I, L, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/evc.h
SIGNATURES = {
    "evc_version": [],
    "evc_last_error": [],
    "evc_launch_count": [],
    "evc_debug_set": [I],
    "evc_frames_pack": [P, I, I, I, P, I, I, I, I, P, P, P, P],
    "evc_frames_pack_u8": [P, P, I, I, I, P, I, I, I, I, P, P, P, P],
    "evc_num_frames_student": [P, I, I, I, P, P],
    "evc_lstm_lengths": [P, I, I, I, I, P, P, P],
    "evc_random_frame_index": [P, P, I, I, P, P],
    "evc_random_sequence_index": [P, P, I, I, P, P],
    "evc_sampled_lengths": [P, I, I, P, P],
    "evc_random_uniform": [C.c_ulonglong, C.c_ulonglong, P, L, P],
    "evc_gemm_bf16": [P, I, L, P, I, L, I, I, I, P, I, L, P, I, I, P],
    "evc_gemm_bf16x2": [P, P, I, L, P, P, I, L, I, I, I, P, I, L, P, I, I, P],
    "evc_gemm_bf16_wgrad": [P, P, I, L, P, P, I, L, I, I, I, P, L, F, P, P],
    "evc_lstm_seq_fwd": [P, L, I, P, P, I, I, I, P, P, P, P, P, L, P],
    "evc_lstm_seq_fwd_steps": [P, L, I, P, P, I, I, I, I, I, P, P, P, P, P, L, P, P, P, P, P],
    "evc_lstm_workspace_bytes": [I, I, I, I],
    "evc_lstm_rec_workspace_bytes": [I, I, I],
    "evc_lstm_seq_fwd_resident": [P, L, I, P, P, I, I, I, P, P, P, P, P, L, P],
    "evc_lstm_seq_bwd": [P, I, I, I, I, P, P, P, P, P, L, P, L, P, P, P, P, P, L, P, P, P, P],
    "evc_state_pack": [P, P, P, P, I, I, P, P, P, P, P, P],
    "evc_cast_bf16": [P, L, I, I, P, P, P],
    "evc_fill_f32": [P, L, F, P],
    "evc_moe_mix_fwd": [P, L, P, L, I, I, I, P, P],
    "evc_moe_mix_bwd": [P, L, P, L, P, I, I, I, P, L, P, L, P, P, P],
    "evc_ce_kl_loss": [P, P, P, I, I, F, F, P, P, P, P],
    "evc_moe_mix_loss": [P, L, P, L, P, P, I, I, I, F, F, P, P, P, P, L, P, L, P, P, P],
    "evc_reduce_rows": [P, I, F, P, P],
    "evc_adam_lr": [P, F, F, F, P, P],
    "evc_rep_loss": [P, P, I, I, F, P, P, P],
    "evc_colsum_bf16": [P, L, I, L, P, P],
    "evc_sumsq": [P, P, F, L, P, P, P],
    "evc_clip_adam": [P, P, P, P, L, P, F, F, P, F, F, F, P, I, L, P, P],
    "evc_clip_adam_fused": [P, P, P, P, L, P, F, F, P, F, F, F, P, I, L, P, P, P, P, P, P],
    "evc_reg_cross": [P, L, P, P, L, P, I, I, P, P],
    "evc_batch_metrics": [P, P, I, I, I, P, P, P, P, P, P, P, P, P, P, P],
    "evc_topk": [P, I, I, I, P, P, P, P, P],
}
_RESTYPES = {"evc_last_error": C.c_char_p, "evc_launch_count": C.c_longlong,
             "evc_lstm_workspace_bytes": C.c_longlong, "evc_lstm_rec_workspace_bytes": C.c_longlong}


class EvcError(RuntimeError):
    pass


def _load(path: str = LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


lib = _load()
if os.environ.get("EVC_DEBUG"):          # profiling experiments only (ablation bits of evc_debug_set)
    lib.evc_debug_set(int(os.environ["EVC_DEBUG"]))


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise EvcError(f"{what or 'libevc'} failed ({rc}): {lib.evc_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device).  The library is used
    by one process per GPU with `torch.cuda.set_device(local_rank)` called first (bench.py, the launchers):
    kernels run on the CURRENT device, so tensors of another device are rejected by `ops._cuda`."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib.evc_launch_count())
