"""Noise floor of the 200-step loss criterion (north_star: "loss within 1 % over the first 200 steps").

Not a test: a CPU experiment that replays the joint teacher+student training curve of
tests/test_gpu_parity.py::_curve with the oracle in several arithmetic modes and reports, per loss term,
the relative deviation from the float64 oracle at every step:

  f32      the float32 oracle (what the reference's TensorFlow CPU kernels compute in)
  tf32     matmul operands rounded to 10 mantissa bits (tcgen05 kind::tf32), f32 elsewhere
  bf16     matmul operands rounded to bf16 (the default GPU path: bf16 operands, f32 accumulate)
  bf16x2   matmul operands split hi+lo in bf16, three products hi*hi + hi*lo + lo*hi (~16 mantissa bits)

    python tests/noise_floor.py [--lr 1e-3] [--steps 200] [--modes f32,tf32,bf16,bf16x2]

Output: one table per mode (max / median / 95th percentile of the relative error per term, steps above
1 %) and tests/golden/noise_floor_lr<lr>.json, which tests/test_gpu_parity.py::test_loss_curve_200_steps
uses as the measured floor.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hlstm_oracle as O  # noqa: E402

SMALL = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
TERMS = ("teacher_loss", "student_loss", "l_ce", "l_rep", "l_pred")


def _round_mantissa(x: torch.Tensor, bits: int) -> torch.Tensor:
    """Round-to-nearest-even of an f32 tensor to `bits` explicit mantissa bits."""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    half = (1 << (drop - 1)) - 1
    lsb = (i >> drop) & 1
    return ((i + half + lsb) >> drop << drop).view(torch.float32)


def _q(x: torch.Tensor, mode: str):
    if mode == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if mode == "tf32":
        return _round_mantissa(x, 10)
    return x


def _site_modes(mode: str):
    """'bf16' -> the same mode at every product; 'rec=bf16x2,moe=bf16,dgrad=...' -> per site.  Sites:
    rec_fwd / rec_dgrad / rec_wgrad (the LSTM products and their two gradient products) and moe_fwd /
    moe_dgrad / moe_wgrad; a key without suffix sets all three, 'default=' the rest."""
    sites = ["rec_fwd", "rec_dgrad", "rec_wgrad", "moe_fwd", "moe_dgrad", "moe_wgrad"]
    if "=" not in mode:
        return {k: mode for k in sites}
    spec = dict(kv.split("=") for kv in mode.split(","))
    out = {k: spec.get("default", "bf16") for k in sites}
    for k, v in spec.items():
        for s_ in sites:
            if s_ == k or s_.startswith(k + "_"):
                out[s_] = v
    return out


class _QMatmul(torch.autograd.Function):
    """a @ b with the operands of the product and of both gradient products rounded as the GPU path does."""

    @staticmethod
    def forward(ctx, a, b, modes, site):
        ctx.save_for_backward(a, b)
        ctx.modes, ctx.site = modes, site
        return _qmm(a, b, modes[site + "_fwd"])

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        return (_qmm(g, b.t(), ctx.modes[ctx.site + "_dgrad"]), _qmm(a.t(), g, ctx.modes[ctx.site + "_wgrad"]),
                None, None)


def _qmm(a, b, mode):
    if mode == "f32":
        return a @ b
    if mode == "bf16x2":
        ah, bh = _q(a, "bf16"), _q(b, "bf16")
        al, bl = _q(a - ah, "bf16"), _q(b - bh, "bf16")
        return ah @ bh + (ah @ bl + al @ bh)
    return _q(a, mode) @ _q(b, mode)


def _patched(mode):
    """oracle functions with every matmul replaced by the rounded product."""
    modes = _site_modes(mode)

    def cell(x, c, h, kernel, bias, forget_bias=1.0):
        z = _QMatmul.apply(torch.cat([x, h], dim=1), kernel, modes, "rec") + bias
        i, j, f, o = torch.chunk(z, 4, dim=1)
        c_new = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        return c_new, torch.tanh(c_new) * torch.sigmoid(o)

    def moe(state, params, scope, vocab_size, num_mixtures):
        wg = params[f"{scope}/classifier/gates/weights"]
        we = params[f"{scope}/classifier/experts/weights"]
        be = params[f"{scope}/classifier/experts/biases"]
        g = _QMatmul.apply(state, wg, modes, "moe").reshape(-1, num_mixtures + 1)
        e = (_QMatmul.apply(state, we, modes, "moe") + be).reshape(-1, num_mixtures)
        p = (torch.softmax(g, dim=1)[:, :num_mixtures] * torch.sigmoid(e)).sum(1)
        return p.reshape(-1, vocab_size)
    return cell, moe


def curve(mode: str, lr: float, steps: int, B: int = 16, NB: int = 8, cfg=SMALL):
    """Losses of `steps` joint training steps over NB rotating batches in arithmetic `mode` ('f64' = truth)."""
    dtype = torch.float64 if mode == "f64" else torch.float32
    batches = [O.synthetic_batch(B, seed=100 + i, num_features=cfg["feature_size"], vocab_size=cfg["vocab_size"])
               for i in range(NB)]
    T = O.init_params("model", 0, dtype=dtype, **cfg)
    S = O.init_params("model_student", 1, dtype=dtype, **cfg)
    ot, os_ = O.TFAdam(T, lr=lr), O.TFAdam(S, lr=lr)
    saved = O.basic_lstm_cell, O.moe_predictions
    if mode not in ("f64", "f32"):
        O.basic_lstm_cell, O.moe_predictions = _patched(mode)
    out = {k: [] for k in TERMS}
    try:
        for it in range(steps):
            x, nf, lab = batches[it % NB]
            ref = O.teacher_student_train_step(torch.from_numpy(x).to(dtype), nf, torch.from_numpy(lab), T, S, ot, os_,
                                               vocab_size=cfg["vocab_size"], num_mixtures=cfg["num_mixtures"])
            for k in TERMS:
                out[k].append(float(ref[k]))
    finally:
        O.basic_lstm_cell, O.moe_predictions = saved
    return {k: np.array(v) for k, v in out.items()}


def summarize(ref, got):
    s = {}
    for k in TERMS:
        rel = np.abs(got[k] - ref[k]) / (np.abs(ref[k]) + 1e-9)
        s[k] = {"max": float(rel.max()), "argmax": int(rel.argmax()), "median": float(np.median(rel)),
                "p95": float(np.percentile(rel, 95)), "steps_over_1pct": int((rel > 0.01).sum()),
                "first_over_1pct": int(np.argmax(rel > 0.01)) if (rel > 0.01).any() else -1}
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--modes", default="f32,tf32,bf16,bf16x2")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = curve("f64", a.lr, a.steps)
    result = {"lr": a.lr, "steps": a.steps, "config": SMALL, "batch": 16, "rotating_batches": 8,
              "f64": {k: v.tolist() for k, v in ref.items()}, "modes": {}}
    for mode in a.modes.split(";" if "=" in a.modes else ","):
        got = curve(mode, a.lr, a.steps)
        s = summarize(ref, got)
        result["modes"][mode] = {"summary": s, "rel": {k: (np.abs(got[k] - ref[k]) / (np.abs(ref[k]) + 1e-9)).tolist()
                                                         for k in TERMS}}
        print(f"== {mode} vs f64, lr {a.lr}, {a.steps} steps")
        for k in TERMS:
            print(f"  {k:13s} max {s[k]['max']:.3e} @ {s[k]['argmax']:3d}  median {s[k]['median']:.2e}  "
                  f"p95 {s[k]['p95']:.2e}  steps>1%: {s[k]['steps_over_1pct']:3d} (first {s[k]['first_over_1pct']})")
    out = a.out or os.path.join(ROOT, "tests", "golden", f"noise_floor_lr{a.lr:g}.json")
    if out != "-":
        with open(out, "w") as f:
            json.dump(result, f)
        print("wrote", out)


if __name__ == "__main__":
    main()
