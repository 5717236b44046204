"""Generate golden vectors for the top-k / GAP@20 / hit@1 / PERR part of the hot path by
running the REFERENCE's own numpy code (code_student_uniform/eval_util.py:17-124 and
average_precision_calculator.py) in this container.

Run once, here (needs /root/reference; it does not exist on the GPU box):
    python tests/golden/make_golden_eval.py
The reference imports ``tensorflow.python.platform.gfile`` at module scope but never
uses it (eval_util.py:8), so an empty stub module is registered before the import.
Outputs: tests/golden/eval_golden.npz (committed).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/code_student_uniform"


def _import_reference_eval_util():
    for name in ("tensorflow", "tensorflow.python", "tensorflow.python.platform",
                 "tensorflow.python.platform.gfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tensorflow.python.platform"].gfile = sys.modules["tensorflow.python.platform.gfile"]
    sys.path.insert(0, REF)
    import eval_util  # noqa: E402  (the reference module)
    return eval_util


def make_case(seed, batch, vocab, ties):
    rng = np.random.default_rng(seed)
    p = rng.random((batch, vocab), dtype=np.float32)
    if ties:  # quantise so that equal predictions straddle the top-k boundary
        p = np.round(p * 64).astype(np.float32) / 64
    y = np.zeros((batch, vocab), dtype=np.float32)
    for b in range(batch):
        k = max(1, int(rng.poisson(3.4)))
        y[b, rng.choice(vocab, size=min(k, vocab), replace=False)] = 1
        # make the model "good": raise the predictions of some true labels
        hot = np.flatnonzero(y[b])[: max(1, k // 2)]
        p[b, hot] = np.minimum(p[b, hot] + 0.5, 1.0).astype(np.float32)
    return p, y


def main():
    eu = _import_reference_eval_util()
    out = {}
    cases = [("small", 11, 16, 100, False), ("ties", 12, 16, 100, True),
             ("k_gt_vocab", 13, 4, 12, False), ("yt8m", 14, 8, 4716, False)]
    for name, seed, batch, vocab, ties in cases:
        p, y = make_case(seed, batch, vocab, ties)
        k = 20
        sets = np.full((batch, min(k, vocab)), -1, dtype=np.int32)
        for b in range(batch):
            trip = eu.top_k_triplets(p[b], y[b], k)
            sets[b] = np.sort(np.array([t[0] for t in trip], dtype=np.int32))
        out[name + "/predictions"] = p
        out[name + "/labels"] = y
        out[name + "/topk_sorted_indices"] = sets
        out[name + "/hit_at_one"] = np.float64(eu.calculate_hit_at_one(p, y))
        out[name + "/perr"] = np.float64(eu.calculate_precision_at_equal_recall_rate(p, y))
        out[name + "/gap"] = np.float64(eu.calculate_gap(p, y, top_k=k))
        m = eu.EvaluationMetrics(vocab, k)
        m.accumulate(p[: batch // 2], y[: batch // 2], np.ones(batch // 2))
        m.accumulate(p[batch // 2:], y[batch // 2:], np.ones(batch - batch // 2))
        g = m.get()
        out[name + "/epoch_gap"] = np.float64(g["gap"])
        out[name + "/epoch_hit_at_one"] = np.float64(g["avg_hit_at_one"])
        out[name + "/epoch_perr"] = np.float64(g["avg_perr"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()
                          if "predictions" not in k and "labels" not in k})


if __name__ == "__main__":
    main()
