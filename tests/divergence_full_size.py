"""Evidence for bench.py's learning-rate note: the reference graph itself (float64 / float32
CPU oracle, full model dimensions, Adam lr 1e-3, clip 1.0) on bench.py's synthetic inputs -- uniform random uint8
features, full-length videos, random labels, NB rotating batches -- reaches inf/NaN losses within a few dozen
steps, i.e. the reference's check_numerics would abort the same run; at lr 1e-5 it stays finite.

    python tests/divergence_full_size.py [--batch 8] [--batches 8] [--steps 40] [--lr 1e-3] [--dtype f32]
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hlstm_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--batches", type=int, default=8)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--out", default=None)
a = ap.parse_args()
dt = torch.float64 if a.dtype == "f64" else torch.float32
torch.set_num_threads(os.cpu_count() or 1)
T = O.init_params("model", 0, dtype=dt)
S = O.init_params("model_student", 1, dtype=dt)
ot, os_ = O.TFAdam(T, lr=a.lr), O.TFAdam(S, lr=a.lr)
batches = [O.synthetic_batch(a.batch, seed=1234 + i, full_length=True) for i in range(a.batches)]
rows = []
for it in range(a.steps):
    x, nf, lab = batches[it % a.batches]
    t0 = time.time()
    r = O.teacher_student_train_step(torch.from_numpy(x).to(dt), nf, torch.from_numpy(lab), T, S, ot, os_)
    row = {k: float(r[k]) for k in ("teacher_loss", "student_loss", "l_ce", "l_rep", "l_pred")}
    row["min_student_p"] = float(r["student_predictions"].min())
    row["max_abs_state"] = float(r["teacher_state"].abs().max())
    rows.append(row)
    print(it, {k: f"{v:.4g}" for k, v in row.items()}, f"{time.time() - t0:.1f}s", flush=True)
    if not all(np.isfinite(v) for v in row.values()):
        print("non-finite loss at step", it)
        break
if a.out:
    json.dump({"args": vars(a), "rows": rows}, open(a.out, "w"), indent=1)
