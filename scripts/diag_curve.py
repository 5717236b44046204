"""Loss-curve parity over many steps (north_star: loss within 1 % over the first 200 steps)."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
kw = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
B, NB = 16, 8
cfg = ModelConfig(**kw)
batches = [O.synthetic_batch(B, seed=100 + i, num_features=128, vocab_size=200) for i in range(NB)]
tr = TeacherStudentTrainer(cfg, batch_size=B, base_learning_rate=lr)
T = O.init_params("model", 0, dtype=torch.float64, **kw); S = O.init_params("model_student", 1, dtype=torch.float64, **kw)
ot, os_ = O.TFAdam(T, lr=lr), O.TFAdam(S, lr=lr)
worst = {}
hist = {}
for it in range(steps):
    x, nf, lab = batches[it % NB]
    tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda())
    got = tr.fetch()
    ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, ot, os_, vocab_size=200, num_mixtures=2)
    rel = {k: abs(got[k] - float(ref[k])) / (abs(float(ref[k])) + 1e-9) for k in ("teacher_loss", "student_loss", "l_ce", "l_rep", "l_pred")}
    for k, v in rel.items():
        worst[k] = max(worst.get(k, 0), v)
        hist.setdefault(k, []).append(v)
    if it % 20 == 0 or it == steps - 1:
        print(it, {k: f"{got[k]:.4f}/{float(ref[k]):.4f}" for k in ("teacher_loss", "student_loss", "l_rep", "l_pred")}, {k: f"{v:.2e}" for k, v in rel.items()}, flush=True)
print("worst rel", {k: f"{v:.3e}" for k, v in worst.items()})
print("median rel", {k: f"{float(np.median(v)):.3e}" for k, v in hist.items()})
print("frac of steps within 1%", {k: f"{float(np.mean(np.array(v) < 0.01)):.3f}" for k, v in hist.items()})
