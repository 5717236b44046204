"""Per-kernel ncu captures at the bench configuration (cfg #1, B=256).

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:<name> -c <n> -o gpurun_out/<piece> python scripts/prof_kernels.py <piece>

Two warm training steps run outside the profiled range; then ONE piece of the teacher's step runs
between cudaProfilerStart/Stop so that ncu only sees (and replays) those launches:
  bwd    backward recurrence of RNN_L1 cell 1 (14 recurrent dgrad GEMMs + 15 cell kernels)
  wgrad  weight-gradient GEMMs + bias column sums of RNN_L1 cell 1, then the input-gradient GEMM
  fwd    forward of RNN_L1 cell 0 (15 fused GEMM + cell-epilogue steps)
  adam   the teacher's clip + Adam pass (11 sumsq + 11 clip_adam launches)
  head   MoE logit GEMMs + the fused classifier-head kernel + its backward GEMMs
  rec    forward of RNN_L2 cell 0 (256 rows x 20 steps): the resident-weights persistent recurrence
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("EVC_OVERLAP", "0")          # one stream: the captured launches are the serial ones
import torch
from efficientvideoclassification_youtube8m_b200 import synthetic as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer

piece = sys.argv[1] if len(sys.argv) > 1 else "bwd"
B = 256
cfg = ModelConfig()
x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", base_learning_rate=1e-5)
xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
for _ in range(2):
    tr.step(xd, nfd, labd)
tr.forward_backward(xd, nfd, labd.view(torch.uint8))
torch.cuda.synchronize()
t = tr.t_eng
H, D, S = cfg.lstm_cells, cfg.feature_size, cfg.state_size
a, b = t.l1
torch.cuda.profiler.start()
if piece == "bwd":
    t._cell_bwd(b, 0, 1, H, t.len_l1, None, t.dl2_in, 2 * H, t.scr_l1)
elif piece == "wgrad":
    t._fused_norms = tr.teacher.fused_norms()      # as inside lstm_backward: |dW|^2 from the GEMM epilogue
    t._cell_wgrad(b, 0, 1, a.h_all[1:].view(-1, H), H)
    t._fused_norms = False
    t._cell_dx(b, 0, 1, H, t.dx_l1)
elif piece == "fwd":
    t._cell_fwd(a, t.x, t.R1 * D, D, 0, 0, t.len_l1)
elif piece == "rec":
    t._cell_fwd(t.l2[0], t.l2_in, B * S, S, 1, 0, t.len_l2)
elif piece == "adam":
    tr.teacher.apply_gradients(tr.lr, tr.clip, tr.penalty)
elif piece == "head":
    t.classifier_forward(mix=False)
    t.classifier_loss_fused(labd.view(torch.uint8), None, 1.0 / B, 0.0, tr.rows[0], None)
    t.classifier_backward(None, logits_done=True)
else:
    raise SystemExit("unknown piece " + piece)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", piece)
