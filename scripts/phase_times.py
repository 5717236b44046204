"""Warm-cache, in-situ time of every phase of the joint training step (cfg #1, B=256, one stream, PDL on):
CUDA events between the phases of the same launch sequence `TeacherStudentTrainer.step` issues
(EVC_OVERLAP=0).  ncu's per-launch times are cold-cache and serialised; these are not.

    python scripts/phase_times.py [steps]        # on a B200; prints a table (ms per step, share)
"""
import collections
import os
import sys

os.environ["EVC_OVERLAP"] = "0"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from efficientvideoclassification_youtube8m_b200 import ops
from efficientvideoclassification_youtube8m_b200 import synthetic as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import MAX_FRAMES, TeacherStudentTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, cfg = 256, ModelConfig()
x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", base_learning_rate=1e-5)
xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
labd = torch.from_numpy(lab).cuda().view(torch.uint8)
H, D, S = cfg.lstm_cells, cfg.feature_size, cfg.state_size

marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def forward(tag, e, src, frame_idx, num_frames):
    R1, ell, C = e.R1, e.ell, e.C
    ops.frames_pack(src, frame_idx, e.K, C, True, out_bf16=e.x)
    ops.lstm_lengths(num_frames, C, ell, e.len_l1, e.len_l2)
    mark(tag + " fwd pack")
    a, b = e.l1
    e._cell_fwd(a, e.x, R1 * D, D, 0, 0, e.len_l1)
    mark(tag + " fwd L1 cell0")
    e._cell_fwd(b, a.h_all[1:], R1 * H, H, 0, 1, e.len_l1)
    mark(tag + " fwd L1 cell1")
    ops.state_pack(a.c_all[ell], a.h_all[ell], b.c_all[ell], b.h_all[ell], R1, H, out_bf16=e.l2_in)
    a2, b2 = e.l2
    e._cell_fwd(a2, e.l2_in, B * S, S, 1, 0, e.len_l2)
    e._cell_fwd(b2, a2.h_all[1:], B * H, H, 1, 1, e.len_l2)
    ops.state_pack(a2.c_all[C], a2.h_all[C], b2.c_all[C], b2.h_all[C], B, H, out_bf16=e.state_bf16, out_f32=e.state)
    mark(tag + " fwd L2 (both cells)")
    e.classifier_forward(mix=False)
    mark(tag + " fwd MoE GEMMs")


def lstm_backward(tag, e):
    R1, ell, C = e.R1, e.ell, e.C
    a2, b2 = e.l2
    a, b = e.l1
    e._cell_bwd(b2, 1, 1, H, e.len_l2, None, e.dstate, 2 * H, e.scr_l2)
    e._cell_wgrad(b2, 1, 1, a2.h_all[1:].view(-1, H), H)
    e._cell_dx(b2, 1, 1, H, e.dx_l2)
    e._cell_bwd(a2, 1, 0, S, e.len_l2, e.dx_l2, e.dstate, 0, e.scr_l2)
    e._cell_wgrad(a2, 1, 0, e.l2_in.view(-1, S), S)
    e._cell_dx(a2, 1, 0, S, e.dl2_in)
    mark(tag + " bwd L2 (all)")
    e._cell_bwd(b, 0, 1, H, e.len_l1, None, e.dl2_in, 2 * H, e.scr_l1)
    mark(tag + " bwd L1 cell1 recurrence")
    e._cell_wgrad(b, 0, 1, a.h_all[1:].view(-1, H), H)
    mark(tag + " bwd L1 cell1 wgrad+colsum")
    e._cell_dx(b, 0, 1, H, e.dx_l1)
    mark(tag + " bwd L1 cell1 dX")
    e._cell_bwd(a, 0, 0, D, e.len_l1, e.dx_l1, e.dl2_in, 0, e.scr_l1)
    mark(tag + " bwd L1 cell0 recurrence")
    e._cell_wgrad(a, 0, 0, e.x.view(-1, D), D)
    mark(tag + " bwd L1 cell0 wgrad+colsum")


def one_step():
    t, s = tr.t_eng, tr.s_eng
    mark("start")
    forward("T", t, xd, None, nfd)
    ops.num_frames_student(nfd, tr.every_n, MAX_FRAMES, tr.nf_student)
    forward("S", s, xd, tr.frame_idx, tr.nf_student)
    t.classifier_loss_fused(labd, None, 1.0 / B, 0.0, tr.rows[0], None)
    t.classifier_backward(None, logits_done=True)
    mark("T head loss + MoE bwd")
    lstm_backward("T", t)
    ops.rep_loss(t.state, s.state, 4.0 / B, tr.rows[3], s.dstate)
    s.classifier_loss_fused(labd, t.pred, 1.0 / B, 1.0, tr.rows[1], tr.rows[2])
    s.classifier_backward(None, dstate_preset=True, logits_done=True)
    mark("S head loss + MoE bwd")
    lstm_backward("S", s)
    tr.teacher.apply_gradients(tr.lr, tr.clip, tr.penalty)
    mark("T clip+Adam")
    tr.student.apply_gradients(tr.lr, tr.clip, tr.penalty)
    mark("S clip+Adam")


for _ in range(3):
    tr.step(xd, nfd, labd)
torch.cuda.synchronize()
acc = collections.OrderedDict()
for _ in range(steps):
    marks.clear()
    one_step()
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
tot = sum(acc.values())
print(f"{'phase':34s} {'ms/step':>8s} {'share':>6s}")
for k, v in acc.items():
    print(f"{k:34s} {v / steps:8.3f} {100 * v / tot:5.1f}%")
print(f"{'total':34s} {tot / steps:8.3f}")
