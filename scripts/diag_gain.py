import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
SMALL = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
for cfgkw,B in [(SMALL,24)]:
  for gain in [1.0,2.0,3.0,4.0,6.0]:
    for stress in [False, True]:
        cfg = ModelConfig(**cfgkw)
        x, nf, lab = O.synthetic_batch(B, seed=7, num_features=cfg.feature_size, vocab_size=cfg.vocab_size, stress=stress)
        tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", lstm_gain=gain)
        T = O.init_params("model", 0, dtype=torch.float64, gain=gain, **cfgkw)
        S = O.init_params("model_student", 1, dtype=torch.float64, gain=gain, **cfgkw)
        # oracle with bf16-rounded weights too (to separate operand rounding from logic)
        xd = torch.from_numpy(x).cuda(); nfd = torch.from_numpy(nf).cuda(); labd = torch.from_numpy(lab).cuda()
        tr.forward_backward(xd, nfd, labd.view(torch.uint8)); torch.cuda.synchronize()
        ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, None, None, clip_gradient_norm=0.0, regularization_penalty=0.0, vocab_size=cfg.vocab_size, num_mixtures=cfg.num_mixtures)
        es = (tr.t_eng.state.cpu().double()-ref["teacher_state"]).abs().max().item()
        ep = (tr.t_eng.pred.cpu().double()-ref["teacher_predictions"]).abs().max().item()
        ess = (tr.s_eng.state.cpu().double()-ref["student_state"]).abs().max().item()
        eps_ = (tr.s_eng.pred.cpu().double()-ref["student_predictions"]).abs().max().item()
        smax = ref["teacher_state"].abs().max().item(); pr = (ref["teacher_predictions"].min().item(), ref["teacher_predictions"].max().item())
        rels = []
        for params, grads in [(tr.teacher, ref["teacher_grads"]), (tr.student, ref["student_grads"])]:
            for n in params.names:
                g, r = params.g[n].cpu().double(), grads[n]
                rels.append(((g-r).norm()/(r.norm()+1e-30)).item())
        print(f"gain {gain} stress {stress}: T state err {es:.2e} (max|s| {smax:.2f}) pred err {ep:.2e} range {pr[0]:.3f}-{pr[1]:.3f}; S state {ess:.2e} pred {eps_:.2e}; grad rel max {max(rels):.2e}")
