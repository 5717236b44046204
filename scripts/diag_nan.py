import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = ModelConfig()
x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda")
xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda().view(torch.uint8)
def chk(tag):
    bad = []
    for eng, nm in [(tr.t_eng,'T'),(tr.s_eng,'S')]:
        for k in ['x','state','G','E','pred','dP','dstate','dx_l2','dl2_in','dx_l1','dG','dE']:
            t = getattr(eng,k).float()
            if not torch.isfinite(t).all(): bad.append(f"{nm}.{k}")
        for li,l in enumerate(eng.l1+eng.l2):
            for k in ['h_all','c_all','gates','dz']:
                t = getattr(l,k).float()
                if not torch.isfinite(t).all(): bad.append(f"{nm}.layer{li}.{k}")
        for n in eng.p.names:
            for d,dn in [(eng.p.g,'g'),(eng.p.w,'w'),(eng.p.m,'m'),(eng.p.v,'v')]:
                if not torch.isfinite(d[n]).all(): bad.append(f"{nm}.{dn}.{n.split('/')[-4:]}" )
    print(tag, "BAD:", bad, "losses", tr.losses.tolist()[:4], "normsq T", tr.teacher.normsq.tolist(), flush=True)
    return bad
for it in range(6):
    tr.forward_backward(xd, nfd, labd); torch.cuda.synchronize()
    if chk(f"it{it} after fwd/bwd"): break
    tr.apply_gradients(); torch.cuda.synchronize()
    if chk(f"it{it} after adam"): break
