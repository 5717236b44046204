"""Run a few joint T+S training steps at the bench configuration (for ncu)."""
import sys; sys.path.insert(0, '.')
import torch
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = 256
x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
tr = TeacherStudentTrainer(ModelConfig(), batch_size=B, device="cuda", base_learning_rate=1e-5)
xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
torch.cuda.synchronize()
for _ in range(steps - 1):
    tr.step(xd, nfd, labd)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one warm step is listed
tr.step(xd, nfd, labd)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(tr.fetch())
