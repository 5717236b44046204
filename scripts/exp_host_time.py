"""Host-side (launch) time of one training step vs its GPU time."""
import sys, time; sys.path.insert(0, '.')
import torch
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer, StudentEvaluator
B = 256
x, nf, lab = O.synthetic_batch(B, seed=1234, full_length=True)
tr = TeacherStudentTrainer(ModelConfig(), batch_size=B, device="cuda", base_learning_rate=1e-5)
xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
for _ in range(3): tr.step(xd, nfd, labd)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): tr.step(xd, nfd, labd)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"train: host launch {1e3*(t1-t0)/10:.2f} ms/step (queue back-pressure included), total {1e3*(t2-t0)/10:.2f} ms/step")
hs = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.step(xd, nfd, labd); hs.append(time.perf_counter() - t0)
torch.cuda.synchronize()
print("train: host time of one step issued into an empty queue: %.2f ms (min of 5)" % (1e3 * min(hs)))
ev = StudentEvaluator(tr.student, 1024)
xi, nfi, _ = O.synthetic_batch(1024, seed=99, full_length=True)
dxi, dnfi = torch.from_numpy(xi).cuda(), torch.from_numpy(nfi).cuda()
for _ in range(3): ev.step(dxi, dnfi)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): ev.step(dxi, dnfi)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"student infer B=1024: host launch {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
