"""Run under torchrun with N >= 2 GPUs: the sharded optimizer (reduce-scatter + row-block Adam + all-gather
of the bf16 operands) must produce the same weights as the replicated one (all-reduce + full Adam)."""
import os, sys; sys.path.insert(0, '.')
import torch, torch.distributed as dist
from oracle import hlstm_oracle as O
from efficientvideoclassification_youtube8m_b200.params import ModelConfig
from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
kw = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
cfg = ModelConfig(**kw)
B = 16
x, nf, lab = O.synthetic_batch(B, seed=500 + rank, num_features=128, vocab_size=200)
xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
a = TeacherStudentTrainer(cfg, batch_size=B, shard_optimizer=False, lstm_gain=2.0)
b = TeacherStudentTrainer(cfg, batch_size=B, shard_optimizer=True, lstm_gain=2.0)
for it in range(4):
    a.step(xd, nfd, labd); b.step(xd, nfd, labd)
fa, fb = a.fetch(), b.fetch()
worst = 0.0
for pa, pb in [(a.teacher, b.teacher), (a.student, b.student)]:
    pb.sync_master_weights(rank, world)
    for n in pa.names:
        d = (pa.w[n] - pb.w[n]).abs().max().item()
        worst = max(worst, d)
        if n in pa.shadow:
            ds = (pa.shadow[n].float() - pb.shadow[n].float()).abs().max().item()
            worst = max(worst, ds)
t = torch.tensor([worst], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("max |w_replicated - w_sharded| over all ranks:", t.item())
    print("losses replicated", {k: round(v, 5) for k, v in fa.items()})
    print("losses sharded   ", {k: round(v, 5) for k, v in fb.items()})
    assert t.item() < 2e-3, "sharded optimizer diverges from the replicated one"   # first Adam steps move weights by ~lr
    print("OK")
dist.destroy_process_group()
