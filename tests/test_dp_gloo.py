"""CPU suite: the data-parallel plumbing with world_size 2 on the gloo backend (no GPU): every
rank computes a different flat gradient, the step's allreduce leaves the mean on all ranks, and
replicated Adam states stay bit-identical."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from efficientvideoclassification_youtube8m_b200.steps import _Base

    class P:                       # stands in for HLstmParams: the allreduce only touches flat_g
        flat_g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    # gloo has no AVG: the step falls back to SUM / world
    b = _Base.__new__(_Base)
    b._pending = []
    _Base._allreduce(b, P, 0, 4)
    _Base._allreduce(b, P, 4, None)       # the step reduces the flat buffer in two slices
    b._finish_allreduce()
    want = torch.arange(10, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(P.flat_g, want)
    gathered = [torch.zeros(10) for _ in range(world)]
    dist.all_gather(gathered, P.flat_g)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    out[rank] = bool(ok and same and _Base._world() == world)
    dist.destroy_process_group()


def test_gradient_average_world2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
