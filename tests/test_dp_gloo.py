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


def _stats_from_oracle(O, p, a, loss, k):
    """batch_stats() computed on the host by the oracle (the product computes the top-k on the GPU)."""
    import numpy as np
    n = p.shape[0]
    if n == 0:
        z = np.zeros((0, k))
        return {"n": 0, "idx": z.astype(np.int32), "val": z.astype(np.float32), "lab": z.astype(np.uint8),
                "num_positives": np.zeros(p.shape[1]), "hit_sum": 0.0, "perr_sum": 0.0, "loss_sum": 0.0}
    idx, val = O.top_k(p, k)
    return {"n": n, "idx": idx, "val": val.astype(np.float32), "lab": np.take_along_axis(a, idx, axis=1),
            "num_positives": a.sum(0).astype(np.float64), "hit_sum": O.hit_at_one(p, a) * n,
            "perr_sum": O.perr(p, a) * n, "loss_sum": float(loss) * n}


def _eval_batches(world):
    """Two evaluation batches per rank, the last one ragged (rank 0: 5 videos, rank 1: none)."""
    import numpy as np
    rng = np.random.default_rng(5)
    V = 40
    out = []
    for b, sizes in enumerate([(6, 6), (5, 0)]):
        per_rank = []
        for r in range(world):
            n = sizes[r]
            p = rng.random((n, V)).astype(np.float32)        # distinct values: no tie order involved
            a = (rng.random((n, V)) < 0.1).astype(np.uint8)
            per_rank.append((p, a, 1.0 + b + 0.1 * r))
        out.append(per_rank)
    return V, out


def _metrics_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.eval_util import EvaluationMetrics
    V, batches = _eval_batches(world)
    m = EvaluationMetrics(V, 5, distributed=True)
    per_batch = []
    for per_rank in batches:
        p, a, loss = per_rank[rank]
        per_batch.append(m.accumulate_stats(_stats_from_oracle(O, p, a, loss, 5)))
    res = m.get()
    out[rank] = (per_batch, res["avg_hit_at_one"], res["avg_perr"], res["avg_loss"], res["gap"],
                 [float(x) for x in res["aps"]], m.num_examples)
    dist.destroy_process_group()


def test_distributed_evaluation_metrics_world2():
    """evaluation shards the videos over the ranks and all-gathers k triplets per video; every
    rank ends with the metrics one process computes on the concatenated batches."""
    import numpy as np
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.eval_util import EvaluationMetrics
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_metrics_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        got = dict(out)
    assert got[0] == got[1]                                   # all ranks hold the same global metrics
    V, batches = _eval_batches(world)
    single = EvaluationMetrics(V, 5)
    all_p, all_a, want_batch = [], [], []
    for per_rank in batches:
        p = np.concatenate([x[0] for x in per_rank])
        a = np.concatenate([x[1] for x in per_rank])
        loss = sum(x[2] * x[0].shape[0] for x in per_rank) / p.shape[0]
        want_batch.append(single.accumulate_stats(_stats_from_oracle(O, p, a, loss, 5)))
        all_p.append(p)
        all_a.append(a)
    want = single.get()
    assert len(np.unique(np.concatenate(all_p))) == 17 * V
    per_batch, hit, perr, loss, gap, aps, n = got[0]
    assert n == 17
    for g, w in zip(per_batch, want_batch):
        for k in w:
            assert abs(g[k] - w[k]) < 1e-12, (k, g[k], w[k])
    assert abs(hit - want["avg_hit_at_one"]) < 1e-12 and abs(perr - want["avg_perr"]) < 1e-12
    assert abs(loss - want["avg_loss"]) < 1e-12
    assert abs(gap - want["gap"]) < 1e-12 and np.allclose(aps, want["aps"], atol=1e-12)
    # and the global GAP equals the oracle's AP over the pooled top-k triplets of all 17 videos
    # (distinct prediction values: no tie order involved)
    assert abs(gap - O.gap(np.concatenate(all_p), np.concatenate(all_a), 5)) < 1e-12


def _reader_worker(rank, world, port, tmp, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from efficientvideoclassification_youtube8m_b200 import readers
    rd = readers.YT8MFrameFeatureReader(num_classes=10, feature_sizes=[8], feature_names=["rgb"])
    pat = os.path.join(tmp, "train*.tfrecord")
    res = {}
    for mode in ("min", "pad"):
        sizes, ids = [], []
        for b in readers.get_input_data_batches(rd, pat, 4, num_epochs=2, shuffle=False, rank=rank, world=world,
                                                uneven=mode, pin_memory=False):
            # a collective per batch, like the training step / distributed metrics: hangs if the ranks disagree
            t = torch.tensor([float(len(b[0]))])
            dist.all_reduce(t)
            sizes.append((len(b[0]), tuple(b[1].shape[1:]), tuple(b[2].shape[1:])))
            ids += b[0]
        res[mode] = (sizes, ids)
    out[rank] = res
    dist.destroy_process_group()


def test_uneven_shards_keep_ranks_in_step_world2(tmp_path):
    """shards hold different numbers of videos; ranks must run the same number of steps per epoch
    ('min': stop with the first rank to run out; 'pad': empty batches until the last rank is done, no video
    lost) or the per-step collectives hang."""
    import numpy as np
    from efficientvideoclassification_youtube8m_b200 import readers
    rng = np.random.default_rng(2)
    counts = [9, 2, 5]                      # rank 0 reads shards 0 and 2 (14 videos), rank 1 shard 1 (2 videos)
    for k, n in enumerate(counts):
        recs = [readers.make_sequence_example(f"s{k}v{i}", [k], {"rgb": rng.integers(0, 256, (2, 8), dtype=np.uint8)})
                for i in range(n)]
        readers.write_tfrecord(str(tmp_path / f"train{k}.tfrecord"), recs)
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_reader_worker, args=(world, _free_port(), str(tmp_path), out), nprocs=world, join=True)
        got = dict(out)
    # min: one step per epoch on both ranks (rank 1 has a single batch of 2 videos)
    assert [s[0] for s in got[0]["min"][0]] == [4, 4] and [s[0] for s in got[1]["min"][0]] == [2, 2]
    # pad: rank 0 has 4 batches per epoch (4,4,4,2); rank 1 follows with empty batches of the right shapes
    assert [s[0] for s in got[0]["pad"][0]] == [4, 4, 4, 2] * 2
    assert [s[0] for s in got[1]["pad"][0]] == [2, 0, 0, 0] * 2
    assert all(s[1:] == ((300, 8), (10,)) for r in (0, 1) for s in got[r]["pad"][0])
    ids = got[0]["pad"][1] + got[1]["pad"][1]
    assert len(ids) == 2 * sum(counts) and len(set(ids)) == sum(counts)      # every video, every epoch


def _gather_worker(rank, world, port, out):
    """The identity behind engine._classifier_wgrad_gathered, on the CPU: the reduce-scatter(AVG) row block of the
    per-rank weight gradients X_r^T dL_r equals (1/world) * X_all[:, r0:r1]^T dL_all over the all-gathered batch, with
    the row blocks of HLstmParams.row_block."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from efficientvideoclassification_youtube8m_b200.params import HLstmParams
    B, S, N = 6, 8, 10
    g = torch.Generator().manual_seed(100 + rank)
    X = torch.randn(B, S, generator=g, dtype=torch.float64)
    dL = torch.randn(B, N, generator=g, dtype=torch.float64)
    # (a) what the reduce-scatter mode leaves on this rank: the average of the local products, own row block
    local = X.t() @ dL
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    local /= world

    class P:                                   # row_block only reads the shapes
        shapes = {"w": (S, N)}
    r0, r1 = HLstmParams.row_block(P, "w", rank, world)
    # (b) the gathered-batch form: all-gather the operands, one contraction, alpha = 1 / world
    Xs, dLs = [torch.zeros_like(X) for _ in range(world)], [torch.zeros_like(dL) for _ in range(world)]
    dist.all_gather(Xs, X)
    dist.all_gather(dLs, dL)
    X_all, dL_all = torch.cat(Xs), torch.cat(dLs)
    mine = (1.0 / world) * (X_all[:, r0:r1].t() @ dL_all)
    blocks = [None] * world
    dist.all_gather_object(blocks, (r0, r1))
    out[rank] = {"ok": bool(torch.allclose(mine, local[r0:r1], rtol=1e-12, atol=1e-12)), "blocks": blocks}
    dist.destroy_process_group()


def test_gathered_batch_weight_gradient_equals_reduce_scatter_world2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_gather_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res[0]["ok"] and res[1]["ok"]
    assert res[0]["blocks"] == [(0, 4), (4, 8)]          # the row blocks tile the matrix
