"""Integer / index side of the path (bit-exact) and the eval metrics against the reference's
golden vectors, through the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz"))


def _boundary_equal(p_row, got, want, k):
    if set(got) == set(want):
        return True
    kth = np.sort(p_row)[::-1][min(k, len(p_row)) - 1]
    return all(p_row[i] == kth for i in set(got) ^ set(want))


@pytest.mark.parametrize("case", ["small", "ties", "k_gt_vocab", "yt8m"])
def test_topk_against_reference_golden_and_oracle(case):
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import ops
    p, y = GOLD[case + "/predictions"], GOLD[case + "/labels"]
    want = GOLD[case + "/topk_sorted_indices"]
    idx, val, lab = ops.topk(torch.from_numpy(p).cuda(), 20, torch.from_numpy(y != 0).cuda().view(torch.uint8))
    idx, val, lab = idx.cpu().numpy(), val.cpu().numpy(), lab.cpu().numpy()
    oi, ov = O.top_k(p, 20)
    assert np.array_equal(idx, oi) and np.array_equal(val, ov)          # exact, ties -> lower class index
    assert np.array_equal(lab, np.take_along_axis(y, oi, axis=1).astype(np.uint8))
    for b in range(p.shape[0]):
        assert _boundary_equal(p[b], idx[b].tolist(), want[b].tolist(), 20)


@pytest.mark.parametrize("case", ["small", "k_gt_vocab", "yt8m"])
def test_metrics_against_reference_golden(case):
    from efficientvideoclassification_youtube8m_b200 import eval_util
    p, y = GOLD[case + "/predictions"], GOLD[case + "/labels"]
    assert abs(eval_util.calculate_hit_at_one(p, y) - float(GOLD[case + "/hit_at_one"])) < 1e-12
    assert abs(eval_util.calculate_precision_at_equal_recall_rate(p, y) - float(GOLD[case + "/perr"])) < 1e-6
    assert abs(eval_util.calculate_gap(p, y, 20) - float(GOLD[case + "/gap"])) < 1e-6
    m = eval_util.EvaluationMetrics(p.shape[1], 20)
    h = p.shape[0] // 2
    m.accumulate(p[:h], y[:h], np.ones(h))
    m.accumulate(p[h:], y[h:], np.ones(p.shape[0] - h))
    g = m.get()
    assert abs(g["gap"] - float(GOLD[case + "/epoch_gap"])) < 1e-6
    assert abs(g["avg_hit_at_one"] - float(GOLD[case + "/epoch_hit_at_one"])) < 1e-9
    with pytest.raises(ValueError):
        eval_util.top_k_by_class(p, y, 0)


def test_topk_full_size_properties():
    """BASELINE config #2 size: [1024, 4716]; sortedness, membership, agreement with torch.topk values."""
    from efficientvideoclassification_youtube8m_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.rand(1024, 4716, device="cuda", generator=g)
    idx, val, _ = ops.topk(p, 20)
    tv, ti = torch.topk(p, 20, dim=1)
    assert torch.equal(val, tv)
    assert torch.equal(torch.gather(p, 1, idx.long()), val)
    assert bool((val[:, :-1] >= val[:, 1:]).all())


def test_num_frames_student_bit_exact_all_lengths():
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import ops
    n = np.arange(0, 301, dtype=np.int32)
    for every_n in (1, 2, 3, 4, 5, 6, 10, 12, 15, 20, 30, 60):
        got = ops.num_frames_student(torch.from_numpy(n).cuda(), every_n).cpu().numpy()
        assert np.array_equal(got, O.num_frames_student(n, every_n)), every_n


def test_lstm_lengths_bit_exact():
    from efficientvideoclassification_youtube8m_b200 import ops
    n = np.arange(0, 301, dtype=np.int32)
    for chunks, ell, dt in ((20, 15, torch.int32), (5, 6, torch.int64), (5, 12, torch.int64), (5, 2, torch.int64)):
        nn = torch.from_numpy(np.minimum(n, chunks * ell)).cuda().to(dt)
        l1, l2 = ops.lstm_lengths(nn, chunks, ell)
        nv = nn.cpu().numpy().astype(np.int64)
        want1 = np.stack([np.minimum(ell, np.maximum(0, nv - ell * c)) for c in range(chunks)]).reshape(-1)
        want2 = np.ceil(nv.astype(np.float32) / np.float32(ell)).astype(np.int32)
        assert np.array_equal(l1.cpu().numpy(), want1) and np.array_equal(l2.cpu().numpy(), want2)


def test_gather_and_random_samplers_bit_exact():
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import model_utils, nn_ops, ops
    rng = np.random.default_rng(0)
    B, T, D = 6, 300, 64
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    nf = np.array([1, 7, 30, 150, 299, 300], dtype=np.int32)
    xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
    for every_n in (5, 10, 20, 30):
        got = nn_ops.sample_every_n(xd, every_n).cpu().numpy()
        assert np.array_equal(got, x[:, O.uniform_frame_indices(every_n)])      # indices exact, copy bit-exact
    u = rng.random((B, 30), dtype=np.float32)
    idx = ops.random_frame_index(torch.from_numpy(u).cuda(), nfd).cpu().numpy()
    assert np.array_equal(idx, O.random_frame_indices(u, nf))
    got = model_utils.SampleRandomFrames(xd, nfd.view(-1, 1), 30, u=torch.from_numpy(u).cuda()).cpu().numpy()
    assert np.array_equal(got, O.gather_frames(torch.from_numpy(x), idx).numpy())
    u1 = rng.random((B,), dtype=np.float32)
    sidx = ops.random_sequence_index(torch.from_numpy(u1).cuda(), nfd, 30).cpu().numpy()
    assert np.array_equal(sidx, O.random_sequence_indices(u1, nf, 30))
    got = model_utils.SampleRandomSequence(xd, nfd.view(-1, 1), 30, u=torch.from_numpy(u1).cuda()).cpu().numpy()
    assert np.array_equal(got, O.gather_frames(torch.from_numpy(x), sidx).numpy())


def test_random_uniform_bit_exact_and_empty_video_samplers():
    """evc_random_uniform == the oracle's Philox4x32-10 stream bit for bit; the samplers with their own draws
    equal the oracle's index rule on those draws; a video with num_frames = 0 never indexes before its row."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import model_utils, ops
    for n, seed, off in [(1, 0, 0), (7, 5, 0), (4096 + 3, 2 ** 40 + 17, 2 ** 33 + 5)]:
        u = torch.empty(n, dtype=torch.float32, device="cuda")
        ops.random_uniform(u, seed, off)
        assert np.array_equal(u.cpu().numpy(), O.philox_uniform(seed, off, n))
    rng = np.random.default_rng(4)
    B, K = 5, 30
    x = rng.standard_normal((B, 300, 64)).astype(np.float32)
    nf = np.array([0, 1, 29, 31, 300], dtype=np.int32)
    x[np.arange(300)[None, :] >= nf[:, None]] = 0.0
    xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
    model_utils.set_random_seed(11)
    got = model_utils.SampleRandomFrames(xd, nfd.view(-1, 1), K).cpu().numpy()
    want_idx = O.random_frame_indices(O.philox_uniform(11, 0, B * K).reshape(B, K), nf)
    assert np.array_equal(got, O.gather_frames(torch.from_numpy(x), want_idx).numpy())
    got = model_utils.SampleRandomSequence(xd, nfd.view(-1, 1), K).cpu().numpy()      # next counters of the stream
    u1 = O.philox_uniform(11, (B * K + 3) // 4, B)
    sidx = np.maximum(O.random_sequence_indices(u1, nf, K), 0)     # n = 0: the reference's -1 is clamped to frame 0
    assert np.array_equal(got, O.gather_frames(torch.from_numpy(x), sidx).numpy())
    assert np.all(got[0] == 0)                                       # the empty video samples zero frames


def test_l2_normalize_and_zero_frames():
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import nn_ops
    x, nf, _ = O.synthetic_batch(4, seed=2, num_features=1152, vocab_size=10)
    got = nn_ops.l2_normalize(torch.from_numpy(x).cuda()).cpu()
    want = O.l2_normalize(torch.from_numpy(x).double())
    assert (got.double() - want).abs().max().item() < 1e-6
    assert torch.all(got[0, int(nf[0]):] == 0)          # zero-padded frames stay zero (max(ss, 1e-12))


def test_every_n_sweep_student_forward():
    """BASELINE config #5: every_n in {5,10,20,30} -> student T in {60,30,15,10}, chunks of {12,6,3,2}."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import HLstmParams, ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import StudentEvaluator
    kw = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
    cfg = ModelConfig(**kw)
    B = 12
    x, nf, lab = O.synthetic_batch(B, seed=21, num_features=128, vocab_size=200)
    S = O.init_params("model_student", 1, dtype=torch.float64, **kw)
    params = HLstmParams("model_student", cfg, "cuda", seed=1)
    xn = O.l2_normalize(torch.from_numpy(x).double())
    for every_n in (5, 10, 20, 30):
        ev = StudentEvaluator(params, B, every_n=every_n)
        pred, idx, val, tl = ev.step(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda())
        with torch.no_grad():
            nfs = O.num_frames_student(nf, every_n)
            ss, sp = O.student_forward(O.sample_uniform(xn, every_n), nfs, S, vocab_size=200, num_mixtures=2)
        assert np.array_equal(ev.nf_student.cpu().numpy(), nfs)
        assert (pred.cpu().double() - sp).abs().max().item() < 1e-3, every_n
        # top-20 of the CUDA predictions equals the oracle's top-20 rule applied to the same predictions
        oi, _ = O.top_k(pred.cpu().numpy(), 20)
        assert np.array_equal(idx.cpu().numpy(), oi)
    with pytest.raises(ValueError):
        StudentEvaluator(params, B, every_n=7)           # 43 frames do not split into 5 chunks (F13)


def test_format_lines_matches_reference_rule():
    """inference_ensemble.py:63-74: top_k classes sorted by descending score, "%i %f" pairs."""
    from efficientvideoclassification_youtube8m_b200 import eval_util
    p = GOLD["small/predictions"]
    ids = [("vid%02d" % i).encode() for i in range(p.shape[0])]
    lines = list(eval_util.format_lines(ids, p, 5))
    for b, line in enumerate(lines):
        top = np.argpartition(p[b], -5)[-5:]
        # equal scores: the reference keeps argpartition's arbitrary order, here the lower class comes first
        want = sorted([(int(c), float(p[b][c])) for c in top], key=lambda q: (-q[1], q[0]))
        assert line == ids[b].decode() + "," + " ".join("%i %f" % pair for pair in want) + "\n"
