"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs and weights.  Tolerances: indices / lengths / top-k exact; predictions 1e-3 abs
(north_star: FP32-accumulate with BF16 GEMM operands); losses 1 %; gradients 3 % (l2, relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)


def _oracle_params(O, scope, seed, cfgkw, gain):
    return O.init_params(scope, seed, dtype=torch.float64, gain=gain, **cfgkw)


def _setup(cfgkw, B, gain, seed=7, stress=False):
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    cfg = ModelConfig(**cfgkw)
    x, nf, lab = O.synthetic_batch(B, seed=seed, num_features=cfg.feature_size, vocab_size=cfg.vocab_size,
                                   stress=stress)
    tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", lstm_gain=gain)
    T = _oracle_params(O, "model", 0, cfgkw, gain)
    S = _oracle_params(O, "model_student", 1, cfgkw, gain)
    # identical f32 weights on both sides
    for n in tr.teacher.names:
        assert torch.equal(tr.teacher.w[n].cpu().double(), T[n].float().double())
    return O, cfg, tr, (x, nf, lab), T, S


# gain scales the LSTM kernels: 1.0 is the reference's own initial distribution (all gates near
# sigma(0)); 2.0 drives the gates out of the linear regime (|state| up to ~0.9, predictions 0.24-0.44).
# Larger gains make the recurrence chaotic and amplify the bf16 operand rounding beyond 1e-3
# (measured table in DESIGN.md), so they are not parity cases.
@pytest.mark.parametrize("gain,stress", [(1.0, False), (2.0, True)])
def test_forward_backward_small(gain, stress):
    B = 24
    O, cfg, tr, (x, nf, lab), T, S = _setup(SMALL, B, gain, stress=stress)
    xd = torch.from_numpy(x).cuda()
    nfd = torch.from_numpy(nf).cuda()
    labd = torch.from_numpy(lab).cuda()
    tr.forward_backward(xd, nfd, labd.view(torch.uint8))
    torch.cuda.synchronize()
    kw = dict(vocab_size=cfg.vocab_size, num_mixtures=cfg.num_mixtures)
    ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, None, None,
                                       clip_gradient_norm=0.0, regularization_penalty=0.0, **kw)
    # integer side: exact
    assert np.array_equal(tr.nf_student.cpu().numpy(), ref["num_frames_student"])
    # forward
    for name, mine, theirs in [("teacher_state", tr.t_eng.state, ref["teacher_state"]),
                               ("student_state", tr.s_eng.state, ref["student_state"])]:
        err = (mine.cpu().double() - theirs).abs().max().item()
        assert err < 2e-2, (name, err)
    for name, mine, theirs in [("teacher_pred", tr.t_eng.pred, ref["teacher_predictions"]),
                               ("student_pred", tr.s_eng.pred, ref["student_predictions"])]:
        err = (mine.cpu().double() - theirs).abs().max().item()
        assert err < 1e-3, (name, err)
    v = tr.losses.cpu().tolist()
    for got, key in zip(v[:4], ["teacher_ce", "l_ce", "l_pred", "l_rep"]):
        want = float(ref[key])
        assert abs(got - want) <= 0.01 * abs(want) + 1e-5, (key, got, want)
    # gradients (unclipped, reg term excluded on both sides)
    for params, grads in [(tr.teacher, ref["teacher_grads"]), (tr.student, ref["student_grads"])]:
        for n in params.names:
            g, r = params.g[n].cpu().double(), grads[n]
            den = r.norm().item()
            if den < 1e-12:
                assert g.norm().item() < 1e-6, n
                continue
            rel = ((g - r).norm() / den).item()
            assert rel < 3e-2, (n, rel, den)


def test_forward_full_size_edges():
    """Full model dimensions (1152-d, 1024 cells, 4716 classes), edge-case lengths forced in."""
    full = dict(feature_size=1152, lstm_cells=1024, vocab_size=4716, num_mixtures=2)
    B = 16
    O, cfg, tr, (x, nf, lab), T, S = _setup(full, B, 2.0, stress=True)
    xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
    tr.t_eng.forward(xd, None, True, nfd)
    from efficientvideoclassification_youtube8m_b200 import ops
    ops.num_frames_student(nfd, 10, 300, tr.nf_student)
    tr.s_eng.forward(xd, tr.frame_idx, True, tr.nf_student)
    torch.cuda.synchronize()
    xn = O.l2_normalize(torch.from_numpy(x).double())
    with torch.no_grad():
        ts, tp = O.teacher_forward(xn, nf, T)
        nfs = O.num_frames_student(nf, 10)
        ss, sp = O.student_forward(O.sample_uniform(xn, 10), nfs, S)
    assert np.array_equal(tr.nf_student.cpu().numpy(), nfs)
    assert (tr.t_eng.pred.cpu().double() - tp).abs().max().item() < 1e-3
    assert (tr.s_eng.pred.cpu().double() - sp).abs().max().item() < 1e-3
    assert (tr.t_eng.state.cpu().double() - ts).abs().max().item() < 2e-2
    assert (tr.s_eng.state.cpu().double() - ss).abs().max().item() < 2e-2


def test_train_steps_loss_curve_small():
    """A few joint T+S steps (clip + TF-Adam): losses within 1 % of the float64 oracle each step."""
    B, steps = 16, 6
    O, cfg, tr, (x, nf, lab), T, S = _setup(SMALL, B, 2.0, stress=True)
    opt_t, opt_s = O.TFAdam(T), O.TFAdam(S)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    kw = dict(vocab_size=cfg.vocab_size, num_mixtures=cfg.num_mixtures)
    for it in range(steps):
        tr.step(xd, nfd, labd)
        got = tr.fetch()
        ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S,
                                           opt_t, opt_s, **kw)
        for key in ["teacher_loss", "student_loss", "l_ce", "l_rep"]:
            want = float(ref[key])
            assert abs(got[key] - want) <= 0.01 * abs(want) + 1e-4, (it, key, got[key], want)
    assert got["global_step"] == 2 * steps


def test_student_finetune_step_cfg4_like():
    """train_finetune.py step (final_loss = penalty*reg + L_CE) with 4 mixtures (BASELINE config #4 shape
    family: moe_num_mixtures 4, clip_gradient_norm 1.0), checked against the oracle for 4 steps."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import StudentFinetuneTrainer
    kw = dict(feature_size=128, lstm_cells=256, vocab_size=120, num_mixtures=4)
    cfg = ModelConfig(**kw)
    B = 16
    x, nf, lab = O.synthetic_batch(B, seed=31, num_features=128, vocab_size=120, stress=True)
    tr = StudentFinetuneTrainer(cfg, batch_size=B, lstm_gain=2.0)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=2.0, **kw)
    opt = O.TFAdam(S)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    for it in range(4):
        tr.step(xd, nfd, labd)
        got = tr.fetch()
        ref = O.student_finetune_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), S, opt,
                                      vocab_size=120, num_mixtures=4)
        for k in ("l_ce", "student_loss"):
            want = float(ref[k])
            assert abs(got[k] - want) <= 0.01 * abs(want) + 1e-4, (it, k, got[k], want)
        err = (tr.s_eng.pred.cpu().double() - ref["student_predictions"]).abs().max().item()
        assert err < 1e-3, (it, err)
    assert got["global_step"] == 4


def test_quantized_input_path_matches_dequantized():
    """uint8 tfrecord features -> Dequantize (utils.py:9-25) + zero padding (readers.py:173) fused into the pack
    kernel: bit-identical to feeding the dequantised float32 batch."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import ops
    from efficientvideoclassification_youtube8m_b200.params import HLstmParams, ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherEvaluator
    B, D = 8, 128
    x, nf, _ = O.synthetic_batch(B, seed=41, num_features=D, vocab_size=50)
    q = O.quantized_batch(B, seed=41, num_features=D)
    assert np.array_equal(np.where(np.arange(300)[None, :, None] < nf[:, None, None], O.dequantize(q), 0), x)
    qd, nfd = torch.from_numpy(q).cuda(), torch.from_numpy(nf).cuda()
    out = torch.empty(B, 300, D, device="cuda")
    ops.frames_pack_u8(qd, nfd, None, 300, 1, False, out_f32=out)
    assert np.array_equal(out.cpu().numpy(), x)                      # Dequantize + padding bit-exact
    cfg = ModelConfig(feature_size=D, lstm_cells=128, vocab_size=50, num_mixtures=2)
    ev = TeacherEvaluator(HLstmParams("model", cfg, "cuda", seed=0), B)
    p_q = ev.step(qd, nfd)[0].clone()
    p_f = ev.step(torch.from_numpy(x).cuda(), nfd)[0]
    assert torch.equal(p_q, p_f)


def test_checkpoint_roundtrip_and_student_conversion(tmp_path):
    """name -> array checkpoints keyed by the TF variable names; T+S -> S conversion
    (train_convert_model.py:496-517 keeps the 11 model_student/* variables)."""
    from efficientvideoclassification_youtube8m_b200.params import HLstmParams, ModelConfig
    cfg = ModelConfig(feature_size=128, lstm_cells=128, vocab_size=50, num_mixtures=2)
    s = HLstmParams("model_student", cfg, "cuda", seed=5)
    path = str(tmp_path / "student.npz")
    s.save(path)
    z = np.load(path)
    assert sorted(z.files) == sorted(s.names)
    assert "model_student/RNN_L2/rnn/multi_rnn_cell/cell_1/basic_lstm_cell/bias" in z.files
    s2 = HLstmParams("model_student", cfg, "cuda", seed=None)
    s2.load(path)
    for n in s.names:
        assert torch.equal(s.w[n], s2.w[n])
        if n in s.shadow:
            assert torch.equal(s.shadow[n], s2.shadow[n])
    t = HLstmParams("model", cfg, "cuda", seed=None)
    t.load(path, from_scope="model_student")                         # re-home under another scope
    assert torch.equal(t.w[t.names[0]], s.w[s.names[0]])
    with pytest.raises(ValueError):
        HLstmParams("model_student", ModelConfig(feature_size=128, lstm_cells=256, vocab_size=50), "cuda",
                    seed=None).load(path)


def _curve(lr, steps, precise=False):
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B, NB = 16, 8
    cfg = ModelConfig(**SMALL)
    batches = [O.synthetic_batch(B, seed=100 + i, num_features=cfg.feature_size, vocab_size=cfg.vocab_size)
               for i in range(NB)]
    tr = TeacherStudentTrainer(cfg, batch_size=B, base_learning_rate=lr, precise=precise)
    T = O.init_params("model", 0, dtype=torch.float64, **SMALL)
    S = O.init_params("model_student", 1, dtype=torch.float64, **SMALL)
    ot, os_ = O.TFAdam(T, lr=lr), O.TFAdam(S, lr=lr)
    rel = {k: [] for k in ("teacher_loss", "student_loss", "l_ce", "l_rep", "l_pred")}
    for it in range(steps):
        x, nf, lab = batches[it % NB]
        tr.step(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda())
        got = tr.fetch()
        ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, ot, os_,
                                           vocab_size=cfg.vocab_size, num_mixtures=cfg.num_mixtures)
        for k in rel:
            # (1e-3 absolute: L_PRED starts at 2e-4, where the float32 graph itself is 0.5 % off the float64 one)
            rel[k].append(max(abs(got[k] - float(ref[k])) - 1e-5, 0.0) / (abs(float(ref[k])) + 1e-9))
    return {k: np.array(v) for k, v in rel.items()}


def _floor(lr):
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", f"noise_floor_lr{lr:g}.json")) as f:
        return json.load(f)["modes"]


def test_loss_curve_200_steps_precise_mode_at_reference_lr():
    """north_star: loss within 1 % of the reference graph over the first 200 steps -- at the reference's own learning
    rate 1e-3 (train.py:75), 8 rotating batches, clip + TF-Adam on both models, EVERY loss term at EVERY step, in the
    split-bf16 mode.  tests/noise_floor.py (CPU, committed result in tests/golden/noise_floor_lr0.001.json) shows why
    the mode exists: the float32 graph follows the float64 one to 2e-5, operands rounded to bf16 at ANY single
    product move single steps by > 10 %, tf32 operands by up to 1.8 %, split-bf16 by 3e-4."""
    rel = _curve(1e-3, 200, precise=True)
    for k, v in rel.items():
        assert v.max() < 0.01, (k, v.max(), int(v.argmax()))
    # and an order of magnitude inside the budget for the terms that are not differences of nearly equal numbers
    assert rel["teacher_loss"].max() < 2e-3 and rel["l_ce"].max() < 2e-3, (rel["teacher_loss"].max(), rel["l_ce"].max())


def test_loss_curve_200_steps_default_mode():
    """The default mode (plain bf16 operands, f32 accumulation -- what north_star's 1e-3 prediction tolerance names).
    At lr 1e-4 every loss term stays within 1 % at every one of 200 steps.  At lr 1e-3 the deviation is the
    operand rounding amplified by 200 optimizer steps, and it must be no larger than what the CPU oracle shows when
    ITS matmul operands are rounded to bf16 (tests/noise_floor.py mode 'bf16'): the GPU path has no error source
    beyond that rounding."""
    rel = _curve(1e-4, 200)
    floor4 = _floor(1e-4)["bf16"]["summary"]
    for k, v in rel.items():
        if k == "l_pred":
            # a difference of nearly equal logs (2e-4 at the start): the bf16-emulated oracle itself reaches 0.63 %
            # here, so this term is held to the emulation's level like the lr 1e-3 curve below, not to a flat 1 %
            assert v.max() <= 3.0 * floor4[k]["max"] + 2e-3, (k, v.max(), floor4[k]["max"])
            assert np.percentile(v, 95) < 0.01, (k, float(np.percentile(v, 95)))
        else:
            assert v.max() < 0.01, (k, v.max(), int(v.argmax()))
    rel = _curve(1e-3, 200)
    floor = _floor(1e-3)["bf16"]["summary"]
    for k, v in rel.items():
        # same error class as the emulation, term by term.  Both curves are single samples of an amplifying
        # process (the GPU's float atomics make even two GPU runs differ by this much late in the curve: measured
        # p95 ratios between 0.9 and 2.2 and median ratios between 1.1 and 1.8 over several runs,
        # profiles/r02_curve_stats.txt), hence factors rather than equality.
        assert np.percentile(v, 95) <= 3.0 * floor[k]["p95"] + 2e-3, (k, float(np.percentile(v, 95)), floor[k]["p95"])
        assert v.max() <= 4.0 * floor[k]["max"] + 5e-3, (k, v.max(), floor[k]["max"])
        assert np.median(v) <= 3.0 * floor[k]["median"] + 2e-3, (k, float(np.median(v)), floor[k]["median"])


@pytest.mark.parametrize("B", [1, 3])
def test_ragged_tiny_batches_and_empty_video(B):
    """Edge cases: batch smaller than any tile (rows 20*B / 5*B < 128), a video with num_frames = 0 (all
    lengths zero -> zero state -> p = sum_m softmax(0)[m]*sigmoid(b)), one with a single frame."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    cfg = ModelConfig(**SMALL)
    x, nf, lab = O.synthetic_batch(B, seed=77, num_features=cfg.feature_size, vocab_size=cfg.vocab_size)
    nf[0] = 0
    x[0] = 0.0
    tr = TeacherStudentTrainer(cfg, batch_size=B, lstm_gain=2.0)
    T = O.init_params("model", 0, dtype=torch.float64, gain=2.0, **SMALL)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=2.0, **SMALL)
    tr.forward_backward(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(),
                        torch.from_numpy(lab).cuda().view(torch.uint8))
    torch.cuda.synchronize()
    ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, None, None,
                                       clip_gradient_norm=0.0, regularization_penalty=0.0,
                                       vocab_size=cfg.vocab_size, num_mixtures=cfg.num_mixtures)
    assert torch.all(tr.t_eng.state[0] == 0) and torch.all(tr.s_eng.state[0] == 0)
    assert (tr.t_eng.pred.cpu().double() - ref["teacher_predictions"]).abs().max().item() < 1e-3
    assert (tr.s_eng.pred.cpu().double() - ref["student_predictions"]).abs().max().item() < 1e-3
    for params, grads in [(tr.teacher, ref["teacher_grads"]), (tr.student, ref["student_grads"])]:
        for n in params.names:
            g, r = params.g[n].cpu().double(), grads[n]
            den = r.norm().item()
            if den < 1e-12:
                assert g.norm().item() < 1e-6, n
            else:
                assert ((g - r).norm() / den).item() < 3e-2, n


def test_tfrecord_to_gpu_step(tmp_path):
    """End of the input path: TFRecord shard -> TF-free reader (uint8 batches) -> training step on the GPU, equal to
    feeding the dequantised float32 batch the reference's reader would produce."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import readers
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    rng = np.random.default_rng(3)
    frames = [300, 120, 7, 45]
    recs = []
    for i, n in enumerate(frames):
        f = {"rgb": rng.integers(0, 256, size=(n, 96), dtype=np.uint8),
             "audio": rng.integers(0, 256, size=(n, 32), dtype=np.uint8)}
        recs.append(readers.make_sequence_example(f"v{i}", sorted(rng.choice(200, 3, replace=False).tolist()), f))
    path = str(tmp_path / "shard.tfrecord")
    readers.write_tfrecord(path, recs)
    rd = readers.YT8MFrameFeatureReader(num_classes=200, feature_sizes=[96, 32], feature_names=["rgb", "audio"])
    ids, xq, y, nf = next(rd.batches([path], 4))
    cfg = ModelConfig(**SMALL)
    a = TeacherStudentTrainer(cfg, batch_size=4)
    b = TeacherStudentTrainer(cfg, batch_size=4)
    a.step(xq.cuda(), nf.cuda(), y.cuda())                                  # quantised path
    xf = np.where(np.arange(300)[None, :, None] < nf.numpy()[:, None, None], O.dequantize(xq.numpy()), 0.0)
    b.step(torch.from_numpy(xf.astype(np.float32)).cuda(), nf.cuda(), y.cuda())   # reference-style float batch
    fa, fb = a.fetch(), b.fetch()
    for k in fa:       # (the regulariser's sum of squares is accumulated with float atomics: order-dependent last bits)
        assert abs(fa[k] - fb[k]) <= 1e-6 * abs(fb[k]), k
    assert torch.equal(a.s_eng.pred, b.s_eng.pred) and torch.equal(a.t_eng.pred, b.t_eng.pred)


@pytest.mark.parametrize("mode", [7, 31])
def test_stream_schedules_equal_the_single_stream_step(mode, monkeypatch):
    """EVC_OVERLAP only changes WHERE kernels run (streams joined by events), never what they compute.
    B=64 puts RNN_L1 of the teacher (1280 rows) on the fused-epilogue path, so the step-by-step
    interleaving of the two cells (evc_lstm_seq_fwd_steps) is exercised.

    (a) From identical weights, one forward+backward: states and predictions are bit-identical (no atomics
        upstream of them), gradients equal to the run-to-run noise of the single-stream step (two
        reductions use float atomics whose order varies: the bias column sums and the split-K d(state) of
        the classifier).
    (b) After three training steps the weights agree except for the few elements whose gradient is within
        that noise of zero: Adam normalises the gradient, so such an element may move by ~lr per step in
        either direction in ANY two runs, also of the single-stream step (measured beside it as the
        baseline).  A misordered kernel would instead change a large fraction of a tensor."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import StudentFinetuneTrainer, TeacherStudentTrainer
    cfg = ModelConfig(**SMALL)
    B, lr = 64, 1e-4
    x, nf, lab = O.synthetic_batch(B, seed=21, num_features=cfg.feature_size, vocab_size=cfg.vocab_size)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()

    def make(overlap, cls):
        monkeypatch.setenv("EVC_OVERLAP", str(overlap))
        return cls(cfg, batch_size=B, device="cuda", base_learning_rate=lr)

    def models(tr):
        return ([tr.teacher] if hasattr(tr, "teacher") else []) + [tr.student]

    # ---- (a) one forward + backward
    a, a2, b = make(0, TeacherStudentTrainer), make(0, TeacherStudentTrainer), make(mode, TeacherStudentTrainer)
    assert a.student_stream is None and b.student_stream is not None
    for tr in (a, a2, b):
        tr.forward_backward(xd, nfd, labd.view(torch.uint8))
    torch.cuda.synchronize()
    for ea, eb in ((a.t_eng, b.t_eng), (a.s_eng, b.s_eng)):
        assert torch.equal(ea.state, eb.state) and torch.equal(ea.pred, eb.pred)
    assert torch.equal(a.rows, b.rows)
    for ma, ma2, mb in zip(models(a), models(a2), models(b)):
        for n in ma.names:
            ga, ga2, gb = ma.g[n], ma2.g[n], mb.g[n]
            # measured: up to 2e-4 of the largest element between any two runs (the atomics' 1e-7 flips bf16
            # roundings of the gate gradients, which the weight-gradient sums then average)
            noise = (ga - ga2).abs().max().item()
            assert (ga - gb).abs().max().item() <= 5 * noise + 2e-3 * ga.abs().max().item() + 1e-12, n
    del a, a2, b

    # ---- (b) three training steps
    def run(overlap, cls):
        tr = make(overlap, cls)
        for _ in range(3):
            tr.step(xd, nfd, labd)
        out = tr.fetch()
        torch.cuda.synchronize()
        return tr, out

    for cls in (TeacherStudentTrainer, StudentFinetuneTrainer):
        a, fa = run(0, cls)
        a2, fa2 = run(0, cls)               # the single-stream step's own run-to-run noise
        b, fb = run(mode, cls)
        for k in fa:
            noise = abs(fa[k] - fa2[k])
            assert abs(fa[k] - fb[k]) <= 20 * noise + 1e-3 * abs(fa[k]) + 1e-5, (k, fa[k], fa2[k], fb[k])
        noise = (a.s_eng.pred - a2.s_eng.pred).abs().max().item()
        assert (a.s_eng.pred - b.s_eng.pred).abs().max().item() <= 20 * noise + 5e-4
        for ma, ma2, mb in zip(models(a), models(a2), models(b)):
            for n in ma.names:
                d = (ma.w[n] - mb.w[n]).abs()
                assert d.max().item() <= 3 * 2 * 3.2 * lr * 1.05, (n, d.max().item())   # Adam's largest step, both ways
                frac = (d > 1e-6).float().mean().item()
                frac0 = ((ma.w[n] - ma2.w[n]).abs() > 1e-6).float().mean().item()
                assert frac <= 20 * frac0 + 0.02, (n, frac, frac0)


def test_norms_from_the_producers_equal_the_sumsq_pass(monkeypatch):
    """The default single-GPU step takes the clipped gradient norms in the weight-gradient GEMM epilogues (+ <g, w> from
    the logits, sum w^2 from the previous clip+Adam launch); EVC_FUSED_NORMS=0 runs the sumsq pass over g and w
    instead.  (a) on the same gradients the assembled norm equals sum (g + wd w)^2 of every matrix; (b) the
    regulariser bookkeeping (sum w^2 before / after the update, also after weights were set from outside);
    (c) both modes train alike (a regularisation penalty large enough for the wd terms to matter)."""
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer, _as_u8
    from oracle import hlstm_oracle as O
    cfg = ModelConfig(**SMALL)
    B, penalty = 16, 3e6
    wd = penalty * cfg.l2_penalty
    x, nf, lab = O.synthetic_batch(B, seed=11, num_features=cfg.feature_size, vocab_size=cfg.vocab_size, stress=True)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()

    def make(fused):
        monkeypatch.setenv("EVC_FUSED_NORMS", "1" if fused else "0")
        tr = TeacherStudentTrainer(cfg, batch_size=B, lstm_gain=2.0, regularization_penalty=penalty,
                                   clip_gradient_norm=0.05)
        assert tr.teacher.fused_norms() == fused
        return tr

    # (a), (b)
    tr = make(True)
    for rnd in range(2):
        if rnd == 1:             # weights set from outside: sum w^2 has to be taken again
            for p in (tr.teacher, tr.student):
                p.w[p.gates_w].mul_(1.5)
                p.refresh_shadows()
                assert not p._wsq_valid
        tr.forward_backward(xd, nfd, _as_u8(labd))
        torch.cuda.synchronize()
        before = {}
        for p in (tr.teacher, tr.student):
            assert p._fused_ready == {0, 2, 4, 6, 8, 9}
            aux = p.norm_aux.cpu().double()
            for i in sorted(p._fused_ready):
                n = p.names[i]
                g, w = p.g[n].double(), p.w[n].double()
                reg = n in (p.gates_w, p.experts_w)
                want = ((g + wd * w) ** 2).sum().item() if reg else (g ** 2).sum().item()
                got = aux[i, 0].item()
                if reg:
                    assert abs(aux[i, 1].item() - (g * w).sum().item()) <= 2e-3 * (g.norm() * w.norm()).item(), n
                    got += 2 * wd * aux[i, 1].item() + wd * wd * (w ** 2).sum().item()
                assert abs(got - want) <= 2e-5 * want, (n, got, want)
            before[p.scope] = [(p.w[n].double() ** 2).sum().item() for n in (p.gates_w, p.experts_w)]
        tr.apply_gradients()
        torch.cuda.synchronize()
        for p in (tr.teacher, tr.student):
            assert p._wsq_valid and not p._fused_ready
            for k, n in enumerate((p.gates_w, p.experts_w)):
                i = p.names.index(n)
                assert abs(p.wsq[i].item() - before[p.scope][k]) <= 1e-5 * before[p.scope][k]
                after = (p.w[n].double() ** 2).sum().item()
                assert abs(p.wsq_next[i].item() - after) <= 1e-5 * after

    # (c)
    def run(fused):
        tr = make(fused)
        out = []
        for it in range(4):
            tr.step(xd, nfd, labd)
            out.append(tr.fetch())
        torch.cuda.synchronize()
        return tr, out

    a, la = run(False)
    b, lb = run(True)
    for it in range(4):
        for k in ("teacher_loss", "student_loss", "teacher_reg", "student_reg"):
            assert abs(la[it][k] - lb[it][k]) <= 1e-4 * abs(la[it][k]) + 1e-7, (it, k, la[it][k], lb[it][k])
    assert la[0]["teacher_reg"] > 0
    for pa, pb in ((a.teacher, b.teacher), (a.student, b.student)):
        for n in pa.names:
            # Adam's update is nearly invariant to the clip scale (and flips with the sign of near-zero gradient
            # elements, which the atomically accumulated split-K sums perturb from run to run); its first moment is
            # linear in the scale (part (a) pins the norms themselves to 2e-5; this is the end-to-end sanity check,
            # with room for the run-to-run noise of four chaotic steps)
            dm = (pa.m[n] - pb.m[n]).norm().item()
            assert dm <= 1e-2 * pa.m[n].norm().item() + 1e-12, (n, dm, pa.m[n].norm().item())
