"""validate.py's teacher+student evaluation graph (TeacherStudentEvaluator) and the random frame samplers inside
the fused steps (BASELINE config #5), against the float64 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)


def _params(scope, seed, gain):
    from efficientvideoclassification_youtube8m_b200.params import HLstmParams, ModelConfig
    return HLstmParams(scope, ModelConfig(**SMALL), "cuda", seed, gain)


@pytest.mark.parametrize("overlap", ["0", "7"])
def test_teacher_student_evaluator_matches_validate_graph(overlap, monkeypatch):
    """validate.py:109-189: student predictions, student label loss (CrossEntropyLoss, batch mean) and the
    state-matching loss mean_b sum_j (teacher_state - student_state)^2; edge-case lengths incl. an empty video."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentEvaluator
    monkeypatch.setenv("EVC_OVERLAP", overlap)
    B, gain = 24, 2.0
    x, nf, lab = O.synthetic_batch(B, seed=19, num_features=128, vocab_size=200, stress=True)
    nf[5] = 0
    x[5] = 0.0
    ev = TeacherStudentEvaluator(_params("model", 0, gain), _params("model_student", 1, gain), B)
    assert (ev.student_stream is not None) == (overlap == "7")
    pred, idx, val, tl = ev.step(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda())
    torch.cuda.synchronize()
    T = O.init_params("model", 0, dtype=torch.float64, gain=gain, **SMALL)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=gain, **SMALL)
    with torch.no_grad():
        ref = O.teacher_student_losses(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S,
                                       vocab_size=200, num_mixtures=2)
    assert (pred.cpu().double() - ref["student_predictions"]).abs().max().item() < 1e-3
    got = ev.losses.cpu().tolist()
    assert abs(got[0] - float(ref["l_ce"])) <= 0.01 * abs(float(ref["l_ce"]))
    assert abs(got[1] - float(ref["l_rep"])) <= 0.01 * abs(float(ref["l_rep"])) + 1e-5
    rows = ((ref["teacher_state"] - ref["student_state"]) ** 2).sum(1)
    assert (ev.state_loss_rows.cpu().double() - rows).abs().max().item() <= 0.02 * rows.max().item() + 1e-5
    assert ev.state_loss_rows[5].item() == 0.0                      # empty video: both states are zero
    oi, _ = O.top_k(pred.cpu().numpy(), 20)
    assert np.array_equal(idx.cpu().numpy(), oi)                    # top-k of the student's predictions, exact
    assert np.array_equal(tl.cpu().numpy(), np.take_along_axis(lab, oi, axis=1).astype(np.uint8))


def test_validate_evaluation_loop_epoch_metrics():
    """The loop of validate.py:255-290 over host batches with a ragged last batch: epoch GAP / hit@1 / avg loss
    equal to metrics accumulated from the oracle's top-k on the same (GPU) predictions, plus the per-batch
    student_loss column."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.eval_util import EvaluationMetrics
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentEvaluator, evaluation_loop
    B = 8
    ev = TeacherStudentEvaluator(_params("model", 0, 2.0), _params("model_student", 1, 2.0), B)
    batches = []
    for i, n in enumerate([8, 8, 3]):
        x, nf, lab = O.synthetic_batch(n, seed=50 + i, num_features=128, vocab_size=200, stress=True)
        q = np.clip(np.round((x + 2.0) * 255.0 / 4.0), 0, 255).astype(np.uint8)      # uint8 features, as a reader yields
        batches.append(([f"v{i}_{j}" for j in range(n)], torch.from_numpy(q), torch.from_numpy(lab), torch.from_numpy(nf)))
    logs = []
    out = evaluation_loop(ev, batches, EvaluationMetrics(200, 20), log=logs.append)
    assert out["examples_processed"] == 19 and len(logs) == 3 and "student_loss" in logs[0]
    # recompute from the evaluator's own predictions
    m, preds, labs, sl = EvaluationMetrics(200, 20), [], [], []
    for ids, q, lab, nf in batches:
        n = q.shape[0]
        pad = lambda t: torch.cat([t, t.new_zeros((B - n,) + tuple(t.shape[1:]))]) if n < B else t  # noqa: E731
        p = ev.step(pad(q).cuda(), pad(nf).cuda(), pad(lab).cuda())[0][:n].cpu().numpy()
        sl.append((float(ev.state_loss_rows[:n].mean().item()), n))
        preds.append(p)
        labs.append(lab.numpy())
        m.accumulate(p, lab.numpy(), ev.rows[:n])
    want = m.get()
    assert abs(out["gap"] - want["gap"]) < 1e-12 and abs(out["avg_hit_at_one"] - want["avg_hit_at_one"]) < 1e-12
    assert abs(out["avg_loss"] - want["avg_loss"]) < 1e-6
    # (near-initial weights predict ~1/3 everywhere: GAP under that many exact ties depends on the tie order, which
    # the accumulators define per batch -- the device metrics are checked against the oracle without ties in
    # test_gpu_launchers.py::test_device_batch_metrics_match_host_and_golden)
    assert abs(out["avg_student_state_loss"] - sum(v * n for v, n in sl) / 19) < 1e-6


@pytest.mark.parametrize("sampling", ["random_frames", "random_sequence"])
def test_random_sampling_joint_step_matches_oracle(sampling):
    """config #5 'random': the student's frames come from SampleRandomFrames / SampleRandomSequence
    (model_utils.py:11-58) given the uniform draws u; indices bit-exact, the joint step (losses, gradients)
    within the usual tolerances of the oracle fed with the same sampled input."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B, K, gain = 16, 30, 2.0
    x, nf, lab = O.synthetic_batch(B, seed=23, num_features=128, vocab_size=200, stress=True)
    nf[3] = 0
    x[3] = 0.0
    rng = np.random.default_rng(8)
    tr = TeacherStudentTrainer(ModelConfig(**SMALL), batch_size=B, lstm_gain=gain, sampling=sampling)
    if sampling == "random_frames":
        u = rng.random((B, K), dtype=np.float32)
        idx = O.random_frame_indices(u, nf)
    else:
        u = rng.random((B,), dtype=np.float32)
        idx = np.maximum(O.random_sequence_indices(u, nf, K), 0)
    tr.forward_backward(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(),
                        torch.from_numpy(lab).cuda().view(torch.uint8), u=torch.from_numpy(u).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(tr.frame_idx_rand.cpu().numpy(), idx)
    nfs = np.where(nf > 0, K, 0).astype(np.int64)
    assert np.array_equal(tr.nf_student.cpu().numpy(), nfs)
    T = O.init_params("model", 0, dtype=torch.float64, gain=gain, **SMALL)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=gain, **SMALL)
    xn = O.l2_normalize(torch.from_numpy(x).double())
    ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, None, None,
                                       clip_gradient_norm=0.0, regularization_penalty=0.0, vocab_size=200,
                                       num_mixtures=2, student_input=O.gather_frames(xn, idx), student_num_frames=nfs)
    assert (tr.s_eng.pred.cpu().double() - ref["student_predictions"]).abs().max().item() < 1e-3
    v = tr.losses.cpu().tolist()
    for got, key in zip(v[:4], ["teacher_ce", "l_ce", "l_pred", "l_rep"]):
        want = float(ref[key])
        assert abs(got - want) <= 0.01 * abs(want) + 1e-5, (key, got, want)
    for n in tr.student.names:
        g, r = tr.student.g[n].cpu().double(), ref["student_grads"][n]
        assert ((g - r).norm() / r.norm()).item() < 3e-2, n


def test_random_sampling_own_draws_and_evaluator():
    """Without `u` the step draws from the library's Philox stream (seed, counter advancing per step): the indices
    of two consecutive steps equal the oracle's index rule on the oracle's Philox stream."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.steps import StudentEvaluator
    B, K = 12, 30
    x, nf, lab = O.synthetic_batch(B, seed=29, num_features=128, vocab_size=200)
    ev = StudentEvaluator(_params("model_student", 1, 1.0), B, sampling="random_frames", sampling_seed=77)
    xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
    per_step = (B * K + 3) // 4
    for step in range(2):
        pred = ev.step(xd, nfd)[0]
        u = O.philox_uniform(77, step * per_step, B * K).reshape(B, K)
        assert np.array_equal(ev.frame_idx_rand.cpu().numpy(), O.random_frame_indices(u, nf))
    S = O.init_params("model_student", 1, dtype=torch.float64, **SMALL)
    xn = O.l2_normalize(torch.from_numpy(x).double())
    with torch.no_grad():
        _, sp = O.student_forward(O.gather_frames(xn, O.random_frame_indices(u, nf)),
                                  np.where(nf > 0, K, 0).astype(np.int64), S, vocab_size=200, num_mixtures=2)
    assert (pred.cpu().double() - sp).abs().max().item() < 1e-3
