"""Parity at the shapes bench.py runs (VERDICT r1 "next" #1): the large-GEMM code paths (N-fastest tile order,
multi-round split-K, all operand majors), the full model dimensions (1152-d, 1024 cells, 4716 classes, 2 mixtures)
forward + backward + one clip/Adam step against the float64 oracle, and a cfg #1-sized forward (B = 256) with
edge-case lengths.  Tolerances as everywhere: predictions 1e-3 abs, gradients 3 % (relative l2), losses 1 %."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FULL = dict(feature_size=1152, lstm_cells=1024, vocab_size=4716, num_mixtures=2)


def _bf16_operand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _ref_matmul(a, b):
    """f64 product of the bf16-rounded operands, in chunks (the oracle for a bf16-operand / f32-accumulate GEMM)."""
    out = torch.empty(a.shape[0], b.shape[1], dtype=torch.float64, device="cuda")
    step = 4096
    for i in range(0, a.shape[0], step):
        out[i:i + step] = a[i:i + step].double() @ b.double()
    return out


# (M, N, K): 40960 x 512 x 1024 has an 80 MB A operand -> N-fastest tile order (evc_gemm.cu: n_fastest) and 640
# tiles = 5 rounds over the SMs; 1024 x 512 x 40960 is the weight-gradient shape (K = rows*steps) where split-K
# with several rounds is requested explicitly.  Same paths as the step's dX (76800x1024x4096) and wgrad
# (1024x4096x76800) GEMMs at a size the f64 check finishes in seconds.
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split_k,c_bf16", [
    (40960, 512, 1024, False, False, 1, False),     # dX: A K-major [M,K], B K-major (stored [N,K])
    (40960, 512, 1024, False, True, 1, True),       # forward projection: B MN-major (TF weight layout), bf16 out
    (1024, 512, 40960, True, True, 1, False),       # wgrad: both operands MN-major over the activation buffers
    (1024, 512, 40960, True, True, 8, False),       # ... with split-K atomics
    (40960, 384, 1024, False, False, 2, False),     # N tail (384 = 256 + 128) + split-K 2, 3 rounds
])
def test_gemm_large_paths(M, N, K, a_mn, b_mn, split_k, c_bf16):
    from efficientvideoclassification_youtube8m_b200 import ops
    A = _bf16_operand((M, K), 1, 0.5)
    Bm = _bf16_operand((K, N), 2, 0.5)
    a_st = A.t().contiguous() if a_mn else A                     # stored [K][M] when MN-major
    b_st = Bm if b_mn else Bm.t().contiguous()                   # stored [N][K] when K-major
    out = torch.zeros(M, N, dtype=torch.bfloat16 if c_bf16 else torch.float32, device="cuda")
    ops.gemm(a_st, b_st, M, N, K, out, a_mn=a_mn, b_mn=b_mn, split_k=split_k)
    torch.cuda.synchronize()
    ref = _ref_matmul(A, Bm)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # f32 accumulation of K products in the tensor core (measured 5e-5 of the largest element at K = 40960 without
    # split-K, 1e-5 with 8 splits); bf16 output adds 2^-9 relative.  A mis-addressed tile is an O(1) error.
    tol = (4e-3 if c_bf16 else 1e-4) * scale + 1e-4
    assert err <= tol, (err, scale)


def _grad_report(params, grads, tol=3e-2):
    bad = []
    for n in params.names:
        g, r = params.g[n].cpu().double(), grads[n]
        den = r.norm().item()
        rel = ((g - r).norm() / den).item() if den > 1e-12 else g.norm().item()
        if rel >= tol:
            bad.append((n, rel, den))
    return bad


def test_full_dimension_forward_backward_adam_step():
    """Full model dimensions at B = 64: R1 = 1280 rows puts the teacher's RNN_L1 on the clustered fused-epilogue
    path with K = 2176, RNN_L2 / the student on split-K slabs at H = 1024, the MoE on 4716 x (2M+1) columns.
    One joint training step (clip 1.0 + TF-Adam) against the float64 oracle: predictions 1e-3, states 2e-2, the
    five losses 1 %, all 22 gradients 3 %, and the updated weights."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B, gain = 64, 2.0
    cfg = ModelConfig(**FULL)
    x, nf, lab = O.synthetic_batch(B, seed=11, stress=True)
    tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", lstm_gain=gain)
    T = O.init_params("model", 0, dtype=torch.float64, gain=gain)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=gain)
    w0 = {n: tr.student.w[n].clone() for n in (tr.student.names[0], tr.student.names[8])}
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    # gradients before clipping / regularisation: forward_backward leaves them in params.g
    tr.forward_backward(xd, nfd, labd.view(torch.uint8))
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, None, None,
                                       clip_gradient_norm=0.0, regularization_penalty=0.0)
    assert np.array_equal(tr.nf_student.cpu().numpy(), ref["num_frames_student"])
    for name, mine, theirs, tol in [("teacher_state", tr.t_eng.state, ref["teacher_state"], 2e-2),
                                    ("student_state", tr.s_eng.state, ref["student_state"], 2e-2),
                                    ("teacher_pred", tr.t_eng.pred, ref["teacher_predictions"], 1e-3),
                                    ("student_pred", tr.s_eng.pred, ref["student_predictions"], 1e-3)]:
        err = (mine.cpu().double() - theirs).abs().max().item()
        assert err < tol, (name, err)
    v = tr.losses.cpu().tolist()
    for got, key in zip(v[:4], ["teacher_ce", "l_ce", "l_pred", "l_rep"]):
        want = float(ref[key])
        assert abs(got - want) <= 0.01 * abs(want) + 1e-5, (key, got, want)
    bad = _grad_report(tr.teacher, ref["teacher_grads"]) + _grad_report(tr.student, ref["student_grads"])
    assert not bad, bad
    # the optimizer step from the same state: per-variable clip + first TF-Adam step moves a weight by
    # lr_t * m / (sqrt(v) + eps) = lr * sign(g) (1 - O(eps/|g|)); compare where the oracle's clipped gradient is
    # not noise, on one LSTM kernel and one classifier matrix of the student
    tr.apply_gradients()
    torch.cuda.synchronize()
    lr = 1e-3
    for n in w0:
        g = O.clip_by_norm(ref["student_grads"][n] + (2.0 * 1e-8 * S[n] if "classifier" in n else 0.0), 1.0)
        want = S[n] - lr * g / (g.abs() + 1e-8 / (1 - 0.999) ** 0.5)
        got = tr.student.w[n].cpu().double()
        big = g.abs() > 0.05 * g.abs().max()
        assert big.float().mean().item() > 1e-4, n
        err = (got - want)[big].abs().max().item()
        assert err < 5e-5, (n, err)                       # 5 % of one Adam step (lr = 1e-3), f32 weights
        moved = (tr.student.w[n] - w0[n]).abs().max().item()
        assert 0.5 * lr < moved < 1.5 * lr, (n, moved)


def test_forward_cfg1_batch_256_edge_lengths():
    """cfg #1 forward at the bench's batch size (B = 256: teacher R1 = 5120 rows, 40 M-tiles in clustered pairs,
    RNN_L2 at 256 rows; student R1 = 1280) with the edge-case lengths forced into the first rows."""
    import os
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import ops
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B, gain = 256, 2.0
    cfg = ModelConfig(**FULL)
    x, nf, lab = O.synthetic_batch(B, seed=12, stress=True)
    nf[13], nf[14] = 0, 2                                    # an empty video and one shorter than every chunk
    x[np.arange(300)[None, :] >= nf[:, None]] = 0.0
    tr = TeacherStudentTrainer(cfg, batch_size=B, device="cuda", lstm_gain=gain)
    xd, nfd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda()
    tr.t_eng.forward(xd, None, True, nfd)
    ops.num_frames_student(nfd, 10, 300, tr.nf_student)
    tr.s_eng.forward(xd, tr.frame_idx, True, tr.nf_student)
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    T = O.init_params("model", 0, dtype=torch.float64, gain=gain)
    S = O.init_params("model_student", 1, dtype=torch.float64, gain=gain)
    xn = O.l2_normalize(torch.from_numpy(x).double())
    with torch.no_grad():
        ts, tp = O.teacher_forward(xn, nf, T)
        nfs = O.num_frames_student(nf, 10)
        ss, sp = O.student_forward(O.sample_uniform(xn, 10), nfs, S)
    assert np.array_equal(tr.nf_student.cpu().numpy(), nfs)
    assert torch.all(tr.t_eng.state[13] == 0) and torch.all(tr.s_eng.state[13] == 0)
    for name, mine, theirs, tol in [("teacher_pred", tr.t_eng.pred, tp, 1e-3), ("student_pred", tr.s_eng.pred, sp, 1e-3),
                                    ("teacher_state", tr.t_eng.state, ts, 2e-2), ("student_state", tr.s_eng.state, ss, 2e-2)]:
        err = (mine.cpu().double() - theirs).abs().max().item()
        assert err < tol, (name, err)
