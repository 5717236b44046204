"""tcgen05 GEMM family against torch (bf16-rounded operands, f32/f64 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (300, 200, 136), (512, 1024, 1152),
                                   (96, 4716 * 3, 320)])
def test_gemm_all_majors(a_mn, b_mn, M, N, K):
    from efficientvideoclassification_youtube8m_b200 import ops
    Mp, Np, Kp = ops.pad8(M, 8), ops.pad8(N, 8), ops.pad8(K, 8)
    # storage with padded pitches (TMA needs 16-byte multiples)
    A = _mk((K, Mp), 1) if a_mn else _mk((M, Kp), 1)
    B = _mk((K, Np), 2) if b_mn else _mk((N, Kp), 2)
    Al = (A[:, :M].t() if a_mn else A[:, :K]).double()
    Bl = (B[:, :N] if b_mn else B[:, :K].t()).double()
    ref = Al @ Bl
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(A, B, M, N, K, out, a_mn=a_mn, b_mn=b_mn, bias=bias)
    torch.cuda.synchronize()
    err = (out.double() - (ref + bias.double())).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_gemm_bf16_out_and_splitk():
    from efficientvideoclassification_youtube8m_b200 import ops
    M, N, K = 256, 512, 4096
    A, B = _mk((M, K), 3), _mk((N, K), 4)
    ref = A.double() @ B.double().t()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, B, M, N, K, out)
    assert (out.double() - ref).abs().max().item() < 0.02 * ref.abs().max().item()
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, acc, split_k=8)
    torch.cuda.synchronize()
    assert (acc.double() - 1 - ref).abs().max().item() < 1e-3 * ref.abs().max().item()


def _lstm_ref(x, W, b, seq_len, T, H, dtype=torch.float64):
    """dynamic_rnn(BasicLSTMCell) with autograd, x [T,rows,Kx]."""
    rows = x.shape[1]
    c = torch.zeros(rows, H, dtype=dtype, device=x.device)
    h = torch.zeros(rows, H, dtype=dtype, device=x.device)
    hs = []
    for t in range(T):
        z = torch.cat([x[t], h], 1) @ W + b
        i, j, f, o = z.chunk(4, 1)
        cn = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
        hn = torch.tanh(cn) * torch.sigmoid(o)
        live = (t < seq_len).unsqueeze(1)
        c = torch.where(live, cn, c)
        h = torch.where(live, hn, h)
        hs.append(hn)
    return c, h, torch.stack(hs)


@pytest.mark.parametrize("use_ws", [False, True, "fused"])
@pytest.mark.parametrize("rows,Kx,H,T", [(200, 128, 128, 5), (256, 1152, 1024, 3), (1300, 256, 128, 4),
                                         (5120, 1152, 1024, 2), (19000, 128, 256, 3)])
def test_lstm_seq_fwd_bwd(rows, Kx, H, T, use_ws):
    """use_ws=False: fused-epilogue kernels (forward: GEMM + cell epilogue; backward: dgrad GEMM + cell-backward
    epilogue); True: forward slab path for <=1024 rows, backward slab path (split-K / stream-K GEMM + cell kernel with
    fused bias sums); "fused": workspace present but the fused backward kernel forced (cta_group::2 pairs above 128 rows,
    last step through the stand-alone cell kernel).
    The last two shapes have more dgrad tiles than SM pairs: with a workspace the recurrent dgrad runs on the
    stream-K schedule (5120 x 1024: the teacher's RNN_L1 at B = 256; 19000 x 256: ragged last tile, one N tile)."""
    from efficientvideoclassification_youtube8m_b200 import _lib, ops
    torch.manual_seed(0)
    dev = "cuda"
    if use_ws == "fused":         # with a workspace, force the fused dgrad + cell-backward kernel (default: slab path)
        _lib.lib.evc_debug_set(4096)
    try:
        _lstm_seq_fwd_bwd_body(ops, rows, Kx, H, T, bool(use_ws), dev)
    finally:
        _lib.lib.evc_debug_set(0)


def _lstm_seq_fwd_bwd_body(ops, rows, Kx, H, T, use_ws, dev):
    ws = torch.empty(ops.lstm_workspace_bytes(rows, H, Kx), dtype=torch.uint8, device=dev) if use_ws else None
    x = (torch.randn(T, rows, Kx, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(Kx + H, 4 * H, device=dev) * (2.0 / (Kx + H) ** 0.5)).to(torch.bfloat16)
    b = torch.randn(4 * H, device=dev) * 0.1
    seq_len = torch.randint(0, T + 1, (rows,), device=dev, dtype=torch.int32)
    seq_len[:4] = torch.tensor([0, 1, T, T - 1], dtype=torch.int32)
    h_all = torch.zeros(T + 1, rows, H, dtype=torch.bfloat16, device=dev)
    c_all = torch.zeros(T + 1, rows, H, device=dev)
    gates = torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev)
    ops.lstm_seq_fwd(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws)
    torch.cuda.synchronize()

    xd = x.double().requires_grad_(True)
    Wd = W.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    c_ref, h_ref, hs_ref = _lstm_ref(xd, Wd, bd, seq_len, T, H)
    assert (c_all[T].double() - c_ref).abs().max().item() < 2e-2
    assert (h_all[T].double() - h_ref).abs().max().item() < 2e-2

    # backward: loss = sum(dc_f*c_T) + sum(dh_f*h_T) + sum(dh_ext * h_new_t)
    dc_f = torch.randn(rows, H, device=dev)
    dh_f = torch.randn(rows, H, device=dev)
    dh_ext = torch.randn(T, rows, H, device=dev) * 0.5
    live_all = (torch.arange(T, device=dev).view(T, 1) < seq_len.view(1, rows)).unsqueeze(2)
    loss = (dc_f.double() * c_ref).sum() + (dh_f.double() * h_ref).sum() + \
        (torch.where(live_all, dh_ext.double(), torch.zeros_like(dh_ext.double())) * hs_ref).sum()
    gx, gW, gb = torch.autograd.grad(loss, [xd, Wd, bd])

    dz = torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev)
    dh_pass = torch.zeros(rows, H, device=dev)
    dc = torch.zeros(rows, H, device=dev)
    dh_ext_masked = torch.where(live_all, dh_ext, torch.zeros_like(dh_ext)).contiguous()
    # the bias gradient is accumulated by the kernel that writes dz -- the cell kernel of the slab path or the fused
    # dgrad + cell-backward epilogue (pre-filled with garbage: the call zeroes it)
    db_fused = torch.full((4 * H,), 7.0, device=dev)
    ops.lstm_seq_bwd(W, Kx, rows, H, T, seq_len, gates, c_all, dh_ext_masked, dh_f, H, dc_f, H, dh_pass, dc, dz, ws,
                     dbias=db_fused)
    # wgrad: dW = [x | h_prev]^T dz ; dX = dz Wx^T ; db = colsum dz
    dW = torch.zeros(Kx + H, 4 * H, device=dev)
    R = T * rows
    ops.gemm(x.view(R, Kx), dz.view(R, 4 * H), Kx, 4 * H, R, dW[:Kx], a_mn=True, b_mn=True, ldc=4 * H)
    ops.gemm(h_all.view((T + 1) * rows, H), dz.view(R, 4 * H), H, 4 * H, R, dW[Kx:], a_mn=True, b_mn=True, ldc=4 * H)
    dX = torch.zeros(R, Kx, device=dev)
    ops.gemm(dz.view(R, 4 * H), W, R, Kx, 4 * H, dX)
    db = torch.zeros(4 * H, device=dev)
    ops.colsum_bf16(dz, R, 4 * H, 4 * H, db)
    torch.cuda.synchronize()

    def rel(a, bref):
        return ((a.double() - bref).norm() / (bref.norm() + 1e-30)).item()
    assert rel(dW, gW) < 3e-2, rel(dW, gW)
    assert rel(dX.view(T, rows, Kx), gx) < 3e-2, rel(dX.view(T, rows, Kx), gx)
    assert rel(db, gb) < 3e-2, rel(db, gb)
    # same sums of the same bf16 values, only the float summation order differs
    assert rel(db_fused, gb) < 3e-2
    assert (db_fused - db).abs().max().item() <= 1e-4 * db.abs().max().item() + 1e-6


@pytest.mark.parametrize("rows,Kx,H,T,train", [(256, 4096, 1024, 6, True), (1280, 1152, 1024, 6, True),
                                               (200, 128, 128, 5, True), (128, 256, 256, 3, False),
                                               (700, 1024, 1024, 4, True), (2048, 128, 1024, 2, False)])
def test_lstm_seq_fwd_resident_matches_per_step_path(rows, Kx, H, T, train):
    """The persistent resident-weights recurrence (evc_lstm_seq_fwd_resident: hoisted input projection + one launch
    for all T steps, Wh slices in shared memory) against the f64 reference and against the per-step split-K path:
    RNN_L2 shape (256 rows, Kx = 4H), the student's RNN_L1 (1280 rows: 5 tiles per CTA), partial tiles, one row
    group, ragged / zero lengths, with and without saved gates."""
    from efficientvideoclassification_youtube8m_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    need = ops.lstm_rec_workspace_bytes(rows, H, T)
    assert need > 0
    assert ops.lstm_rec_workspace_bytes(5120, 1024, 15) == 0          # the teacher's RNN_L1: too many rows for one wave
    raw = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    ws_rec = raw[(-raw.data_ptr()) % 1024:][:need]
    ws = torch.empty(ops.lstm_workspace_bytes(rows, H, Kx), dtype=torch.uint8, device=dev)
    x = (torch.randn(T, rows, Kx, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(Kx + H, 4 * H, device=dev) * (2.0 / (Kx + H) ** 0.5)).to(torch.bfloat16)
    b = torch.randn(4 * H, device=dev) * 0.1
    seq_len = torch.randint(0, T + 1, (rows,), device=dev, dtype=torch.int32)
    seq_len[:4] = torch.tensor([0, 1, T, T - 1], dtype=torch.int32)

    def run(fn, *extra):
        h_all = torch.zeros(T + 1, rows, H, dtype=torch.bfloat16, device=dev)
        c_all = torch.zeros(T + 1, rows, H, device=dev)
        gates = torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev) if train else None
        fn(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, *extra)
        torch.cuda.synchronize()
        return h_all, c_all, gates

    h1, c1, g1 = run(ops.lstm_seq_fwd_resident, ws_rec)
    h2, c2, g2 = run(ops.lstm_seq_fwd, ws)
    c_ref, h_ref, _ = _lstm_ref(x.double(), W.double(), b.double(), seq_len, T, H)
    assert (c1[T].double() - c_ref).abs().max().item() < 2e-2
    assert (h1[T].double() - h_ref).abs().max().item() < 2e-2
    # the two paths differ only in where the f32 partial sums are rounded (Zx + bias is formed first here)
    assert (c1 - c2).abs().max().item() < 2e-3 and (h1.float() - h2.float()).abs().max().item() < 2e-2
    if train:
        live = (torch.arange(T, device=dev).view(T, 1) < seq_len.view(1, rows)).unsqueeze(2)
        assert ((g1.float() - g2.float()) * live).abs().max().item() < 2e-2
    assert torch.all(h1[:, 0] == 0) and torch.all(c1[:, 0] == 0)        # the zero-length row never leaves the zero state
    # a second run over the same buffers (flags are reset by the call): deterministic, bit-identical
    h3, c3, _ = run(ops.lstm_seq_fwd_resident, ws_rec)
    assert torch.equal(h1, h3) and torch.equal(c1, c3)


def _split(t):
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_split_bf16_precise_mode(a_mn, b_mn):
    """evc_gemm_bf16x2: A_hi*B_hi + A_hi*B_lo + A_lo*B_hi in one f32 accumulator; ~2^-16 relative operand error
    instead of 2^-9 (measured against the f64 product of the f32 operands), all operand majors, split-K and M/N
    tails."""
    from efficientvideoclassification_youtube8m_b200 import ops
    torch.manual_seed(3)
    M, N, K = 712, 520, 2048          # M, N tails; pitches stay multiples of 8 elements (TMA)
    A = torch.randn(M, K, device="cuda")
    Bm = torch.randn(K, N, device="cuda")
    ref = A.double() @ Bm.double()
    scale = ref.abs().max().item()
    ah, al = _split(A.t().contiguous() if a_mn else A)
    bh, bl = _split(Bm if b_mn else Bm.t().contiguous())
    out = torch.zeros(M, N, device="cuda")
    ops.gemm(ah, bh, M, N, K, out, a_mn=a_mn, b_mn=b_mn, A_lo=al, B_lo=bl)
    err2 = (out.double() - ref).abs().max().item()
    plain = torch.zeros(M, N, device="cuda")
    ops.gemm(ah, bh, M, N, K, plain, a_mn=a_mn, b_mn=b_mn)
    err1 = (plain.double() - ref).abs().max().item()
    assert err2 < 2e-5 * scale, (err2, scale)            # measured ~3e-6: f32 accumulation + the dropped lo*lo term
    assert err1 > 50 * err2                               # the plain bf16 product is two orders of magnitude coarser
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(ah, bh, M, N, K, acc, a_mn=a_mn, b_mn=b_mn, split_k=4, accumulate=True, A_lo=al, B_lo=bl)
    torch.cuda.synchronize()
    assert (acc.double() - 1 - ref).abs().max().item() < 2e-5 * scale
    with pytest.raises(Exception, match="lo plane"):
        ops.gemm(ah, bh, M, N, K, out, a_mn=a_mn, b_mn=b_mn, A_lo=al)


def test_lstm_seq_precise_mode_fwd_bwd():
    """One BasicLSTM layer in split-bf16 mode (lo planes of x, W, h, gates, dz) against the f64 reference: state and
    gradient errors two orders of magnitude below the plain-bf16 tolerances, at a row count above the small-row limit
    (the precise mode routes every row count through the slab path)."""
    from efficientvideoclassification_youtube8m_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    rows, Kx, H, T = 1300, 256, 128, 4
    ws = torch.empty(ops.lstm_workspace_bytes(rows, H, Kx, True), dtype=torch.uint8, device=dev)
    xf = torch.randn(T, rows, Kx, device=dev) * 0.5
    Wf = torch.randn(Kx + H, 4 * H, device=dev) * (2.0 / (Kx + H) ** 0.5)
    b = torch.randn(4 * H, device=dev) * 0.1
    x, x_lo = _split(xf)
    W, W_lo = _split(Wf)
    seq_len = torch.randint(0, T + 1, (rows,), device=dev, dtype=torch.int32)
    seq_len[:4] = torch.tensor([0, 1, T, T - 1], dtype=torch.int32)
    h_all, h_lo = (torch.zeros(T + 1, rows, H, dtype=torch.bfloat16, device=dev) for _ in range(2))
    c_all = torch.zeros(T + 1, rows, H, device=dev)
    gates, gates_lo = (torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev) for _ in range(2))
    ops.lstm_seq_fwd(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws, x_lo=x_lo, W_lo=W_lo,
                     h_lo_all=h_lo, gates_lo_all=gates_lo)
    xd = xf.double().requires_grad_(True)
    Wd = Wf.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    c_ref, h_ref, hs_ref = _lstm_ref(xd, Wd, bd, seq_len, T, H)
    assert (c_all[T].double() - c_ref).abs().max().item() < 2e-4
    assert ((h_all[T].float() + h_lo[T].float()).double() - h_ref).abs().max().item() < 2e-4
    dc_f, dh_f = torch.randn(rows, H, device=dev), torch.randn(rows, H, device=dev)
    loss = (dc_f.double() * c_ref).sum() + (dh_f.double() * h_ref).sum()
    gx, gW, gb = torch.autograd.grad(loss, [xd, Wd, bd])
    dz, dz_lo = (torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev) for _ in range(2))
    dh_pass, dc = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev)
    db = torch.zeros(4 * H, device=dev)
    ops.lstm_seq_bwd(W, Kx, rows, H, T, seq_len, gates, c_all, None, dh_f, H, dc_f, H, dh_pass, dc, dz, ws, dbias=db,
                     W_lo=W_lo, gates_lo_all=gates_lo, dz_lo_all=dz_lo)
    R = T * rows
    dW = torch.zeros(Kx + H, 4 * H, device=dev)
    ops.gemm(x.view(R, Kx), dz.view(R, 4 * H), Kx, 4 * H, R, dW[:Kx], a_mn=True, b_mn=True, ldc=4 * H,
             A_lo=x_lo.view(R, Kx), B_lo=dz_lo.view(R, 4 * H))
    ops.gemm(h_all.view(-1, H), dz.view(R, 4 * H), H, 4 * H, R, dW[Kx:], a_mn=True, b_mn=True, ldc=4 * H,
             A_lo=h_lo.view(-1, H), B_lo=dz_lo.view(R, 4 * H))
    dX = torch.zeros(R, Kx, device=dev)
    ops.gemm(dz.view(R, 4 * H), W, R, Kx, 4 * H, dX, A_lo=dz_lo.view(R, 4 * H), B_lo=W_lo)
    torch.cuda.synchronize()

    def rel(a, bref):
        return ((a.double() - bref).norm() / (bref.norm() + 1e-30)).item()
    assert rel(dW, gW) < 3e-4, rel(dW, gW)
    assert rel(dX.view(T, rows, Kx), gx) < 3e-4, rel(dX.view(T, rows, Kx), gx)
    assert rel(db, gb) < 3e-4, rel(db, gb)


@pytest.mark.parametrize("rows,Kx,H,T,want_ks", [(100, 1024, 128, 4, 8), (256, 4096, 1024, 3, 4), (256, 1024, 1024, 3, 4),
                                                 (512, 1152, 1024, 2, 2), (130, 512, 256, 3, 4), (16, 512, 128, 5, 4)])
def test_lstm_cluster_step_matches_slab_path(rows, Kx, H, T, want_ks):
    """Small-row forward steps as one cluster split-K kernel each (partial sums through DSMEM, fused cell update;
    csrc/evc_cluster.cuh) against the f64 reference and the slab path (split-K GEMM + cell kernel): cluster sizes 8, 4
    and 2, the RNN_L2 shapes (256 rows, Kx = 4H and H), ragged row counts, zero / partial lengths, saved gates."""
    from efficientvideoclassification_youtube8m_b200 import _lib, ops
    torch.manual_seed(2)
    dev = "cuda"
    tiles = -(-rows // 128) * (H // 64)
    ks = 8
    while ks >= 2 and tiles * ks > 148:
        ks //= 2
    while ks >= 2 and Kx // 64 < 2 * ks:
        ks //= 2
    assert ks == want_ks                                  # (documents which cluster size the shape exercises)
    ws = torch.empty(ops.lstm_workspace_bytes(rows, H, Kx), dtype=torch.uint8, device=dev)
    x = (torch.randn(T, rows, Kx, device=dev) * 0.5).to(torch.bfloat16)
    W = (torch.randn(Kx + H, 4 * H, device=dev) * (2.0 / (Kx + H) ** 0.5)).to(torch.bfloat16)
    b = torch.randn(4 * H, device=dev) * 0.1
    seq_len = torch.randint(0, T + 1, (rows,), device=dev, dtype=torch.int32)
    seq_len[:4] = torch.tensor([0, 1, T, T - 1], dtype=torch.int32)

    def run(debug):
        _lib.lib.evc_debug_set(debug)
        try:
            h_all = torch.zeros(T + 1, rows, H, dtype=torch.bfloat16, device=dev)
            c_all = torch.zeros(T + 1, rows, H, device=dev)
            gates = torch.zeros(T, rows, 4 * H, dtype=torch.bfloat16, device=dev)
            n0 = _lib.launch_count()
            ops.lstm_seq_fwd(x, rows * Kx, Kx, W, b, rows, H, T, seq_len, h_all, c_all, gates, ws)
            torch.cuda.synchronize()
            return h_all, c_all, gates, _lib.launch_count() - n0
        finally:
            _lib.lib.evc_debug_set(0)

    h1, c1, g1, n1 = run(16384)          # cluster kernel (off by default: measured slower): one launch per step
    h2, c2, g2, n2 = run(8192)           # slab path: GEMM + cell kernel per step
    assert n1 == T and n2 == 2 * T
    c_ref, h_ref, _ = _lstm_ref(x.double(), W.double(), b.double(), seq_len, T, H)
    assert (c1[T].double() - c_ref).abs().max().item() < 2e-2
    assert (h1[T].double() - h_ref).abs().max().item() < 2e-2
    live = (torch.arange(T, device=dev).view(T, 1) < seq_len.view(1, rows)).unsqueeze(2)
    assert (c1 - c2).abs().max().item() < 2e-3 and (h1.float() - h2.float()).abs().max().item() < 2e-2
    assert ((g1.float() - g2.float()) * live).abs().max().item() < 2e-2
    assert torch.all(h1[:, 0] == 0) and torch.all(c1[:, 0] == 0)        # the zero-length row keeps the zero state
    h3, c3, g3, _ = run(16384)
    assert torch.equal(h1, h3) and torch.equal(c1, c3) and torch.equal(g1, g3)     # deterministic


@pytest.mark.parametrize("M,N,K,precise", [(1152, 4096, 2560, False), (512, 1412, 256, False), (300, 712, 192, True)])
def test_gemm_wgrad_sumsq_epilogue(M, N, K, precise):
    """Weight-gradient form (A and B MN-major): same C as the plain GEMM and sum(C^2) from the epilogue, accumulated
    over two launches into one slot (the x and h parts of one LSTM kernel matrix)."""
    from efficientvideoclassification_youtube8m_b200 import ops
    Mp, Np = ops.pad8(M, 8), ops.pad8(N, 8)
    A, B = _mk((K, Mp), 11), _mk((K, Np), 12)
    A_lo = (_mk((K, Mp), 13) * 2.0 ** -9) if precise else None
    B_lo = (_mk((K, Np), 14) * 2.0 ** -9) if precise else None
    ref = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, M, N, K, ref, a_mn=True, b_mn=True, A_lo=A_lo, B_lo=B_lo)
    out = torch.full((M, N), float("nan"), device="cuda")
    ss = torch.zeros(2, device="cuda")
    ops.gemm_wgrad(A, B, M, N, K, out, ss[0:1], a_mn=True, b_mn=True, A_lo=A_lo, B_lo=B_lo)
    ops.gemm_wgrad(A, B, M, N, K, out, ss[0:1], a_mn=True, b_mn=True, A_lo=A_lo, B_lo=B_lo)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    want = 2.0 * ref.double().pow(2).sum().item()
    assert abs(ss[0].item() - want) <= 2e-5 * want, (ss[0].item(), want)
    assert ss[1].item() == 0.0
    # rejected before any launch: bf16 C has no sumsq form, unaligned C
    from efficientvideoclassification_youtube8m_b200._lib import EvcError
    with pytest.raises(EvcError):
        ops.gemm_wgrad(A, B, M, N, K, torch.empty(M * N + 1, device="cuda")[1:].view(M, N), ss[0:1], a_mn=True,
                       b_mn=True)


@pytest.mark.parametrize("precise", [False, True])
def test_reg_cross_is_the_inner_product_of_weight_gradient_and_weights(precise):
    """<X^T dL, w> = sum dL * (X w): evc_reg_cross against the explicit weight gradient (f64)."""
    from efficientvideoclassification_youtube8m_b200 import ops
    B, S, N = 48, 256, 1412
    g = torch.Generator(device="cpu").manual_seed(5)
    X = torch.randn(B, S, generator=g).to(torch.bfloat16).cuda()
    W = (torch.randn(S, N, generator=g) * 0.05).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ld = ops.pad8(N, 64)
    dl = torch.zeros(B, ld, dtype=torch.bfloat16, device="cuda")
    dl[:, :N] = (torch.randn(B, N, generator=g) * 1e-2).to(torch.bfloat16).cuda()
    dl_lo = None
    if precise:
        dl_lo = torch.zeros(B, ld, dtype=torch.bfloat16, device="cuda")
        dl_lo[:, :N] = (torch.randn(B, N, generator=g) * 1e-4).to(torch.bfloat16).cuda()
    logits = (X.double() @ W.double() + bias.double()).float().contiguous()
    out = torch.zeros(1, device="cuda")
    ops.reg_cross(logits, logits.stride(0), dl, ld, bias, B, N, out, dl_lo)
    d = dl[:, :N].double() + (dl_lo[:, :N].double() if precise else 0.0)
    want = ((X.double().t() @ d) * W.double()).sum().item()
    assert abs(out.item() - want) <= 1e-4 * abs(want) + 1e-6, (out.item(), want)


def test_clip_adam_fused_norm_parts_equal_the_sumsq_pass():
    """|g + wd w|^2 assembled from sum g^2, <g, w> and sum w^2 gives the update of the sumsq pass; wsq_out receives
    the squared norm of the updated weights."""
    from efficientvideoclassification_youtube8m_b200 import ops
    n, cols = 64 * 1024, 1024
    g0 = torch.Generator(device="cpu").manual_seed(9)
    w = torch.randn(n, generator=g0).cuda()
    g = (torch.randn(n, generator=g0) * 0.05).cuda()
    wd, clip = 0.3, 1.0          # a weight decay large enough for every norm term to matter
    lr_t = torch.tensor([1e-3], device="cuda")

    def run(fused):
        ww, m, v = w.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        sh = torch.zeros(n // cols, cols, dtype=torch.bfloat16, device="cuda")
        ns, wsq_out = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
        if fused:
            parts = torch.stack([(g * g).sum(), (g * w).sum(), (w * w).sum()]).float().contiguous()
            ops.clip_adam(ww, g, m, v, ns, clip, wd, lr_t, 0.9, 0.999, 1e-8, sh, cols, cols,
                          normsq_fused=parts[0:1], reg_cross=parts[1:2], reg_wsq=parts[2:3], wsq_out=wsq_out)
        else:
            ops.sumsq(g, w, wd, ns)
            ops.clip_adam(ww, g, m, v, ns, clip, wd, lr_t, 0.9, 0.999, 1e-8, sh, cols, cols)
        torch.cuda.synchronize()
        return ww, m, v, sh, wsq_out

    a, b = run(False), run(True)
    for x, y in zip(a[:3], b[:3]):
        assert (x - y).abs().max().item() <= 1e-6 * max(1.0, x.abs().max().item())
    assert (a[3].float() - b[3].float()).abs().max().item() <= 2 ** -7 * a[3].float().abs().max().item()
    want = b[0].double().pow(2).sum().item()
    assert abs(b[4].item() - want) <= 1e-5 * want
