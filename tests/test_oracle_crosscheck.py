"""The TF-1.x arithmetic the oracle restates cannot be pinned against TensorFlow here
(not installable offline).  These tests pin each restated op against an INDEPENDENT implementation of the
same published algorithm that does exist in this image -- PyTorch's own LSTMCell / LSTM (packed sequences),
Categorical + kl_divergence, normalize, clip_grad_norm_, Adam and binary_cross_entropy -- after mapping
TF's conventions (gate order i,j,f,o; forget_bias added at use; kernel rows [x;h]; epsilon placement) onto
PyTorch's.  They do not replace a TensorFlow run; they rule out slips in the restatement itself."""
import math

import numpy as np
import torch

from oracle import hlstm_oracle as O

D64 = torch.float64


def _tf_to_torch_lstm(kernel, bias, in_dim, forget_bias=1.0):
    """TF BasicLSTMCell (kernel [in+H, 4H], columns i|j|f|o) -> torch LSTMCell (rows i|f|g|o)."""
    H = bias.shape[0] // 4
    i, j, f, o = torch.chunk(kernel, 4, dim=1)
    w = torch.cat([i, f, j, o], dim=1)                      # torch order: input, forget, cell(g), output
    bi, bj, bf, bo = torch.chunk(bias, 4)
    b = torch.cat([bi, bf + forget_bias, bj, bo])
    return w[:in_dim].t().contiguous(), w[in_dim:].t().contiguous(), b, H


def test_basic_lstm_cell_equals_torch_lstmcell():
    g = torch.Generator().manual_seed(0)
    R, Din, H = 7, 12, 16
    kernel = torch.randn(Din + H, 4 * H, generator=g, dtype=D64) * 0.4
    bias = torch.randn(4 * H, generator=g, dtype=D64) * 0.3
    x, c, h = (torch.randn(R, n, generator=g, dtype=D64) for n in (Din, H, H))
    w_ih, w_hh, b, _ = _tf_to_torch_lstm(kernel, bias, Din)
    cell = torch.nn.LSTMCell(Din, H, dtype=D64)
    with torch.no_grad():
        cell.weight_ih.copy_(w_ih); cell.weight_hh.copy_(w_hh)
        cell.bias_ih.copy_(b); cell.bias_hh.zero_()
        h_ref, c_ref = cell(x, (h, c))
    c_new, h_new = O.basic_lstm_cell(x, c, h, kernel, bias)
    assert torch.allclose(c_new, c_ref, atol=1e-13) and torch.allclose(h_new, h_ref, atol=1e-13)


def test_two_layer_dynamic_rnn_equals_torch_lstm_on_packed_sequences():
    """dynamic_rnn(MultiRNNCell, sequence_length): the final state of a row is its state after its last
    valid step -- what torch.nn.LSTM returns as (h_n, c_n) for a packed sequence."""
    g = torch.Generator().manual_seed(1)
    R, T, Din, H = 9, 6, 10, 8
    cells = [(torch.randn(Din + H, 4 * H, generator=g, dtype=D64) * 0.4, torch.randn(4 * H, generator=g, dtype=D64) * 0.2),
             (torch.randn(2 * H, 4 * H, generator=g, dtype=D64) * 0.4, torch.randn(4 * H, generator=g, dtype=D64) * 0.2)]
    x = torch.randn(R, T, Din, generator=g, dtype=D64)
    lens = torch.tensor([6, 1, 3, 6, 2, 5, 4, 1, 6])
    lstm = torch.nn.LSTM(Din, H, num_layers=2, batch_first=True, dtype=D64)
    with torch.no_grad():
        for layer, ((kernel, bias), in_dim) in enumerate(zip(cells, (Din, H))):
            w_ih, w_hh, b, _ = _tf_to_torch_lstm(kernel, bias, in_dim)
            getattr(lstm, f"weight_ih_l{layer}").copy_(w_ih)
            getattr(lstm, f"weight_hh_l{layer}").copy_(w_hh)
            getattr(lstm, f"bias_ih_l{layer}").copy_(b)
            getattr(lstm, f"bias_hh_l{layer}").zero_()
        packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
        _, (h_n, c_n) = lstm(packed)
    state = O.multi_lstm_dynamic(x, lens, cells)                       # [c0 | h0 | c1 | h1]
    want = torch.cat([c_n[0], h_n[0], c_n[1], h_n[1]], dim=1)
    assert torch.allclose(state, want, atol=1e-13)
    # a row of length 0 keeps the zero state (rows of empty chunks in the lower level)
    z = O.multi_lstm_dynamic(x[:2], torch.tensor([0, 2]), cells)
    assert torch.count_nonzero(z[0]) == 0 and torch.count_nonzero(z[1]) > 0


def test_pred_kl_equals_torch_categorical_kl():
    g = torch.Generator().manual_seed(2)
    pt = torch.rand(5, 30, generator=g, dtype=D64) * 0.9 + 0.01
    ps = torch.rand(5, 30, generator=g, dtype=D64) * 0.9 + 0.01
    want = torch.distributions.kl_divergence(torch.distributions.Categorical(probs=pt),
                                             torch.distributions.Categorical(probs=ps)).sum()
    assert abs(O.pred_kl_loss(pt, ps).item() - want.item()) < 1e-12


def test_l2_normalize_equals_torch_normalize_and_keeps_zero_frames():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 4, 9, generator=g, dtype=D64)
    x[1, 2] = 0.0
    got = O.l2_normalize(x)
    assert torch.allclose(got, torch.nn.functional.normalize(x, dim=-1, eps=1e-12), atol=1e-15)
    assert torch.count_nonzero(got[1, 2]) == 0


def test_cross_entropy_equals_torch_bce_with_the_reference_epsilon():
    g = torch.Generator().manual_seed(4)
    p = torch.rand(6, 25, generator=g, dtype=D64)
    y = torch.rand(6, 25, generator=g) < 0.2
    eps = 1e-5
    # BCE(p, y) with log(p + eps), log(1 - p + eps): evaluate torch's BCE on shifted probabilities per term
    pos = torch.nn.functional.binary_cross_entropy(p + eps, torch.ones_like(p), reduction="none")
    neg = torch.nn.functional.binary_cross_entropy(p - eps, torch.zeros_like(p), reduction="none")
    want = torch.where(y, pos, neg).sum(1).mean()
    assert abs(O.cross_entropy_loss(p, y).item() - want.item()) < 1e-10


def test_clip_by_norm_equals_torch_clip_grad_norm_per_variable():
    g = torch.Generator().manual_seed(5)
    for scale in (0.01, 30.0):                               # below and above the clip norm
        w = torch.nn.Parameter(torch.zeros(11, 7, dtype=D64))
        w.grad = torch.randn(11, 7, generator=g, dtype=D64) * scale
        want_in = w.grad.clone()
        torch.nn.utils.clip_grad_norm_([w], max_norm=1.0)
        got = O.clip_by_norm(want_in, 1.0)
        assert torch.allclose(got, w.grad, rtol=1e-6, atol=0)      # torch divides by (norm + 1e-6)
    assert torch.count_nonzero(O.clip_by_norm(torch.zeros(4, dtype=D64), 1.0)) == 0


def test_tf_adam_equals_torch_adam_up_to_the_epsilon_convention():
    """TF: w -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps).  torch: w -= lr/(1-b1^t) * m/(sqrt(v/(1-b2^t))+eps),
    i.e. the same update with eps_tf = eps_torch*sqrt(1-b2^t).  With eps -> 0 they coincide."""
    g = torch.Generator().manual_seed(6)
    w0 = torch.randn(13, generator=g, dtype=D64)
    grads = [torch.randn(13, generator=g, dtype=D64) for _ in range(5)]
    p = {"w": w0.clone()}
    opt = O.TFAdam(p, lr=1e-3, eps=1e-300)
    wt = torch.nn.Parameter(w0.clone())
    topt = torch.optim.Adam([wt], lr=1e-3, betas=(0.9, 0.999), eps=1e-300)
    for gr in grads:
        opt.apply(p, {"w": gr})
        wt.grad = gr.clone()
        topt.step()
        assert torch.allclose(p["w"], wt.detach(), atol=1e-14)
    # and with the default eps the first TF step is lr * g/(|g| + eps/sqrt(1-b2)) exactly
    p = {"w": w0.clone()}
    O.TFAdam(p, lr=1e-3).apply(p, {"w": grads[0]})
    want = w0 - 1e-3 * grads[0] / (grads[0].abs() + 1e-8 / math.sqrt(1 - 0.999))
    assert torch.allclose(p["w"], want, atol=1e-15)


def test_moe_equals_a_per_class_loop():
    """video_level_models.py:437-447: reshape [-1, M+1] / [-1, M] puts class c, mixture m at column c*(M+1)+m
    / c*M+m.  Checked against an explicit loop over (video, class)."""
    g = torch.Generator().manual_seed(7)
    B, S, V, M = 3, 6, 5, 2
    params = {"s/classifier/gates/weights": torch.randn(S, V * (M + 1), generator=g, dtype=D64),
              "s/classifier/experts/weights": torch.randn(S, V * M, generator=g, dtype=D64),
              "s/classifier/experts/biases": torch.randn(V * M, generator=g, dtype=D64)}
    state = torch.randn(B, S, generator=g, dtype=D64)
    got = O.moe_predictions(state, params, "s", V, M)
    G = (state @ params["s/classifier/gates/weights"]).numpy()
    E = (state @ params["s/classifier/experts/weights"] + params["s/classifier/experts/biases"]).numpy()
    for b in range(B):
        for c in range(V):
            gl = G[b, c * (M + 1):(c + 1) * (M + 1)]
            gate = np.exp(gl - gl.max()); gate /= gate.sum()
            ex = 1.0 / (1.0 + np.exp(-E[b, c * M:(c + 1) * M]))
            assert abs(got[b, c].item() - float((gate[:M] * ex).sum())) < 1e-13


def _torch_lstm2(cells, in_dim, H):
    lstm = torch.nn.LSTM(in_dim, H, num_layers=2, batch_first=True, dtype=D64)
    with torch.no_grad():
        for layer, ((kernel, bias), d) in enumerate(zip(cells, (in_dim, H))):
            w_ih, w_hh, b, _ = _tf_to_torch_lstm(kernel, bias, d)
            getattr(lstm, f"weight_ih_l{layer}").copy_(w_ih)
            getattr(lstm, f"weight_hh_l{layer}").copy_(w_hh)
            getattr(lstm, f"bias_ih_l{layer}").copy_(b)
            getattr(lstm, f"bias_hh_l{layer}").zero_()
    return lstm


def _final_state(lstm, x, length, H):
    """[c0|h0|c1|h1] of ONE sequence after `length` steps from the zero state (zero if length == 0)."""
    if length == 0:
        return torch.zeros(4 * H, dtype=D64)
    with torch.no_grad():
        _, (h_n, c_n) = lstm(x[None, :length])
    return torch.cat([c_n[0, 0], h_n[0, 0], c_n[1, 0], h_n[1, 0]])


def test_hierarchical_state_equals_a_literal_per_video_per_chunk_loop():
    """frame_level_models.py:237-257 executed literally -- video by video, chunk by chunk: tf.split into
    num_inputs_to_lstm chunks, each chunk through RNN_L1 from the zero state for
    min(len, max(0, n - len*i)) steps, the chunk states stacked and run through RNN_L2 for ceil(n/len)
    steps -- with torch.nn.LSTM as the cell engine.  The oracle batches all chunks into one call."""
    g = torch.Generator().manual_seed(11)
    D, H, C, ell = 10, 8, 5, 4                       # 20 frames = 5 chunks x 4
    T = C * ell
    mk = lambda r, c: torch.randn(r, c, generator=g, dtype=D64) * 0.4
    params = {}
    for level, in_dim in (("RNN_L1", D), ("RNN_L2", 4 * H)):
        for cell, d in ((0, in_dim), (1, H)):
            base = f"m/{level}/rnn/multi_rnn_cell/cell_{cell}/basic_lstm_cell"
            params[base + "/kernel"] = mk(d + H, 4 * H)
            params[base + "/bias"] = mk(1, 4 * H)[0]
    num_frames = np.array([1, 3, 4, 5, 8, 9, 12, 16, 17, 19, 20], dtype=np.int32)
    x = torch.randn(len(num_frames), T, D, generator=g, dtype=D64)
    for b, n in enumerate(num_frames):
        x[b, n:] = 0.0                               # readers.py:173 zero padding
    got = O.hlstm_state(x, num_frames, params, "m", C)
    l1 = _torch_lstm2(O._cells(params, "m", "RNN_L1"), D, H)
    l2 = _torch_lstm2(O._cells(params, "m", "RNN_L2"), 4 * H, H)
    for b, n in enumerate(num_frames):
        chunks = [_final_state(l1, x[b, i * ell:(i + 1) * ell], int(min(ell, max(0, n - ell * i))), H)
                  for i in range(C)]
        len2 = int(np.ceil(np.float32(n) / np.float32(ell)))
        want = _final_state(l2, torch.stack(chunks), len2, H)
        assert torch.allclose(got[b], want, atol=1e-13), (b, n)
