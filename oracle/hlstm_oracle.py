"""CPU oracle for the Hierarchical-LSTM teacher-student hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.

PARITY UNPINNED for the TensorFlow-internal arithmetic: the reference is a TF-1.x
graph (``README.md:12``), TensorFlow is not vendored under ``/root/reference`` and is
not installable here (no network), and the reference ships no tests or golden
vectors for this path.  This file restates the published TF r1.3/r1.4 semantics
of the ops the reference calls; the only anchors are the
reference's call sites and the README training log (CE at init = 1914.1,
``README.md:116``), which ``tests/test_oracle.py`` reproduces, plus cross-checks of
every restated op against PyTorch's independent implementations of the same
algorithms (``tests/test_oracle_crosscheck.py``: LSTMCell / packed LSTM, Categorical
KL, Adam, clip_grad_norm_, normalize, BCE).  The top-k / GAP
part of the path *is* pinned: ``tests/golden/make_golden_eval.py`` imports the
reference's own ``eval_util.py`` and records its outputs.

Everything is plain PyTorch on the CPU.  ``dtype=torch.float64`` is the
correctness oracle, ``torch.float32`` is the timed CPU baseline ("port").

Reference files restated (relative to /root/reference/code_student_uniform):
  train.py:253-272,281-334,349-418   step definition (normalise, sampler, losses, train ops)
  train_finetune.py:242-318          student-only step
  frame_level_models.py:200-338      HierarchicalLstmModel.create_model / create_model_inference
  video_level_models.py:397-448      MoeModel.create_model
  losses.py:90-97                    CrossEntropyLoss
  model_utils.py:11-58               SampleRandomSequence / SampleRandomFrames
  eval_util.py:17-124                hit@1, PERR, GAP, top_k_triplets
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

MAX_FRAMES = 300  # train.py:262  max_num_frames_before_sampling


# ----------------------------------------------------------------------------
# weight layout (README.md:98,105; validate.py:350-374; train_convert_model.py:501-511)
# ----------------------------------------------------------------------------
def variable_names(scope: str) -> List[str]:
    """The 11 trainable tensors of one H-LSTM model, in TF creation order."""
    names = []
    for level in ("RNN_L1", "RNN_L2"):
        for cell in (0, 1):
            base = f"{scope}/{level}/rnn/multi_rnn_cell/cell_{cell}/basic_lstm_cell"
            names += [base + "/kernel", base + "/bias"]
    names += [f"{scope}/classifier/gates/weights",
              f"{scope}/classifier/experts/weights",
              f"{scope}/classifier/experts/biases"]
    return names


def variable_shapes(scope: str, feature_size=1152, lstm_cells=1024, vocab_size=4716,
                    num_mixtures=2) -> Dict[str, Tuple[int, ...]]:
    """Shapes of the 11 tensors (SURVEY.md 8a weight-layout contract). lstm_layers == 2."""
    H, S = lstm_cells, 4 * lstm_cells
    n = variable_names(scope)
    return {
        n[0]: (feature_size + H, 4 * H), n[1]: (4 * H,),
        n[2]: (2 * H, 4 * H), n[3]: (4 * H,),
        n[4]: (S + H, 4 * H), n[5]: (4 * H,),
        n[6]: (2 * H, 4 * H), n[7]: (4 * H,),
        n[8]: (S, vocab_size * (num_mixtures + 1)),
        n[9]: (S, vocab_size * num_mixtures),
        n[10]: (vocab_size * num_mixtures,),
    }


def init_params(scope: str, seed: int, dtype=torch.float64, gain: float = 1.0,
                **shape_kw) -> Dict[str, torch.Tensor]:
    """Glorot-uniform weights, zero biases (TF defaults).

    ``gain`` > 1 is the parity "stress variant" (SURVEY 8d): it scales the LSTM
    kernels so that the gates leave the linear regime.
    """
    rng = np.random.RandomState(seed)
    out = {}
    for name, shp in variable_shapes(scope, **shape_kw).items():
        if len(shp) == 1:
            out[name] = torch.zeros(shp, dtype=dtype)
        else:
            limit = math.sqrt(6.0 / (shp[0] + shp[1]))
            w = rng.uniform(-limit, limit, size=shp)
            if "basic_lstm_cell" in name:
                w = w * gain
            out[name] = torch.from_numpy(w).to(dtype)
    return out


# ----------------------------------------------------------------------------
# A.1 / A.2  input normalisation and frame samplers
# ----------------------------------------------------------------------------
def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """train.py:256  tf.nn.l2_normalize(x, dim=2): x * rsqrt(max(sum x^2, 1e-12))."""
    ss = (x * x).sum(dim=-1, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=1e-12))


def uniform_frame_indices(every_n: int) -> List[int]:
    """train.py:265-269  while every_n*k <= 299: append(every_n*k)."""
    out, k = [], 0
    while every_n * k <= 299:
        out.append(every_n * k)
        k += 1
    return out


def num_frames_student(num_frames: np.ndarray, every_n: int) -> np.ndarray:
    """train.py:263-264  int64( (num_frames / 300) * int(300/every_n) ) evaluated in float64."""
    m = int(MAX_FRAMES / every_n)
    q = np.asarray(num_frames, dtype=np.int32).astype(np.float64) / np.float64(MAX_FRAMES)
    return np.trunc(q * np.float64(m)).astype(np.int64)


def sample_uniform(model_input: torch.Tensor, every_n: int) -> torch.Tensor:
    """train.py:270-272  transpose / gather(list_index_to_retain) / transpose."""
    idx = torch.tensor(uniform_frame_indices(every_n), dtype=torch.long)
    return model_input.index_select(1, idx)


def random_frame_indices(u: np.ndarray, num_frames: np.ndarray) -> np.ndarray:
    """model_utils.py:49-53  int32( u[b,k] * float32(num_frames[b]) ), u ~ U[0,1) float32."""
    u = np.asarray(u, dtype=np.float32)
    nf = np.asarray(num_frames).astype(np.float32).reshape(-1, 1)
    return (u * nf).astype(np.float32).astype(np.int32)  # tf.cast truncates toward zero


def random_sequence_indices(u: np.ndarray, num_frames: np.ndarray, num_samples: int) -> np.ndarray:
    """model_utils.py:23-33  start = int32(u[b] * float32(max(n-K,0)+1)); idx = min(start+k, n-1)."""
    u = np.asarray(u, dtype=np.float32).reshape(-1, 1)
    n = np.asarray(num_frames, dtype=np.int32).reshape(-1, 1)
    max_start = np.maximum(n - num_samples, 0)
    start = (u * (max_start + 1).astype(np.float32)).astype(np.float32).astype(np.int32)
    off = np.arange(num_samples, dtype=np.int32).reshape(1, -1)
    return np.minimum(start + off, (n - 1).astype(np.int32))


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; the generator behind tf.random_uniform, [TF
    random_distributions.h PhiloxRandom]) on uint32 arrays: counter [..., 4], key [2] -> [..., 4].
    Pinned by the Random123 known-answer vectors in tests/test_oracle.py."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    m0, m1, mask, s32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF), np.uint64(32)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [((p1 >> s32) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> s32) ^ c[3] ^ k1) & mask, p0 & mask]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_uniform(seed: int, offset: int, n: int) -> np.ndarray:
    """The U[0,1) float32 stream of evc_random_uniform: element i = lane i%4 of counter (offset + i//4) under
    the 64-bit key `seed`, converted like TF's Uint32ToFloat (23 mantissa bits -> [1,2) - 1)."""
    q = (n + 3) // 4
    ctr = np.uint64(offset) + np.arange(q, dtype=np.uint64)
    counter = np.stack([ctr & np.uint64(0xFFFFFFFF), ctr >> np.uint64(32), np.zeros(q, np.uint64),
                        np.zeros(q, np.uint64)], axis=-1)
    r = philox4x32_10(counter, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n]
    return ((np.uint32(127 << 23) | (r & np.uint32(0x7FFFFF))).view(np.float32) - np.float32(1.0)).astype(np.float32)


def gather_frames(model_input: torch.Tensor, index: np.ndarray) -> torch.Tensor:
    """model_utils.py:34-36,55-58  tf.gather_nd(model_input, stack([batch_index, frame_index], 2))."""
    idx = torch.from_numpy(np.asarray(index, dtype=np.int64))
    b = torch.arange(model_input.shape[0]).unsqueeze(1).expand_as(idx)
    return model_input[b, idx]


# ----------------------------------------------------------------------------
# A.3-A.5  BasicLSTMCell / MultiRNNCell(state_is_tuple=False) / dynamic_rnn
# ----------------------------------------------------------------------------
def basic_lstm_cell(x, c, h, kernel, bias, forget_bias=1.0):
    """[TF rnn_cell_impl.BasicLSTMCell] gate order i, j, f, o; kernel rows = [x ; h]."""
    z = torch.cat([x, h], dim=1) @ kernel + bias
    i, j, f, o = torch.chunk(z, 4, dim=1)
    c_new = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
    h_new = torch.tanh(c_new) * torch.sigmoid(o)
    return c_new, h_new


def multi_lstm_dynamic(x_seq, seq_len, cells):
    """dynamic_rnn(MultiRNNCell([BasicLSTMCell]*L, state_is_tuple=False), x, sequence_length).

    x_seq [R, T, Din]; seq_len int tensor [R]; cells = [(kernel, bias), ...].
    Returns the final state [R, 2*H*L] = [c0|h0|c1|h1...] (frame_level_models.py:221-235,
    247-252).  Rows with t >= seq_len keep their state (A.5).
    """
    R, T, _ = x_seq.shape
    H = cells[0][1].shape[0] // 4
    state = [(x_seq.new_zeros(R, H), x_seq.new_zeros(R, H)) for _ in cells]
    for t in range(T):
        live = (t < seq_len).unsqueeze(1)
        inp = x_seq[:, t]
        new_state = []
        for (kernel, bias), (c, h) in zip(cells, state):
            c_new, h_new = basic_lstm_cell(inp, c, h, kernel, bias)
            inp = h_new
            new_state.append((torch.where(live, c_new, c), torch.where(live, h_new, h)))
        state = new_state
    return torch.cat([t_ for ch in state for t_ in ch], dim=1)


def _cells(params, scope, level):
    out = []
    for cell in (0, 1):
        base = f"{scope}/{level}/rnn/multi_rnn_cell/cell_{cell}/basic_lstm_cell"
        out.append((params[base + "/kernel"], params[base + "/bias"]))
    return out


def hlstm_state(model_input, num_frames, params, scope, num_chunks):
    """Two-level LSTM (frame_level_models.py:237-257 teacher, :307-328 student).

    The ``num_chunks`` lower-level dynamic_rnn calls are independent and start
    from a zero state, so they are evaluated as one batch of
    ``num_chunks*B`` rows.
    """
    B, T, D = model_input.shape
    ell = T // num_chunks                                   # len_lower_lstm
    nf = torch.as_tensor(np.asarray(num_frames), dtype=torch.int64)
    chunk = torch.arange(num_chunks, dtype=torch.int64).unsqueeze(1)          # [C,1]
    len_l1 = torch.clamp(nf.unsqueeze(0) - ell * chunk, min=0, max=ell)       # [C,B]
    x = model_input.reshape(B, num_chunks, ell, D).permute(1, 0, 2, 3).reshape(num_chunks * B, ell, D)
    s1 = multi_lstm_dynamic(x, len_l1.reshape(-1), _cells(params, scope, "RNN_L1"))
    l2_in = s1.reshape(num_chunks, B, -1).permute(1, 0, 2)                    # tf.stack(axis=1)
    # ceil in float32 exactly as tf.ceil(tf.cast(n, float32) / len_lower_lstm)
    len_l2 = torch.from_numpy(
        np.ceil(nf.numpy().astype(np.float32) / np.float32(ell)).astype(np.int32).astype(np.int64))
    return multi_lstm_dynamic(l2_in, len_l2, _cells(params, scope, "RNN_L2"))


# ----------------------------------------------------------------------------
# A.7 / A.8  MoE classifier and losses
# ----------------------------------------------------------------------------
def moe_predictions(state, params, scope, vocab_size, num_mixtures):
    """video_level_models.py:421-448 (column c*(M+1)+m <-> class c, mixture m)."""
    wg = params[f"{scope}/classifier/gates/weights"]
    we = params[f"{scope}/classifier/experts/weights"]
    be = params[f"{scope}/classifier/experts/biases"]
    g = (state @ wg).reshape(-1, num_mixtures + 1)
    e = (state @ we + be).reshape(-1, num_mixtures)
    p = (torch.softmax(g, dim=1)[:, :num_mixtures] * torch.sigmoid(e)).sum(1)
    return p.reshape(-1, vocab_size)


def cross_entropy_loss(predictions, labels):
    """losses.py:90-97 (epsilon = 10e-6 = 1e-5)."""
    eps = 10e-6
    y = labels.to(predictions.dtype)
    ce = y * torch.log(predictions + eps) + (1 - y) * torch.log(1 - predictions + eps)
    return (-ce).sum(1).mean()


def regularization_loss(params, scope, l2_penalty=1e-8):
    """slim.l2_regularizer(1e-8) on gates/experts weights = scale * sum(w^2)/2
    (video_level_models.py:428,434; train.py:305-307)."""
    wg = params[f"{scope}/classifier/gates/weights"]
    we = params[f"{scope}/classifier/experts/weights"]
    return l2_penalty * (wg * wg).sum() / 2 + l2_penalty * (we * we).sum() / 2


def rep_loss(teacher_state, student_state):
    """train.py:359-362  reduce_mean_b reduce_sum_j (t - s)^2."""
    return ((teacher_state - student_state) ** 2).sum(1).mean()


def pred_kl_loss(teacher_pred, student_pred):
    """train.py:398-402  reduce_sum_b KL(Categorical(probs=pT) || Categorical(probs=pS));
    Categorical re-normalises through log_softmax(log p)."""
    a = torch.log_softmax(torch.log(teacher_pred), dim=1)
    b = torch.log_softmax(torch.log(student_pred), dim=1)
    return (torch.exp(a) * (a - b)).sum()


# ----------------------------------------------------------------------------
# forward passes of the two plugin entry points
# ----------------------------------------------------------------------------
def teacher_forward(model_input, num_frames, params, scope="model", num_inputs_to_lstm=20,
                    vocab_size=4716, num_mixtures=2):
    """HierarchicalLstmModel.create_model (frame_level_models.py:200-267)."""
    state = hlstm_state(model_input, num_frames, params, scope, num_inputs_to_lstm)
    return state, moe_predictions(state, params, scope, vocab_size, num_mixtures)


def student_forward(model_input_student, num_frames_s, params, scope="model_student",
                    num_inputs_L1=5, vocab_size=4716, num_mixtures=2):
    """HierarchicalLstmModel.create_model_inference (frame_level_models.py:269-338)."""
    state = hlstm_state(model_input_student, num_frames_s, params, scope, num_inputs_L1)
    return state, moe_predictions(state, params, scope, vocab_size, num_mixtures)


# ----------------------------------------------------------------------------
# A.9 / A.10  slim create_train_op: per-variable clip_by_norm + TF Adam
# ----------------------------------------------------------------------------
def clip_by_norm(g, clip_norm):
    """[TF clip_ops.clip_by_norm]  g * c * min(rsqrt(sum g^2), 1/c); all-zero g stays zero."""
    l2 = (g * g).sum()
    scale = clip_norm * torch.minimum(torch.rsqrt(l2), torch.tensor(1.0 / clip_norm, dtype=g.dtype))
    out = g * scale
    return torch.where(l2 > 0, out, torch.zeros_like(g))


class TFAdam:
    """[TF training_ops ApplyAdam] lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for k, g in grads.items():
            self.m[k] += (g - self.m[k]) * (1 - self.b1)
            self.v[k] += (g * g - self.v[k]) * (1 - self.b2)
            params[k] -= lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


def _grads(loss, params):
    names = list(params.keys())
    gs = torch.autograd.grad(loss, [params[n] for n in names])
    return dict(zip(names, gs))


def teacher_student_losses(model_input_raw, num_frames, labels, teacher, student, every_n=10,
                           num_inputs_to_lstm=20, num_inputs_L1=5, vocab_size=4716,
                           num_mixtures=2, regularization_penalty=2.0, student_input=None,
                           student_num_frames=None):
    """Forward half of train.py:253-406.  Returns a dict of tensors (graph attached)."""
    x = l2_normalize(model_input_raw)
    xs = sample_uniform(x, every_n) if student_input is None else student_input
    n_s = num_frames_student(num_frames, every_n) if student_num_frames is None else student_num_frames
    t_state, t_pred = teacher_forward(x, num_frames, teacher, "model", num_inputs_to_lstm,
                                      vocab_size, num_mixtures)
    ce_t = cross_entropy_loss(t_pred, labels)
    reg_t = regularization_loss(teacher, "model")
    final_t = regularization_penalty * reg_t + ce_t                         # train.py:324
    s_state, s_pred = student_forward(xs, n_s, student, "model_student", num_inputs_L1,
                                      vocab_size, num_mixtures)
    l_rep = rep_loss(t_state.detach(), s_state)                              # F9: teacher is constant
    l_ce = cross_entropy_loss(s_pred, labels)
    l_pred = pred_kl_loss(t_pred.detach(), s_pred)
    reg_s = regularization_loss(student, "model_student")
    total_s = l_rep + l_pred + l_ce + l_rep + regularization_penalty * reg_s  # train.py:406 (F7)
    return dict(teacher_state=t_state, teacher_predictions=t_pred, teacher_ce=ce_t,
                teacher_reg=reg_t, teacher_loss=final_t, student_state=s_state,
                student_predictions=s_pred, l_rep=l_rep, l_ce=l_ce, l_pred=l_pred,
                student_reg=reg_s, student_loss=total_s, num_frames_student=n_s)


def teacher_student_train_step(model_input_raw, num_frames, labels, teacher, student,
                               opt_t: Optional[TFAdam], opt_s: Optional[TFAdam],
                               clip_gradient_norm=1.0, **kw):
    """One ``sess.run([train_op, train_student_op, ...])`` of train.py:516 (both models read the
    pre-update weights).  Updates ``teacher`` / ``student`` in place; returns losses and the
    clipped gradients."""
    for p in list(teacher.values()) + list(student.values()):
        p.requires_grad_(True)
    out = teacher_student_losses(model_input_raw, num_frames, labels, teacher, student, **kw)
    g_t = _grads(out["teacher_loss"], teacher)
    g_s = _grads(out["student_loss"], student)
    for p in list(teacher.values()) + list(student.values()):
        p.requires_grad_(False)
    if clip_gradient_norm > 0:
        g_t = {k: clip_by_norm(g, clip_gradient_norm) for k, g in g_t.items()}
        g_s = {k: clip_by_norm(g, clip_gradient_norm) for k, g in g_s.items()}
    if opt_t is not None:
        opt_t.apply(teacher, g_t)
    if opt_s is not None:
        opt_s.apply(student, g_s)
    res = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    res["teacher_grads"], res["student_grads"] = g_t, g_s
    return res


def student_finetune_step(model_input_raw, num_frames, labels, student, opt_s: Optional[TFAdam],
                          every_n=10, num_inputs_L1=5, vocab_size=4716, num_mixtures=2,
                          regularization_penalty=2.0, clip_gradient_norm=1.0):
    """train_finetune.py:242-318: final_loss = penalty*reg + L_CE on the student alone."""
    for p in student.values():
        p.requires_grad_(True)
    x = l2_normalize(model_input_raw)
    xs = sample_uniform(x, every_n)
    n_s = num_frames_student(num_frames, every_n)
    s_state, s_pred = student_forward(xs, n_s, student, "model_student", num_inputs_L1,
                                      vocab_size, num_mixtures)
    l_ce = cross_entropy_loss(s_pred, labels)
    reg_s = regularization_loss(student, "model_student")
    total = regularization_penalty * reg_s + l_ce
    g_s = _grads(total, student)
    for p in student.values():
        p.requires_grad_(False)
    if clip_gradient_norm > 0:
        g_s = {k: clip_by_norm(g, clip_gradient_norm) for k, g in g_s.items()}
    if opt_s is not None:
        opt_s.apply(student, g_s)
    return dict(student_state=s_state.detach(), student_predictions=s_pred.detach(),
                l_ce=l_ce.detach(), student_reg=reg_s.detach(), student_loss=total.detach(),
                student_grads=g_s, num_frames_student=n_s)


# ----------------------------------------------------------------------------
# A.11  top-k / GAP@20 / hit@1 / PERR  (eval_util.py, average_precision_calculator.py)
# ----------------------------------------------------------------------------
def top_k(predictions: np.ndarray, k: int = 20) -> Tuple[np.ndarray, np.ndarray]:
    """eval_util.py:118-124 selects the k largest per row with numpy.argpartition (an
    unordered set; ties at the boundary are implementation-defined).  The oracle fixes
    the order: value descending, lower class index first among equal values."""
    p = np.asarray(predictions)
    k = min(k, p.shape[1])
    order = np.lexsort((np.arange(p.shape[1])[None, :].repeat(p.shape[0], 0), -p), axis=1)[:, :k]
    return order.astype(np.int32), np.take_along_axis(p, order, axis=1)


def hit_at_one(predictions: np.ndarray, actuals: np.ndarray) -> float:
    """eval_util.py:17-31."""
    top = np.argmax(predictions, 1)
    return float(np.average(actuals[np.arange(actuals.shape[0]), top]))


def perr(predictions: np.ndarray, actuals: np.ndarray) -> float:
    """eval_util.py:34-59 precision at equal recall rate (ties resolved like top_k above)."""
    agg = 0.0
    for row in range(actuals.shape[0]):
        nl = int(np.sum(actuals[row]))
        idx, val = top_k(predictions[row:row + 1], nl)
        item = sum(float(actuals[row][c]) for c, v in zip(idx[0], val[0]) if v > 0)
        # num_labels == 0: argpartition(p, -0)[-0:] is every class and no label is set -> 0
        agg += item / idx.shape[1] if nl > 0 else 0.0
    return agg / actuals.shape[0]


def ap_at_n(predictions: Sequence[float], actuals: Sequence[float], n: Optional[int],
            total_num_positives: Optional[float]) -> float:
    """average_precision_calculator.py:166-232 without the seed-0 tie shuffle (:234-240):
    stable sort descending, sum precision * delta_recall."""
    p = np.asarray(predictions, dtype=np.float64)
    a = np.asarray(actuals, dtype=np.float64)
    order = np.argsort(-p, kind="stable")
    numpos = float(np.sum(a > 0)) if total_num_positives is None else float(total_num_positives)
    if numpos == 0:
        return 0.0
    if n is not None:
        numpos = min(numpos, n)
    r = len(order) if n is None else min(len(order), n)
    ap, pos = 0.0, 0.0
    for i in range(r):
        if a[order[i]] > 0:
            pos += 1
            ap += pos / (i + 1) / numpos
    return ap


def gap(predictions: np.ndarray, actuals: np.ndarray, k: int = 20) -> float:
    """eval_util.py:61-79: AP over the pooled per-video top-k triplets, numpos = all labels."""
    idx, val = top_k(predictions, k)
    lab = np.take_along_axis(np.asarray(actuals), idx, axis=1)
    # eval_util regroups triplets by class before flattening (:107-113); replicate that order
    # so that ties between equal predictions fall the same way under the stable sort.
    cls = idx.reshape(-1)
    order = np.argsort(cls, kind="stable")
    return ap_at_n(val.reshape(-1)[order], lab.reshape(-1)[order], None, float(np.sum(actuals)))


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
EDGE_NUM_FRAMES = [1, 5, 6, 9, 10, 14, 15, 16, 29, 30, 150, 299, 300]


def dequantize(q: np.ndarray, max_quantized_value=2.0, min_quantized_value=-2.0) -> np.ndarray:
    """utils.py:9-25 Dequantize in float32: q * (range/255) + (range/512 + min)."""
    rng_ = max_quantized_value - min_quantized_value
    scalar = np.float32(rng_ / 255.0)
    bias = np.float32(rng_ / 512.0 + min_quantized_value)
    return q.astype(np.float32) * scalar + bias


def quantized_batch(batch, seed=1234, num_features=1152, max_frames=MAX_FRAMES):
    """The uint8 features behind synthetic_batch(stress=False) with the same seed (before Dequantize
    and zero padding, i.e. what a tfrecord holds)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(batch, max_frames, num_features), dtype=np.uint8)


def synthetic_batch(batch, seed=1234, num_features=1152, vocab_size=4716, full_length=False,
                    max_frames=MAX_FRAMES, stress=False):
    """Dequantised-uint8 features (utils.py:21-25 Dequantize with readers.py:178-179 defaults
    max=2,min=-2), zero-padded past num_frames (readers.py:173), ~3.4 labels/video."""
    rng = np.random.default_rng(seed)
    if stress:
        x = rng.standard_normal((batch, max_frames, num_features)).astype(np.float32)
    else:
        q = rng.integers(0, 256, size=(batch, max_frames, num_features), dtype=np.uint8)
        x = q.astype(np.float32) * np.float32(4.0 / 255.0) + np.float32(4.0 / 512.0 - 2.0)
    if full_length:
        nf = np.full((batch,), max_frames, dtype=np.int32)
    else:
        nf = rng.integers(1, max_frames + 1, size=(batch,)).astype(np.int32)
        edge = [e for e in EDGE_NUM_FRAMES if e <= max_frames][:batch]
        nf[:len(edge)] = edge
    x[np.arange(max_frames)[None, :] >= nf[:, None]] = 0.0
    lrng = np.random.default_rng(seed + 3087)
    labels = np.zeros((batch, vocab_size), dtype=bool)
    for b in range(batch):
        k = min(max(1, int(lrng.poisson(3.4))), vocab_size)
        labels[b, lrng.choice(vocab_size, size=k, replace=False)] = True
    return x, nf, labels
