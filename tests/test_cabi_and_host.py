"""CPU suite: the C-ABI library loads and exports every symbol include/evc.h declares (no compute
calls without a GPU); host-side logic of the plugin surface."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "evc.h")).read()
    return sorted(set(re.findall(r"\b(evc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "efficientvideoclassification_youtube8m_b200", "libevc.so"))
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libevc.so does not export {s}"
    lib.evc_version.restype = ctypes.c_int
    assert lib.evc_version() >= 1


def test_ctypes_table_matches_header():
    from efficientvideoclassification_youtube8m_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    src = open(os.path.join(ROOT, "include", "evc.h")).read()
    for name, argtypes in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        if name == "evc_last_error":
            n = 0
        assert n == len(argtypes), (name, n, len(argtypes))


def test_variable_names_and_shapes_follow_the_reference_checkpoint_layout():
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig, variable_names, variable_shapes
    names = variable_names("model_student")
    assert names[0] == "model_student/RNN_L1/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"
    assert names[-1] == "model_student/classifier/experts/biases"
    shp = variable_shapes("model", ModelConfig())
    assert shp["model/RNN_L1/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"] == (2176, 4096)
    assert shp["model/RNN_L2/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"] == (5120, 4096)
    assert shp["model/classifier/gates/weights"] == (4096, 14148)
    assert shp["model/classifier/experts/weights"] == (4096, 9432)
    total = sum(int(__import__("numpy").prod(s)) for s in shp.values())
    assert total == 143271128


def test_flags_and_class_lookup():
    from efficientvideoclassification_youtube8m_b200 import frame_level_models, losses, video_level_models
    from efficientvideoclassification_youtube8m_b200.flags import FLAGS
    FLAGS.reset()
    FLAGS.parse("--model HierarchicalLstmModel --batch_size 256 --num_inputs_to_lstm 20 --lstm_layers 2 "
                "--every_n 10".split())
    assert FLAGS.every_n == 10 and FLAGS.batch_size == 256
    with pytest.raises(AttributeError):
        FLAGS.parse(["--no_such_flag", "1"])

    def find_class_by_name(name, modules):                 # train.py:179-182
        modules = [getattr(module, name, None) for module in modules]
        return next(a for a in modules if a)
    assert find_class_by_name(FLAGS.model, [frame_level_models, video_level_models])().__class__.__name__ == \
        "HierarchicalLstmModel"
    assert isinstance(find_class_by_name(FLAGS.label_loss, [losses])(), losses.BaseLoss)
    assert getattr(video_level_models, FLAGS.video_level_classifier_model).__name__ == "MoeModel"
    FLAGS.reset()


def test_scope_requires_variable_scope_and_uniform_indices():
    from efficientvideoclassification_youtube8m_b200 import scope
    from efficientvideoclassification_youtube8m_b200.steps import uniform_frame_indices
    with pytest.raises(RuntimeError):
        scope.root_scope()
    with scope.variable_scope("model"):
        with scope.variable_scope("classifier") as full:
            assert full == "model/classifier" and scope.root_scope() == "model"
    assert uniform_frame_indices(10) == list(range(0, 300, 10))
    assert len(uniform_frame_indices(30)) == 10


def test_average_precision_calculator_matches_reference_golden():
    import numpy as np
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.average_precision_calculator import AveragePrecisionCalculator
    gold = np.load(os.path.join(ROOT, "tests", "golden", "eval_golden.npz"))
    for case in ("small", "yt8m"):
        p, y = gold[case + "/predictions"], gold[case + "/labels"]
        idx, val = O.top_k(p, 20)
        lab = np.take_along_axis(y, idx, axis=1)
        calc = AveragePrecisionCalculator()
        calc.accumulate(val.reshape(-1), lab.reshape(-1), float(y.sum()))
        assert abs(calc.peek_ap_at_n() - float(gold[case + "/gap"])) < 1e-6
    with pytest.raises(ValueError):
        AveragePrecisionCalculator(top_n=-1)


def test_missing_extension_fails_loudly(tmp_path):
    """No CPU fallback: without libevc.so the binding refuses to load, and host tensors are rejected."""
    import torch
    from efficientvideoclassification_youtube8m_b200 import _lib, ops
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib._load(str(tmp_path / "libevc.so"))
    with pytest.raises(ValueError):
        ops.topk(torch.zeros(2, 30), 5)            # CPU tensor


def test_bench_gpu_arm_does_not_import_the_oracle():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_ours"):src.index("def main")]
    assert "oracle" not in body.replace("no oracle here", "")
    pkg = os.path.join(ROOT, "efficientvideoclassification_youtube8m_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            text = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in text and "from oracle" not in text, f


def test_argument_errors_are_reported_before_any_launch():
    """The C ABI validates its arguments on the host and reports through the int return code +
    evc_last_error() (no exceptions, no launch): checked here without a GPU."""
    from efficientvideoclassification_youtube8m_b200 import _lib
    lib = _lib.lib
    # evc_lstm_seq_fwd_steps: empty problem, bad step range, H not a multiple of 64
    args = dict(x=None, stride=0, Kx=128, W=None, bias=None, rows=128, H=128, T=4, t0=0, t1=4)

    def call(**kw):
        a = dict(args, **kw)
        return lib.evc_lstm_seq_fwd_steps(a["x"], a["stride"], a["Kx"], a["W"], a["bias"], a["rows"], a["H"], a["T"],
                                          a["t0"], a["t1"], None, None, None, None, None, 0, None, None, None, None,
                                          None)
    assert call(rows=0) == -1 and b"empty" in lib.evc_last_error()
    for t0, t1 in ((-1, 2), (3, 2), (0, 5), (2, 2)):
        assert call(t0=t0, t1=t1) == -1 and b"step range" in lib.evc_last_error()
    assert call(H=100) == -1 and b"multiples of 64" in lib.evc_last_error()
    with pytest.raises(_lib.EvcError, match="step range"):
        _lib.check(call(t0=3, t1=1), "evc_lstm_seq_fwd_steps")
    # evc_frames_pack: frames must split evenly into chunks (tf.split), features a multiple of 4
    assert lib.evc_frames_pack(None, 2, 300, 128, None, 0, 30, 7, 1, None, None, None, None) == -1
    assert b"split evenly" in lib.evc_last_error()
    assert lib.evc_frames_pack(None, 2, 300, 126, None, 0, 30, 5, 1, None, None, None, None) == -1
    # split-bf16 mode: the residual planes come as a set
    one = ctypes.c_void_p(16)
    assert lib.evc_gemm_bf16x2(one, None, 0, 64, one, one, 0, 64, 128, 128, 64, one, 0, 128, None, 1, 0, None) == -1
    assert b"lo planes" in lib.evc_last_error()
    assert lib.evc_lstm_seq_fwd_steps(None, 0, 128, None, None, 128, 128, 4, 0, 4, None, None, None, None, None, 0,
                                      one, None, None, None, None) == -1 and b"split-bf16" in lib.evc_last_error()
    assert lib.evc_lstm_seq_fwd_resident(None, 0, 128, None, None, 100000, 1024, 4, None, None, None, None, None, 0,
                                         None) == -3 and b"not eligible" in lib.evc_last_error()
    # evc_topk: k must be positive and <= num_classes (eval_util.py:103-104)
    assert lib.evc_topk(None, 4, 100, 0, None, None, None, None, None) == -1
    assert lib.evc_topk(None, 4, 100, 101, None, None, None, None, None) == -1


def test_overlap_mode_default_and_env(monkeypatch):
    from efficientvideoclassification_youtube8m_b200 import engine
    monkeypatch.delenv("EVC_OVERLAP", raising=False)
    assert engine.overlap_mode() == (engine.OVERLAP_STUDENT | engine.OVERLAP_CELLS | engine.OVERLAP_WGRAD)
    monkeypatch.setenv("EVC_OVERLAP", "0")
    assert engine.overlap_mode() == 0


def test_reader_library_exports_every_declared_symbol():
    """include/evc_reader.h <-> libevc_reader.so (the native host-side input path; built by build() with g++)."""
    import __graft_entry__ as g
    g.build()
    src = open(os.path.join(ROOT, "include", "evc_reader.h")).read()
    syms = sorted(set(re.findall(r"\b(evc_[a-z0-9_]+)\s*\(", src)))
    assert len(syms) == 8, syms
    lib = ctypes.CDLL(os.path.join(ROOT, "efficientvideoclassification_youtube8m_b200", "libevc_reader.so"))
    for s in syms:
        assert hasattr(lib, s), f"libevc_reader.so does not export {s}"
    lib.evc_reader_version.restype = ctypes.c_int
    assert lib.evc_reader_version() >= 1
    # argument errors come back as NULL / negative codes with a message, never as a crash
    lib.evc_reader_open.restype = ctypes.c_void_p
    lib.evc_reader_last_error.restype = ctypes.c_char_p
    assert lib.evc_reader_open(None, 0, None, None, 0, 10, 300, 1, 0) is None
    assert b"feature_names is empty" in lib.evc_reader_last_error()
    assert lib.evc_reader_next(None, 4, None, None, None, None, 0) == -1


def test_learning_rate_schedule_follows_exponential_decay_staircase():
    """train.py:222-236: lr = base * decay ** floor(global_step * batch_size / decay_examples); the flag defaults
    (decay 1) keep it constant.  global_step counts train ops: +2 per joint iteration."""
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import _Base
    from efficientvideoclassification_youtube8m_b200.train_ops import exponential_decay
    b = _Base(ModelConfig(), 256, "cpu", 10, 5, 1e-3, 1.0, 2.0)
    b.global_step = 10 ** 6
    assert b.lr == 1e-3
    b = _Base(ModelConfig(), 256, "cpu", 10, 5, 1e-3, 1.0, 2.0, learning_rate_decay=0.95,
              learning_rate_decay_examples=1000.0)
    for step, power in ((0, 0), (2, 0), (4, 1), (6, 1), (8, 2), (40, 10)):
        b.global_step = step
        assert b.lr == pytest.approx(1e-3 * 0.95 ** power, rel=1e-12)
        assert b.lr == pytest.approx(exponential_decay(1e-3, step * 256, 1000.0, 0.95), rel=1e-12)


def test_evaluation_loop_pads_the_last_batch_and_drops_the_padding():
    """steps.evaluation_loop (eval_finetune.py:240-275) with stand-ins for the evaluator and the metrics: the
    plan's batch size is fixed, so the epoch's last batch is padded with empty videos whose rows never reach
    the metrics."""
    import torch
    from efficientvideoclassification_youtube8m_b200.steps import evaluation_loop

    class Evaluator:
        B, device = 4, torch.device("cpu")

        def __init__(self):
            self.seen = []

        def step(self, x, nf, y):
            assert x.shape[0] == self.B and nf.shape[0] == self.B and y.shape[0] == self.B
            self.seen.append(nf.tolist())
            self.rows = x.float().sum(dim=(1, 2))
            return x.float().mean(dim=1), None, None, None

    class Metrics:
        def __init__(self):
            self.calls, self.cleared = [], 0

        def clear(self):
            self.cleared += 1

        def accumulate(self, p, y, loss):
            self.calls.append((p.shape[0], y.shape[0], loss.shape[0], float(loss.sum())))
            return {"hit_at_one": 1.0, "perr": 0.5, "loss": float(loss.mean())}

        def get(self):
            return {"gap": 0.25, "avg_loss": 1.0}

    batches = [(["a", "b", "c", "d"], torch.ones(4, 3, 2, dtype=torch.uint8), torch.zeros(4, 5, dtype=torch.bool),
                torch.tensor([3, 2, 1, 3], dtype=torch.int32)),
               (["e"], torch.full((1, 3, 2), 2, dtype=torch.uint8), torch.ones(1, 5, dtype=torch.bool),
                torch.tensor([2], dtype=torch.int32))]
    ev, m, lines = Evaluator(), Metrics(), []
    out = evaluation_loop(ev, batches, m, log=lines.append)
    assert m.cleared == 1 and ev.seen == [[3, 2, 1, 3], [2, 0, 0, 0]]
    assert m.calls == [(4, 4, 4, 24.0), (1, 1, 1, 12.0)]
    assert out["gap"] == 0.25 and out["examples_processed"] == 5 and out["examples_per_second"] > 0
    assert len(lines) == 2 and lines[1].startswith("examples_processed: 5 | hit_at_one: 1")
    with pytest.raises(ValueError, match="exceeds"):
        evaluation_loop(ev, [(["x"] * 5, torch.ones(5, 3, 2), torch.zeros(5, 5), torch.zeros(5, dtype=torch.int32))], m)
