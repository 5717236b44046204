"""The five entry points (run_train.sh -> run_convert_model.sh -> run_finetune.sh -> run_eval.sh / run_validate.sh)
end to end on tiny TFRecord shards: flags, training loop with per-step device metrics, TF-format checkpoints,
resume, student conversion, evaluation epochs; and the device-resident metrics against the host ones."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _shards(tmp_path, prefix, n_shards, per_shard, seed):
    from efficientvideoclassification_youtube8m_b200 import readers
    rng = np.random.default_rng(seed)
    for k in range(n_shards):
        recs = []
        for i in range(per_shard):
            n = int(rng.integers(5, 301))
            f = {"rgb": rng.integers(0, 256, size=(n, 96), dtype=np.uint8),
                 "audio": rng.integers(0, 256, size=(n, 32), dtype=np.uint8)}
            recs.append(readers.make_sequence_example(f"{prefix}{k}_{i}", sorted(rng.choice(4716, 3, replace=False).tolist()), f))
        readers.write_tfrecord(str(tmp_path / f"{prefix}{k}.tfrecord"), recs, with_crc=True)


COMMON = ["--frame_features", "True", "--feature_names", "rgb, audio", "--feature_sizes", "96, 32", "--model",
          "HierarchicalLstmModel", "--num_inputs_to_lstm", "20", "--lstm_layers", "2", "--lstm_cells", "128",
          "--every_n", "10", "--batch_size", "8"]


def test_train_convert_finetune_eval_validate_pipeline(tmp_path, capsys):
    from efficientvideoclassification_youtube8m_b200 import launchers, tf_checkpoint
    from efficientvideoclassification_youtube8m_b200.flags import FLAGS
    _shards(tmp_path, "train", 3, 8, 1)
    _shards(tmp_path, "validate", 2, 5, 2)
    tdir, fdir = str(tmp_path / "ts") + "/", str(tmp_path / "ft") + "/"
    try:
        FLAGS.reset()
        info = launchers.train_main(COMMON + ["--train_data_pattern", str(tmp_path / "train*.tfrecord"), "--train_dir",
                                              tdir, "--start_new_model", "True", "--num_epochs", "1"])
        assert info["global_step"] == 6                                  # 24 videos / batch 8 = 3 iterations x 2 train ops
        for k in ("hit_at_one", "perr", "gap", "teacher_loss", "l_rep", "l_pred", "l_ce"):
            assert np.isfinite(info[k]), k
        ck = tf_checkpoint.latest_checkpoint(tdir)
        assert ck.endswith("model.ckpt-6") and os.path.exists(ck + ".index") and os.path.exists(ck + ".data-00000-of-00001")
        var = tf_checkpoint.list_variables(ck)
        assert var["model/RNN_L2/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel"][1] == (4 * 128 + 128, 512)
        assert "model_student/classifier/gates/weights/Adam_1" in var and "global_step" in var and "beta1_power_1" in var
        # resume: one more epoch continues from global_step 6 with the optimizer state
        FLAGS.reset()
        info2 = launchers.train_main(COMMON + ["--train_data_pattern", str(tmp_path / "train*.tfrecord"), "--train_dir",
                                               tdir, "--start_new_model", "False", "--num_epochs", "1"])
        assert info2["global_step"] == 12 and tf_checkpoint.latest_checkpoint(tdir).endswith("model.ckpt-12")
        assert not os.path.exists(ck + ".index")                           # Saver(max_to_keep=1)
        assert info2["l_ce"] < info["l_ce"] * 1.5
        # T+S -> student-only
        FLAGS.reset()
        prefix = launchers.convert_main(["--train_dir", tdir, "--output_dir", fdir])
        sv = tf_checkpoint.load_variables(prefix)
        assert len(sv) == 12 and int(sv["global_step"]) == 0
        both = tf_checkpoint.load_variables(tf_checkpoint.latest_checkpoint(tdir), ["model_student/classifier/experts/biases"])
        assert np.array_equal(sv["model_student/classifier/experts/biases"], both["model_student/classifier/experts/biases"])
        # fine-tune from the converted checkpoint
        FLAGS.reset()
        info3 = launchers.finetune_main(COMMON + ["--train_data_pattern", str(tmp_path / "train*.tfrecord"),
                                                  "--train_dir", fdir, "--start_new_model", "False", "--num_epochs", "1"])
        assert info3["global_step"] == 3 and np.isfinite(info3["student_loss"])
        # eval_finetune / validate epochs over 10 videos (batches 8 + 2: the ragged one is padded, nothing dropped)
        FLAGS.reset()
        ev = launchers.eval_main(COMMON + ["--eval_data_pattern", str(tmp_path / "validate*.tfrecord"), "--train_dir",
                                           fdir, "--run_once", "True", "--top_k", "20"])
        assert ev["examples_processed"] == 10 and 0.0 <= ev["gap"] <= 1.0 and ev["epoch_id"] == "3"
        FLAGS.reset()
        va = launchers.validate_main(COMMON + ["--eval_data_pattern", str(tmp_path / "validate*.tfrecord"),
                                               "--train_dir", tdir, "--run_once", "True", "--top_k", "20"])
        assert va["examples_processed"] == 10 and va["avg_student_state_loss"] >= 0.0 and va["epoch_id"] == "12"
    finally:
        FLAGS.reset()
    text = capsys.readouterr().out
    assert "training step 2| Hit@1:" in text and "| Teacher_Loss:" in text and "| L_REP:" in text
    assert "Student_Label_Loss" in text and "examples_processed: 10" in text and "Total time taken is" in text


def test_checkpoint_restores_identical_training_state(tmp_path):
    """Saver round trip of weights + Adam slots + step counters through the TF bundle format: a restored trainer
    takes bit-identical steps."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import launchers
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    cfg = ModelConfig(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)
    x, nf, lab = O.synthetic_batch(8, seed=5, num_features=128, vocab_size=200)
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    a = TeacherStudentTrainer(cfg, batch_size=8, )
    for _ in range(3):
        a.step(xd, nfd, labd)
    prefix = launchers.save_checkpoint(str(tmp_path), [a.teacher, a.student], a.global_step)
    b = TeacherStudentTrainer(cfg, batch_size=8, teacher_seed=None, student_seed=None)
    b.global_step = launchers.restore_checkpoint(prefix, [b.teacher, b.student])
    assert b.global_step == 6 and int(b.student.adam_step.item()) == 3
    for pa, pb in ((a.teacher, b.teacher), (a.student, b.student)):
        assert torch.equal(pa.flat_w, pb.flat_w) and torch.equal(pa.flat_m, pb.flat_m) and torch.equal(pa.flat_v, pb.flat_v)
        for n in pa.shadow:
            assert torch.equal(pa.shadow[n], pb.shadow[n])
    # the forward of the next step is deterministic: identical predictions from the restored state
    a.step(xd, nfd, labd)
    b.step(xd, nfd, labd)
    assert torch.equal(a.s_eng.pred, b.s_eng.pred) and torch.equal(a.t_eng.pred, b.t_eng.pred)


def test_device_batch_metrics_match_host_and_golden():
    """evc_batch_metrics (hit@1 / PERR / GAP on the device, no sort) against the reference-generated golden values,
    the host implementations of eval_util, and the oracle; EvaluationMetrics' device-resident accumulation against the
    host accumulation over several batches with ties and label-less videos."""
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200 import eval_util, ops
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.npz"))
    for case in ("small", "yt8m", "ties"):
        p, y = gold[case + "/predictions"], gold[case + "/labels"]
        pd_, yd = torch.from_numpy(p).cuda(), torch.from_numpy((y != 0).astype(np.uint8)).cuda()
        bm = ops.BatchMetrics(p.shape[0], p.shape[1], 20, "cuda")
        out = bm.run(pd_, yd)[0].tolist()
        assert abs(out[0] - O.hit_at_one(p, y)) < 1e-6
        assert abs(out[1] - O.perr(p, y)) < 1e-6
        assert abs(out[2] - O.gap(p, y, 20)) < 1e-5
        if case != "ties":          # (the reference shuffles ties at random; the golden values hold without ties)
            assert abs(out[0] - float(gold[case + "/hit_at_one"])) < 1e-6
            assert abs(out[1] - float(gold[case + "/perr"])) < 1e-6
            assert abs(out[2] - float(gold[case + "/gap"])) < 1e-5
    rng = np.random.default_rng(3)
    V = 300
    host, dev = eval_util.EvaluationMetrics(V, 20), eval_util.EvaluationMetrics(V, 20)
    for n in (16, 16, 5):
        p = np.round(rng.random((n, V)), 2).astype(np.float32)        # two decimals: many exact ties
        y = (rng.random((n, V)) < 0.02)
        y[0] = False                                                   # a video without labels
        loss = rng.random(n).astype(np.float32)
        a = host.accumulate_stats(eval_util.batch_stats(p, y, loss, 20))
        b = dev.accumulate(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(loss).cuda())
        for k in a:
            assert abs(a[k] - b[k]) < 1e-5, k
    assert dev.num_examples == 0 and len(dev._dev_triplets) == 3        # nothing reached the host calculators yet
    ha, da = host.get(), dev.get()
    assert dev.num_examples == 37
    for k in ("avg_hit_at_one", "avg_perr", "avg_loss", "gap"):
        assert abs(ha[k] - da[k]) < 1e-6, k
    assert np.allclose(ha["aps"], da["aps"], atol=1e-12)
