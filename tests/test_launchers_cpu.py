"""Host-side logic of the entry points that needs no GPU: flag parsing as the run_*.sh lines pass them, the T+S ->
student-only checkpoint conversion (train_convert_model.py:496-517) on TensorFlow-format checkpoint files, and
`tf.train.latest_checkpoint` semantics."""
import os

import numpy as np
import pytest


def _joint_checkpoint(tmp_path, step=36704):
    from efficientvideoclassification_youtube8m_b200 import tf_checkpoint
    from efficientvideoclassification_youtube8m_b200.params import variable_names
    rng = np.random.default_rng(1)
    var = {"global_step": np.array(step, dtype=np.int64), "beta1_power": np.array(0.9, np.float32)}
    for scope in ("model", "model_student"):
        for n in variable_names(scope):
            shape = (6,) if n.endswith(("bias", "biases")) else (4, 6)
            var[n] = rng.standard_normal(shape).astype(np.float32)
            var[n + "/Adam"] = rng.standard_normal(shape).astype(np.float32)
            var[n + "/Adam_1"] = rng.random(shape).astype(np.float32)
    prefix = str(tmp_path / ("model.ckpt-%d" % step))
    tf_checkpoint.save_variables(prefix, var)
    tf_checkpoint.update_checkpoint_state(str(tmp_path), prefix)
    return prefix, var


def test_flags_parse_like_the_run_scripts():
    from efficientvideoclassification_youtube8m_b200.flags import FLAGS
    FLAGS.reset()
    try:
        FLAGS.parse(["--train_data_pattern", "./yt8m/train*.tfrecord", "--train_dir", "./model/", "--frame_features", "True",
                     "--feature_names", "rgb, audio", "--feature_sizes", "1024, 128", "--model", "HierarchicalLstmModel",
                     "--gpu", "0", "--batch_size", "256", "--num_inputs_to_lstm", "20", "--lstm_layers", "2",
                     "--start_new_model", "True", "--num_epochs", "1", "--every_n=10", "--run_once", "False"])
        assert FLAGS.frame_features is True and FLAGS.start_new_model is True and FLAGS.run_once is False
        assert FLAGS.batch_size == 256 and FLAGS.every_n == 10 and FLAGS.num_epochs == 1
        assert FLAGS.train_dir == "./model/" and FLAGS.base_learning_rate == 0.001 and FLAGS.sampling == "uniform"
        with pytest.raises(AttributeError):
            FLAGS.parse(["--no_such_flag", "1"])
    finally:
        FLAGS.reset()
    from efficientvideoclassification_youtube8m_b200.launchers import get_list_of_feature_names_and_sizes
    assert get_list_of_feature_names_and_sizes("rgb, audio", "1024, 128") == (["rgb", "audio"], [1024, 128])


def test_convert_main_keeps_the_eleven_student_variables(tmp_path):
    from efficientvideoclassification_youtube8m_b200 import launchers, tf_checkpoint
    from efficientvideoclassification_youtube8m_b200.flags import FLAGS
    train_dir = tmp_path / "joint"
    train_dir.mkdir()
    prefix, var = _joint_checkpoint(train_dir)
    assert tf_checkpoint.latest_checkpoint(str(train_dir)) == prefix
    out_dir = str(tmp_path / "finetune")
    try:
        got_prefix = launchers.convert_main(["--train_dir", str(train_dir), "--output_dir", out_dir])
    finally:
        FLAGS.reset()
    assert got_prefix == os.path.join(out_dir, "model.ckpt-0") and tf_checkpoint.latest_checkpoint(out_dir) == got_prefix
    got = tf_checkpoint.load_variables(got_prefix)
    names = [n for n in var if n.startswith("model_student/") and not n.endswith(("/Adam", "/Adam_1"))]
    assert len(names) == 11 and sorted(got) == sorted(names + ["global_step"])
    for n in names:
        assert np.array_equal(got[n], var[n])
    assert int(got["global_step"]) == 0                    # fine-tuning starts a new global_step and fresh Adam slots
    with pytest.raises(IOError):
        launchers.convert_main(["--train_dir", str(tmp_path / "empty")])
    FLAGS.reset()


def test_latest_checkpoint_ignores_a_state_file_without_data(tmp_path):
    from efficientvideoclassification_youtube8m_b200 import tf_checkpoint
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-5"\n')
    assert tf_checkpoint.latest_checkpoint(str(tmp_path)) is None
    assert tf_checkpoint.latest_checkpoint(str(tmp_path / "missing")) is None
