"""The reference's own control flow (train.py:253-418: variable scopes, create_model /
create_model_inference, CrossEntropyLoss, L_REP, L_PRED, two create_train_op's) written against
this package's plugin classes, checked against the oracle and against the fused step path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(feature_size=128, lstm_cells=128, vocab_size=200, num_mixtures=2)


def _build_and_run(x, nf, lab, steps):
    """train.py build_graph + Trainer loop, line for line, on the plugin surface."""
    from efficientvideoclassification_youtube8m_b200 import (frame_level_models, losses, nn_ops, scope, train_ops,
                                                             video_level_models)
    from efficientvideoclassification_youtube8m_b200.flags import FLAGS
    FLAGS.reset()
    FLAGS.lstm_cells, FLAGS.every_n, FLAGS.batch_size = KW["lstm_cells"], 10, x.shape[0]
    scope.reset_default_graph()
    train_ops.reset_global_step()

    def find_class_by_name(name, modules):
        modules = [getattr(module, name, None) for module in modules]
        return next(a for a in modules if a)

    model = find_class_by_name(FLAGS.model, [frame_level_models, video_level_models])()
    label_loss_fn = find_class_by_name(FLAGS.label_loss, [losses])()
    optimizer = train_ops.AdamOptimizer(FLAGS.base_learning_rate)
    optimizer_student = train_ops.AdamOptimizer(FLAGS.base_learning_rate)
    model_input_raw, num_frames, labels_batch = x, nf, lab
    out = []
    for _ in range(steps):
        model_input = nn_ops.l2_normalize(model_input_raw)
        num_frames_student = nn_ops.num_frames_student(num_frames, FLAGS.every_n)
        model_input_student = nn_ops.sample_every_n(model_input, FLAGS.every_n)
        with scope.variable_scope("model"):
            teacher_state, result = model.create_model(model_input, num_frames=num_frames, vocab_size=KW["vocab_size"],
                                                       batch_size=FLAGS.batch_size, labels=labels_batch, dropout=0.5)
            predictions = result["predictions"]
            label_loss = label_loss_fn.calculate_loss(predictions, labels_batch)
            reg_loss = losses.regularization_loss("model")
            final_loss = FLAGS.regularization_penalty * reg_loss + label_loss
            train_op = train_ops.create_train_op(final_loss, optimizer,
                                                 variables_to_train=scope.trainable_variables("model"),
                                                 clip_gradient_norm=FLAGS.clip_gradient_norm)
        with scope.variable_scope("model_student"):
            student_state, student_results = model.create_model_inference(
                model_input_student, num_frames=num_frames_student, vocab_size=KW["vocab_size"],
                batch_size=FLAGS.batch_size, labels=labels_batch, every_n=FLAGS.every_n, num_inputs_L1=5, dropout=0.5)
            student_loss_state = losses.representation_matching_loss(teacher_state, student_state)
            student_predictions = student_results["predictions"]
            student_label_loss = label_loss_fn.calculate_loss(student_predictions, labels_batch)
            stud_reg_loss = losses.regularization_loss("model_student")
            pred_loss = losses.prediction_matching_loss(predictions, student_predictions)
            total_student_loss = (student_loss_state + pred_loss + student_label_loss + student_loss_state +
                                  FLAGS.regularization_penalty * stud_reg_loss)
            train_student_op = train_ops.create_train_op(
                total_student_loss, optimizer_student,
                variables_to_train=scope.trainable_variables("model_student"),
                clip_gradient_norm=FLAGS.clip_gradient_norm)
        loss_t = train_op.run()
        loss_s = train_student_op.run()
        out.append(dict(teacher_loss=float(loss_t), student_loss=float(loss_s), l_ce=float(student_label_loss),
                        l_rep=float(student_loss_state), l_pred=float(pred_loss),
                        global_step=train_ops.get_global_step(), predictions=student_predictions.detach().clone()))
    FLAGS.reset()
    return out


def test_reference_control_flow_matches_oracle():
    from oracle import hlstm_oracle as O
    B, steps = 16, 3
    x, nf, lab = O.synthetic_batch(B, seed=11, num_features=KW["feature_size"], vocab_size=KW["vocab_size"])
    got = _build_and_run(torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda(), steps)
    T = O.init_params("model", 0, dtype=torch.float64, **KW)
    S = O.init_params("model_student", 1, dtype=torch.float64, **KW)
    ot, os_ = O.TFAdam(T), O.TFAdam(S)
    for it in range(steps):
        ref = O.teacher_student_train_step(torch.from_numpy(x).double(), nf, torch.from_numpy(lab), T, S, ot, os_,
                                           vocab_size=KW["vocab_size"], num_mixtures=KW["num_mixtures"])
        for k in ("teacher_loss", "student_loss", "l_ce", "l_rep"):
            want = float(ref[k])
            assert abs(got[it][k] - want) <= 0.01 * abs(want) + 1e-4, (it, k, got[it][k], want)
        err = (got[it]["predictions"].cpu().double() - ref["student_predictions"]).abs().max().item()
        assert err < 1e-3, (it, err)
    assert got[-1]["global_step"] == 2 * steps


def test_plugin_path_equals_fused_step_path():
    from oracle import hlstm_oracle as O
    from efficientvideoclassification_youtube8m_b200.params import ModelConfig
    from efficientvideoclassification_youtube8m_b200.steps import TeacherStudentTrainer
    B = 16
    x, nf, lab = O.synthetic_batch(B, seed=12, num_features=KW["feature_size"], vocab_size=KW["vocab_size"])
    xd, nfd, labd = torch.from_numpy(x).cuda(), torch.from_numpy(nf).cuda(), torch.from_numpy(lab).cuda()
    got = _build_and_run(xd, nfd, labd, 2)
    tr = TeacherStudentTrainer(ModelConfig(**KW), batch_size=B)
    for it in range(2):
        tr.step(xd, nfd, labd)
        f = tr.fetch()
        for k in ("teacher_loss", "student_loss", "l_ce", "l_rep"):
            assert abs(got[it][k] - f[k]) <= 2e-4 * abs(f[k]) + 1e-5, (it, k, got[it][k], f[k])


def test_plugin_argument_errors():
    from efficientvideoclassification_youtube8m_b200 import frame_level_models, scope
    scope.reset_default_graph()
    m = frame_level_models.HierarchicalLstmModel()
    x = torch.zeros(4, 300, 128, device="cuda")
    with pytest.raises(RuntimeError):                       # no variable scope
        m.create_model(x, 200, torch.zeros(4, dtype=torch.int32, device="cuda"))
    with scope.variable_scope("model"):
        with pytest.raises(TypeError):                      # teacher lengths are int32
            m.create_model(x, 200, torch.zeros(4, dtype=torch.int64, device="cuda"))
        with pytest.raises(TypeError):                      # host tensors are not accepted (no CPU fallback)
            m.create_model(x.cpu(), 200, torch.zeros(4, dtype=torch.int32))
