"""The reference's five entry points as drivers of the fused steps (same flag names, same log lines):

  train_main      <- train.py:428-560,704-737          (run_train.sh)      joint teacher+student training
  finetune_main   <- train_finetune.py:325-450         (run_finetune.sh)   student-only fine-tuning
  validate_main   <- validate.py:192-404               (run_validate.sh)   teacher+student evaluation
  eval_main       <- eval_finetune.py:177-350          (run_eval.sh)       student evaluation
  convert_main    <- train_convert_model.py:428-520    (run_convert_model.sh)  T+S checkpoint -> student-only

What TF's Supervisor / queue runners / Saver did around the step is plain host code here: the reader generator
(readers.get_input_data_batches), a checkpoint every `save_model_secs` and at the end of training in TF's own
tensor-bundle format (tf_checkpoint.py: variables under their TF names, Adam slots `<var>/Adam`, `<var>/Adam_1`,
`beta1_power`, `beta2_power`, `global_step`), `checkpoint` state file, restore of the latest checkpoint unless
--start_new_model.  One process per GPU (torchrun): ranks read disjoint shards, gradients are averaged by NCCL inside
the step, rank 0 logs and writes checkpoints.  The per-step Hit@1 / PERR / GAP of train.py:522-526 come from the
device (evc_topk + evc_batch_metrics) and travel with the losses in one small copy.
"""
from __future__ import annotations

import logging
import os
import sys
import time
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops, readers, tf_checkpoint
from .eval_util import EvaluationMetrics
from .flags import FLAGS
from .params import HLstmParams, ModelConfig
from .steps import (StudentEvaluator, StudentFinetuneTrainer, TeacherStudentEvaluator, TeacherStudentTrainer,
                    evaluation_loop)

log = logging.getLogger("evc")


class _StdoutHandler(logging.Handler):
    """Log lines go to whatever sys.stdout is at emit time (the reference logs through tf.logging to the
    console; run_*.sh redirect it into the output_* files)."""

    def emit(self, record):
        try:
            sys.stdout.write(self.format(record) + "\n")
        except Exception:       # noqa: BLE001
            self.handleError(record)


# ------------------------------------------------------------------ environment
def _setup_logging():
    if not any(isinstance(h, _StdoutHandler) for h in log.handlers):
        log.addHandler(_StdoutHandler())
        log.setLevel(logging.INFO)
        log.propagate = False


def _init(argv: Optional[List[str]]):
    FLAGS.parse(list(sys.argv[1:] if argv is None else argv))
    _setup_logging()
    if not torch.cuda.is_available():
        raise RuntimeError("the H-LSTM path runs on a CUDA device (sm_100a); there is no CPU fallback")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(FLAGS.gpu if world == 1 else 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    for k in sorted(FLAGS._defs):                      # train.py:706-707 prints the flags
        if rank == 0:
            print("Key: %s Value: %s" % (k, getattr(FLAGS, k)))
    return rank, world, dev


def get_list_of_feature_names_and_sizes(feature_names: str, feature_sizes: str):
    """utils.GetListOfFeatureNamesAndSizes (utils.py:130-148)."""
    names = [n.strip() for n in feature_names.split(",")]
    sizes = [int(s) for s in feature_sizes.split(",")]
    if len(names) != len(sizes):
        log.error("length of the feature names (=%d) != length of feature sizes (=%d)", len(names), len(sizes))
    return names, sizes


def _reader():
    if not FLAGS.frame_features:
        raise NotImplementedError("the hot path is the frame-level H-LSTM (--frame_features True)")
    names, sizes = get_list_of_feature_names_and_sizes(FLAGS.feature_names, FLAGS.feature_sizes)
    return readers.YT8MFrameFeatureReader(feature_names=names, feature_sizes=sizes, max_frames=FLAGS.max_num_frames)


def _model_config(reader) -> ModelConfig:
    if FLAGS.model != "HierarchicalLstmModel" or FLAGS.video_level_classifier_model != "MoeModel":
        raise NotImplementedError("this library implements --model HierarchicalLstmModel with the MoeModel classifier")
    if FLAGS.label_loss != "CrossEntropyLoss" or FLAGS.optimizer != "AdamOptimizer":
        raise NotImplementedError("--label_loss CrossEntropyLoss and --optimizer AdamOptimizer are the fused ones")
    return ModelConfig(feature_size=sum(reader.feature_sizes), lstm_cells=FLAGS.lstm_cells,
                       lstm_layers=FLAGS.lstm_layers, vocab_size=reader.num_classes,
                       num_mixtures=FLAGS.moe_num_mixtures)


# ------------------------------------------------------------------ checkpoints
def checkpoint_variables(param_sets: List[HLstmParams], global_step: int) -> Dict[str, np.ndarray]:
    """name -> array exactly as tf.train.Saver() names them (collective when the optimizer is sharded)."""
    out = {"global_step": np.array(global_step, dtype=np.int64)}
    for i, p in enumerate(param_sets):
        p.sync_all(include_slots=True)
        for n in p.names:
            out[n] = p.w[n].detach().cpu().numpy()
            out[n + "/Adam"] = p.m[n].detach().cpu().numpy()
            out[n + "/Adam_1"] = p.v[n].detach().cpu().numpy()
        t = int(p.adam_step.item())
        sfx = "" if i == 0 else "_%d" % i              # TF uniquifies the second optimizer's accumulators
        out["beta1_power" + sfx] = np.array(0.9 ** (t + 1), dtype=np.float32)
        out["beta2_power" + sfx] = np.array(0.999 ** (t + 1), dtype=np.float32)
        out["evc/%s/adam_step" % p.scope] = np.array(t, dtype=np.int64)      # beta1^t underflows float32 after ~830 steps
    return out


def save_checkpoint(train_dir: str, param_sets: List[HLstmParams], global_step: int, write: bool = True) -> str:
    prefix = os.path.join(train_dir, "model.ckpt-%d" % global_step)
    var = checkpoint_variables(param_sets, global_step)
    if write:
        os.makedirs(train_dir, exist_ok=True)
        old = tf_checkpoint.latest_checkpoint(train_dir)
        tf_checkpoint.save_variables(prefix, var)
        tf_checkpoint.update_checkpoint_state(train_dir, prefix)
        if old and old != prefix:                      # tf.train.Saver(max_to_keep=1), train.py:651
            for ext in (".index", ".data-00000-of-00001", ".npz", ".meta"):
                if os.path.exists(old + ext):
                    os.remove(old + ext)
    return prefix


def restore_checkpoint(prefix: str, param_sets: List[HLstmParams], from_scope: Optional[Dict[str, str]] = None,
                       with_slots: bool = True) -> int:
    """Load the variables of every parameter set from `<prefix>` (TF bundle, or a `.npz` name->array map); returns
    the checkpoint's global_step.  from_scope maps a parameter set's scope to the scope it is read from."""
    if os.path.exists(prefix + ".index"):
        have = tf_checkpoint.list_variables(prefix)
        load = lambda names: tf_checkpoint.load_variables(prefix, names)             # noqa: E731
    else:
        z = np.load(prefix + ".npz")
        have = {k: None for k in z.files}
        load = lambda names: {n: z[n] for n in names}                                 # noqa: E731
    for p in param_sets:
        src = (from_scope or {}).get(p.scope, p.scope)
        names = {n: src + n[len(p.scope):] for n in p.names}
        missing = [v for v in names.values() if v not in have]
        if missing:
            raise KeyError(f"checkpoint {prefix} lacks {missing[:3]}{'...' if len(missing) > 3 else ''}")
        got = load(list(names.values()))
        p.load_state_dict({n: got[names[n]] for n in p.names})
        slots = [names[n] + s for n in p.names for s in ("/Adam", "/Adam_1")]
        if with_slots and all(s in have for s in slots):
            got = load(slots)
            for n in p.names:
                p.m[n].copy_(torch.from_numpy(np.asarray(got[names[n] + "/Adam"])))
                p.v[n].copy_(torch.from_numpy(np.asarray(got[names[n] + "/Adam_1"])))
            key = "evc/%s/adam_step" % src
            if key in have:
                p.adam_step.fill_(int(load([key])[key]))
            elif "beta2_power" in have:
                b2 = float(load(["beta2_power"])["beta2_power"])
                p.adam_step.fill_(max(int(round(np.log(b2) / np.log(0.999))) - 1, 0) if b2 > 0 else 0)
    return int(load(["global_step"])["global_step"]) if "global_step" in have else 0


# ------------------------------------------------------------------ training loops
class Trainer:
    """train.py:428-560 (joint) / train_finetune.py:325-450 (finetune=True)."""

    save_model_secs = 30 * 60                          # train.py:503

    def __init__(self, rank: int, world: int, device, train_dir: str, finetune: bool = False):
        self.rank, self.world, self.device, self.train_dir, self.finetune = rank, world, device, train_dir, finetune
        self.is_master = rank == 0
        self.reader = _reader()
        cfg = _model_config(self.reader)
        kw = dict(batch_size=FLAGS.batch_size, device=device, every_n=FLAGS.every_n,
                  base_learning_rate=FLAGS.base_learning_rate, clip_gradient_norm=FLAGS.clip_gradient_norm,
                  regularization_penalty=FLAGS.regularization_penalty, learning_rate_decay=FLAGS.learning_rate_decay,
                  learning_rate_decay_examples=FLAGS.learning_rate_decay_examples, sampling=FLAGS.sampling,
                  precise=FLAGS.precise)
        if finetune:
            self.step_fn = StudentFinetuneTrainer(cfg, **kw)
            self.param_sets = [self.step_fn.student]
        else:
            self.step_fn = TeacherStudentTrainer(cfg, num_inputs_to_lstm=FLAGS.num_inputs_to_lstm, **kw)
            self.param_sets = [self.step_fn.teacher, self.step_fn.student]
        self.metrics = ops.BatchMetrics(FLAGS.batch_size, cfg.vocab_size, FLAGS.top_k, device)

    def restore(self, start_new_model: bool) -> None:
        """train.py:583-600: keep training from the latest checkpoint unless --start_new_model."""
        if start_new_model:
            log.info("Flag 'start_new_model' is set. Building a new model.")
            return
        latest = tf_checkpoint.latest_checkpoint(self.train_dir)
        if not latest:
            log.info("No checkpoint file found. Building a new model.")
            return
        # fine-tuning starts from the converted student-only checkpoint: weights only when it has no slots
        self.step_fn.global_step = restore_checkpoint(latest, self.param_sets)
        log.info("Restored %s (global_step %d)", latest, self.step_fn.global_step)

    def run(self, start_new_model: bool = False, max_steps: Optional[int] = None) -> Dict[str, float]:
        tr, B = self.step_fn, FLAGS.batch_size
        self.restore(start_new_model)
        start, last_save, steps, info = time.time(), time.time(), 0, {}
        pred_eng = tr.s_eng if self.finetune else tr.t_eng     # the "predictions" collection (train.py:289 / train_finetune.py:279)
        batches = readers.get_input_data_batches(self.reader, FLAGS.train_data_pattern, B, FLAGS.num_epochs,
                                                 rank=self.rank, world=self.world, uneven="min",
                                                 drop_remainder=True, prefetch=2, num_threads=FLAGS.num_readers)
        log.info("Entering training loop.")
        for ids, x, y, nf in batches:                  # fixed-size plan: the epoch's ragged last batch is dropped
            t0 = time.time()
            xd, yd, nd = (t.to(self.device, non_blocking=True) for t in (x, y, nf))
            tr.step(xd, nd, yd)
            m, _, _, _ = self.metrics.run(pred_eng.pred, yd.view(torch.uint8))
            vals = tr.fetch()                          # device->host: the reference's sess.run fetch
            hit, perr, gap, _ = m.tolist()
            dt = time.time() - t0
            steps += 1
            info = dict(vals, hit_at_one=hit, perr=perr, gap=gap, examples_per_second=B * self.world / dt)
            if self.is_master:
                if self.finetune:                      # train_finetune.py:417-421
                    log.info("training step %d| Hit@1: %.2f| PERR: %.2f| GAP: %.2f| Student_Label_Loss: %s",
                             vals["global_step"], hit, perr, gap, round(vals["l_ce"], 2))
                else:                                  # train.py:527-532
                    log.info("training step %d| Hit@1: %.2f| PERR: %.2f| GAP: %.2f| Teacher_Loss: %s| L_REP: %s"
                             "| L_PRED: %s| L_CE: %s", vals["global_step"], hit, perr, gap,
                             round(vals["teacher_loss"], 2), round(vals["l_rep"], 2), round(vals["l_pred"], 2),
                             round(vals["l_ce"], 2))
            if time.time() - last_save > self.save_model_secs:
                save_checkpoint(self.train_dir, self.param_sets, tr.global_step, write=self.is_master)
                last_save = time.time()
            if max_steps is not None and steps >= max_steps:
                break
        log.info("Done training -- epoch limit reached.")
        save_checkpoint(self.train_dir, self.param_sets, tr.global_step, write=self.is_master)
        log.info("Exited training loop.")
        if self.is_master:
            print("Total time taken is " + str(time.time() - start))
        return info


def train_main(argv: Optional[List[str]] = None, finetune: bool = False, max_steps: Optional[int] = None):
    rank, world, dev = _init(argv)
    if not FLAGS.train_dir:
        raise ValueError("--train_dir is required")
    if finetune and FLAGS.start_new_model:
        log.info("fine-tuning from scratch: --start_new_model is set")
    out = Trainer(rank, world, dev, FLAGS.train_dir, finetune).run(FLAGS.start_new_model, max_steps)
    if world > 1:
        dist.destroy_process_group()
    return out


def finetune_main(argv: Optional[List[str]] = None, max_steps: Optional[int] = None):
    return train_main(argv, finetune=True, max_steps=max_steps)


# ------------------------------------------------------------------ evaluation loops
def _evaluate(argv, both: bool):
    """validate.py:306-404 (both=True) / eval_finetune.py:283-350: evaluate the latest checkpoint of --train_dir; with
    --run_once False keep polling for new checkpoints (skipping a global_step already evaluated)."""
    rank, world, dev = _init(argv)
    if FLAGS.eval_data_pattern == "":
        raise IOError("'eval_data_pattern' was not specified. Nothing to evaluate.")
    reader = _reader()
    cfg = _model_config(reader)
    student = HLstmParams("model_student", cfg, dev, seed=None, precise=FLAGS.precise)
    sets = [student]
    if both:
        teacher = HLstmParams("model", cfg, dev, seed=None, precise=FLAGS.precise)
        sets = [teacher, student]
        ev = TeacherStudentEvaluator(teacher, student, FLAGS.batch_size, FLAGS.every_n, FLAGS.num_inputs_to_lstm,
                                     top_k=FLAGS.top_k, sampling=FLAGS.sampling)
    else:
        ev = StudentEvaluator(student, FLAGS.batch_size, FLAGS.every_n, top_k=FLAGS.top_k, sampling=FLAGS.sampling)
    metrics = EvaluationMetrics(reader.num_classes, FLAGS.top_k, distributed=world > 1)
    start, last_step, out = time.time(), -1, None
    while True:
        latest = tf_checkpoint.latest_checkpoint(FLAGS.train_dir)
        if not latest:
            log.info("No checkpoint file found.")
        else:
            step = latest.split("/")[-1].split("-")[-1]
            if step == last_step:
                log.info("skip this checkpoint global_step_val=%s (same as the previous one).", step)
            else:
                log.info("Loading checkpoint for eval: " + latest)
                restore_checkpoint(latest, sets, with_slots=False)
                log.info("enter eval_once loop global_step_val = %s. ", step)
                batches = readers.get_input_data_batches(reader, FLAGS.eval_data_pattern, FLAGS.batch_size, 1,
                                                         shuffle=False, rank=rank, world=world, uneven="pad",
                                                         prefetch=2, num_threads=FLAGS.num_readers)
                out = evaluation_loop(ev, batches, metrics, log=log.info if rank == 0 else None)
                out["epoch_id"] = step
                log.info("Done with batched inference. Now calculating global performance metrics.")
                if rank == 0:
                    log.info("epoch/eval number %s | Avg_Hit@1: %.3f | Avg_PERR: %.3f | MAP: %.3f | GAP: %.3f | "
                             "Avg_Loss: %3f", step, out["avg_hit_at_one"], out["avg_perr"], float(np.mean(out["aps"])),
                             out["gap"], out["avg_loss"])
                    log.info("Average examples processed in one second %0.20f" % out["examples_per_second"])
                last_step = step
        if FLAGS.run_once:
            break
        time.sleep(60)
    if rank == 0:
        print("Total time taken is " + str(time.time() - start))
    if world > 1:
        dist.destroy_process_group()
    return out


def validate_main(argv: Optional[List[str]] = None):
    return _evaluate(argv, both=True)


def eval_main(argv: Optional[List[str]] = None):
    return _evaluate(argv, both=False)


# ------------------------------------------------------------------ T+S checkpoint -> student-only checkpoint
def convert_main(argv: Optional[List[str]] = None):
    """train_convert_model.py:428-520: restore the 11 `model_student/*` variables of the latest joint checkpoint in
    --train_dir and save them alone (fresh optimizer state, global_step 0) as the starting point of fine-tuning, into
    --output_dir (default: <train_dir>_finetune, the directory run_finetune.sh trains in)."""
    FLAGS.parse(list(sys.argv[1:] if argv is None else argv))
    _setup_logging()
    latest = tf_checkpoint.latest_checkpoint(FLAGS.train_dir)
    if not latest:
        raise IOError("No checkpoint file found in " + FLAGS.train_dir)
    scope = "model_student/"
    have = (tf_checkpoint.list_variables(latest) if os.path.exists(latest + ".index")
            else {k: None for k in np.load(latest + ".npz").files})
    names = [n for n in have if n.startswith(scope) and not n.endswith(("/Adam", "/Adam_1"))]
    if len(names) != 11:
        raise KeyError(f"{latest} holds {len(names)} model_student variables, expected 11")
    var = (tf_checkpoint.load_variables(latest, names) if os.path.exists(latest + ".index")
           else {n: np.load(latest + ".npz")[n] for n in names})
    var["global_step"] = np.array(0, dtype=np.int64)
    out_dir = FLAGS.output_dir or (FLAGS.train_dir.rstrip("/") + "_finetune")
    os.makedirs(out_dir, exist_ok=True)
    prefix = os.path.join(out_dir, "model.ckpt-0")
    tf_checkpoint.save_variables(prefix, var)
    tf_checkpoint.update_checkpoint_state(out_dir, prefix)
    log.info("Saved %d student variables of %s to %s", len(names), latest, prefix)
    return prefix
