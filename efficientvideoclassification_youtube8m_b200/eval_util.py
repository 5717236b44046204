"""Evaluation helpers with the reference's names and semantics (code_student_uniform/eval_util.py:17-213).
The per-video top-k selection (eval_util.py:118-124, numpy.argpartition on the host in the
reference) runs on the GPU (evc_topk, exact); only the k triplets per video leave the device."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .average_precision_calculator import (AveragePrecisionCalculator, MeanAveragePrecisionCalculator,
                                           SparseMeanAveragePrecisionCalculator)


def flatten(l):
    return [item for sublist in l for item in sublist]


def _dev(predictions, actuals):
    p = predictions if torch.is_tensor(predictions) else torch.as_tensor(np.asarray(predictions, dtype=np.float32))
    a = actuals if torch.is_tensor(actuals) else torch.as_tensor(np.asarray(actuals))
    if not p.is_cuda:
        p, a = p.cuda(), a.cuda()
    p = p.to(torch.float32).contiguous()
    a = (a if a.dtype == torch.uint8 else (a != 0).view(torch.uint8)).contiguous()
    return p, a


def top_k_device(predictions, labels, k=20):
    """(class idx int32 [B,k], prediction f32 [B,k], label u8 [B,k]) as host numpy arrays."""
    if k <= 0:
        raise ValueError("k must be a positive integer.")
    p, a = _dev(predictions, labels)
    idx, val, lab = ops.topk(p, min(k, p.shape[1]), a)
    return idx.cpu().numpy(), val.cpu().numpy(), lab.cpu().numpy()


def calculate_hit_at_one(predictions, actuals):
    """eval_util.py:17-31."""
    idx, _, lab = top_k_device(predictions, actuals, 1)
    return float(np.average(lab[:, 0].astype(np.float64)))


def calculate_precision_at_equal_recall_rate(predictions, actuals):
    """eval_util.py:34-59: precision over the top-num_labels predictions of each video."""
    p, a = _dev(predictions, actuals)
    num_labels = a.sum(dim=1).cpu().numpy().astype(np.int64)
    kmax = int(num_labels.max()) if len(num_labels) else 0
    if kmax == 0:
        return 0.0
    _, val, lab = top_k_device(p, a, kmax)
    return perr_from_top_k(val, lab, num_labels)


def perr_from_top_k(val, lab, num_labels):
    """eval_util.py:43-59 given every video's top-max(num_labels) predictions (descending) and their labels:
    mean over videos of |{top-num_labels predictions that are > 0 and labelled}| / num_labels; a video without
    labels contributes 0 (argpartition(p, -0)[-0:] is every class and no label is set)."""
    val, lab = np.asarray(val), np.asarray(lab, dtype=np.float64)
    num_labels = np.asarray(num_labels, dtype=np.int64)
    if val.shape[0] == 0:
        return 0.0
    within = np.arange(val.shape[1])[None, :] < num_labels[:, None]
    hits = np.sum(lab * (val > 0) * within, axis=1)
    return float(np.sum(hits[num_labels > 0] / num_labels[num_labels > 0]) / val.shape[0])


def top_k_triplets(predictions, labels, k=20):
    """eval_util.py:118-124 for one video: [(class, prediction, label)] (ordered by value here)."""
    p = torch.as_tensor(np.asarray(predictions, dtype=np.float32)).reshape(1, -1)
    l = torch.as_tensor(np.asarray(labels)).reshape(1, -1)
    idx, val, lab = top_k_device(p, l, k)
    return [(int(i), float(v), float(t)) for i, v, t in zip(idx[0], val[0], lab[0])]


def top_k_by_class(predictions, labels, k=20):
    """eval_util.py:82-116."""
    if k <= 0:
        raise ValueError("k must be a positive integer.")
    p, a = _dev(predictions, labels)
    num_classes = p.shape[1]
    idx, val, lab = top_k_device(p, a, k)
    out_predictions = [[] for _ in range(num_classes)]
    out_labels = [[] for _ in range(num_classes)]
    for c, v, t in zip(idx.reshape(-1), val.reshape(-1), lab.reshape(-1)):
        out_predictions[c].append(float(v))
        out_labels[c].append(float(t))
    out_true_positives = a.sum(dim=0).cpu().numpy().astype(np.float64).tolist()
    return out_predictions, out_labels, out_true_positives


def calculate_gap(predictions, actuals, top_k=20):
    """eval_util.py:61-79 global average precision over the per-video top-k."""
    gap_calculator = AveragePrecisionCalculator()
    sparse_predictions, sparse_labels, num_positives = top_k_by_class(predictions, actuals, top_k)
    gap_calculator.accumulate(flatten(sparse_predictions), flatten(sparse_labels), sum(num_positives))
    return gap_calculator.peek_ap_at_n()


def _loss_rows(loss, n, device):
    """The reference passes the batch-mean label loss (a scalar, validate.py:262-266); this package's evaluators
    keep the per-video cross-entropy rows on the device.  Either becomes an f32 device vector of n rows."""
    if torch.is_tensor(loss) and loss.is_cuda and loss.numel() == n and loss.dtype == torch.float32:
        return loss.reshape(n).contiguous()
    v = np.asarray(loss.detach().cpu().numpy() if torch.is_tensor(loss) else loss, dtype=np.float64).reshape(-1)
    v = np.full(n, float(v.mean()) if v.size else 0.0) if v.size != n else v
    return torch.as_tensor(v.astype(np.float32), device=device)


def batch_stats(predictions, labels, loss, top_k):
    """What one batch contributes to EvaluationMetrics, as small host arrays: the per-video top-k triplets, the
    positives per class, and the hit@1 / PERR / loss sums -- all computed on the GPU (one top-k launch + the
    fused metrics kernels of evc_batch_metrics) and brought to the host in ONE copy."""
    p, a = _dev(predictions, labels)
    n, V = int(a.shape[0]), int(a.shape[1])
    k = min(top_k, V)
    if n == 0:
        z = np.zeros((0, k))
        return {"n": 0, "idx": z.astype(np.int32), "val": z.astype(np.float32), "lab": z.astype(np.uint8),
                "num_positives": np.zeros(V), "hit_sum": 0.0, "perr_sum": 0.0, "loss_sum": 0.0}
    bm = ops.BatchMetrics(n, V, k, p.device, accumulate=True)
    out, idx, val, lab = bm.run(p, a, _loss_rows(loss, n, p.device))
    packed = torch.cat([idx.view(torch.uint8).reshape(-1), val.view(torch.uint8).reshape(-1), lab.reshape(-1),
                        bm.class_pos.view(torch.uint8), bm.sums.view(torch.uint8)]).cpu().numpy()
    o = 0
    idx_h = packed[o:o + n * k * 4].view(np.int32).reshape(n, k); o += n * k * 4
    val_h = packed[o:o + n * k * 4].view(np.float32).reshape(n, k); o += n * k * 4
    lab_h = packed[o:o + n * k].reshape(n, k); o += n * k
    pos_h = packed[o:o + V * 4].view(np.int32).astype(np.float64); o += V * 4
    sums = packed[o:o + 32].view(np.float64)
    return {"n": n, "idx": idx_h, "val": val_h, "lab": lab_h, "num_positives": pos_h,
            "hit_sum": float(sums[1]), "perr_sum": float(sums[2]), "loss_sum": float(sums[3])}


class EvaluationMetrics(object):
    """eval_util.py:126-213.  With `distributed=True` under an initialised torch.distributed group every
    rank evaluates its own shard of the videos and `accumulate` all-gathers the per-batch statistics
    (k triplets per video + label counts), so that all ranks hold the metrics of the global
    batch -- the same numbers one process would compute on the concatenation of the shards in rank order."""

    def __init__(self, num_class, top_k, distributed=False, group=None, vectorized=True):
        self.sum_hit_at_one = 0.0
        self.sum_perr = 0.0
        self.sum_loss = 0.0
        # vectorized: the k triplets per video are grouped by class with one numpy sort (same results and tie
        # order as the reference's per-class python lists, 20x less host time per batch); False keeps the
        # literal list-per-class accumulation of eval_util.py:107-116,157-160
        self.vectorized = vectorized
        self.map_calculator = (SparseMeanAveragePrecisionCalculator(num_class) if vectorized
                               else MeanAveragePrecisionCalculator(num_class))
        self.global_ap_calculator = AveragePrecisionCalculator()
        self.top_k = top_k
        self.num_examples = 0
        self.num_class = num_class
        self.distributed = distributed
        self.group = group
        self._dev_acc, self._dev_triplets, self._dev_examples = None, [], 0

    def accumulate(self, predictions, labels, loss):
        if torch.is_tensor(predictions) and predictions.is_cuda and not self.distributed:
            return self.accumulate_device(predictions, labels, loss)
        return self.accumulate_stats(batch_stats(predictions, labels, loss, self.top_k))

    def accumulate_device(self, predictions, labels, loss, fetch=True):
        """GPU-resident accumulation: the batch's top-k triplets stay in device memory, the hit@1 /
        PERR / loss sums and the per-class positives are added to device accumulators by evc_batch_metrics;
        nothing is copied to the host until `get()`.  fetch=True reads the batch's three numbers back for the
        log line (one 16-byte copy, the reference fetches them with every sess.run); fetch=False returns the
        device vector [hit@1, PERR, GAP, loss] and never synchronises."""
        p, a = _dev(predictions, labels)
        n = int(p.shape[0])
        if n == 0:
            return {"hit_at_one": 0.0, "perr": 0.0, "loss": 0.0}
        d = self._dev_acc
        if d is None or d.B < n or d.out.device != p.device:
            old = d
            d = self._dev_acc = ops.BatchMetrics(max(n, old.B if old else 0), self.num_class, self.top_k, p.device,
                                                 accumulate=True)
            if old is not None:
                d.class_pos.copy_(old.class_pos)
                d.sums.copy_(old.sums)
        out, idx, val, lab = d.run(p, a, _loss_rows(loss, n, p.device), n)
        self._dev_triplets.append((idx, val, lab))
        self._dev_examples += n
        if not fetch:
            return out
        h = out.tolist()
        return {"hit_at_one": h[0], "perr": h[1], "loss": h[3]}

    def _flush_device(self):
        """Move the device-resident accumulators into the host calculators (one copy per tensor, per epoch)."""
        if self._dev_acc is None or not self._dev_triplets:
            return
        rows = [int(t[0].shape[0]) for t in self._dev_triplets]
        k = self._dev_triplets[0][0].shape[1]
        idx = torch.cat([t[0].reshape(-1) for t in self._dev_triplets]).cpu().numpy().reshape(-1, k)
        val = torch.cat([t[1].reshape(-1) for t in self._dev_triplets]).cpu().numpy().reshape(-1, k)
        lab = torch.cat([t[2].reshape(-1) for t in self._dev_triplets]).cpu().numpy().reshape(-1, k)
        sums = self._dev_acc.sums.cpu().numpy()
        pos = self._dev_acc.class_pos.cpu().numpy().astype(np.float64)
        # batch by batch, so that ties between equal predictions keep the arrival order of the host path (the
        # epoch's positives and sums ride on the first batch: the calculators only add them up)
        r0 = 0
        for i, r in enumerate(rows):
            first = i == 0
            self.accumulate_stats({"n": r, "idx": idx[r0:r0 + r], "val": val[r0:r0 + r], "lab": lab[r0:r0 + r],
                                   "num_positives": pos if first else np.zeros_like(pos),
                                   "hit_sum": float(sums[1]) if first else 0.0,
                                   "perr_sum": float(sums[2]) if first else 0.0,
                                   "loss_sum": float(sums[3]) if first else 0.0}, local_only=True)
            r0 += r
        self._dev_triplets, self._dev_examples = [], 0
        self._dev_acc.sums.zero_()
        self._dev_acc.class_pos.zero_()

    def accumulate_stats(self, stats, local_only=False):
        """Fold the statistics of one batch (`batch_stats`) of this rank -- and, when distributed, of the same
        batch index of every other rank -- into the accumulators."""
        parts = [stats]
        if self.distributed and not local_only:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                parts = [None] * dist.get_world_size(self.group)
                dist.all_gather_object(parts, stats, group=self.group)
        n = hit = perr = loss = 0.0
        for st in parts:
            if st["n"] == 0:
                continue
            if self.vectorized:
                cls = np.asarray(st["idx"]).reshape(-1)
                val = np.asarray(st["val"], dtype=np.float64).reshape(-1)
                lab = np.asarray(st["lab"], dtype=np.float64).reshape(-1)
                by_class = np.argsort(cls, kind="stable")      # = flatten(top_k_by_class(...)) order
                self.map_calculator.accumulate(cls, val, lab, st["num_positives"])
                self.global_ap_calculator.accumulate(val[by_class], lab[by_class], float(np.sum(st["num_positives"])))
                n += st["n"]
                hit += st["hit_sum"]
                perr += st["perr_sum"]
                loss += st["loss_sum"]
                continue
            sparse_predictions = [[] for _ in range(self.num_class)]
            sparse_labels = [[] for _ in range(self.num_class)]
            for c, v, t in zip(st["idx"].reshape(-1), st["val"].reshape(-1), st["lab"].reshape(-1)):
                sparse_predictions[c].append(float(v))
                sparse_labels[c].append(float(t))
            num_positives = st["num_positives"].tolist()
            self.map_calculator.accumulate(sparse_predictions, sparse_labels, num_positives)
            self.global_ap_calculator.accumulate(flatten(sparse_predictions), flatten(sparse_labels),
                                                 sum(num_positives))
            n += st["n"]
            hit += st["hit_sum"]
            perr += st["perr_sum"]
            loss += st["loss_sum"]
        self.num_examples += int(n)
        self.sum_hit_at_one += hit
        self.sum_perr += perr
        self.sum_loss += loss
        if n == 0:
            return {"hit_at_one": 0.0, "perr": 0.0, "loss": 0.0}
        return {"hit_at_one": hit / n, "perr": perr / n, "loss": loss / n}

    def get(self):
        self._flush_device()
        if self.num_examples <= 0:
            raise ValueError("total_sample must be positive.")
        return {"avg_hit_at_one": self.sum_hit_at_one / self.num_examples,
                "avg_perr": self.sum_perr / self.num_examples,
                "avg_loss": self.sum_loss / self.num_examples,
                "aps": self.map_calculator.peek_map_at_n(),
                "gap": self.global_ap_calculator.peek_ap_at_n()}

    def clear(self):
        self.sum_hit_at_one = 0.0
        self.sum_perr = 0.0
        self.sum_loss = 0.0
        self.map_calculator.clear()
        self.global_ap_calculator.clear()
        self.num_examples = 0
        self._dev_triplets, self._dev_examples = [], 0
        if self._dev_acc is not None:
            self._dev_acc.sums.zero_()
            self._dev_acc.class_pos.zero_()


def format_lines(video_ids, predictions, top_k):
    """Prediction CSV lines "<video_id>,<class> <score> ... " with the top_k classes by descending score
    (inference_ensemble.py:63-74); the per-video selection runs on the GPU (evc_topk already returns
    the classes in that order)."""
    p = predictions if torch.is_tensor(predictions) else torch.as_tensor(np.asarray(predictions, dtype=np.float32))
    if not p.is_cuda:
        p = p.cuda()
    idx, val, _ = ops.topk(p.to(torch.float32).contiguous(), min(top_k, p.shape[1]))
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for video_index in range(len(video_ids)):
        vid = video_ids[video_index]
        vid = vid.decode("utf-8") if isinstance(vid, (bytes, bytearray)) else str(vid)
        yield vid + "," + " ".join("%i %f" % (int(c), float(s)) for c, s in zip(idx[video_index], val[video_index])) + "\n"
