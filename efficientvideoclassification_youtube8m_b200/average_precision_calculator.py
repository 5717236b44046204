"""Average precision over accumulated (prediction, label) pairs
(code_student_uniform/average_precision_calculator.py:48-261 semantics, numpy arrays instead of
a python heap).  Ties between equal predictions: the reference shuffles with random.seed(0)
(:234-240), whose stream differs between Python 2 and 3; here ties keep arrival order."""
from __future__ import annotations

import numbers

import numpy as np


class AveragePrecisionCalculator(object):
    def __init__(self, top_n=None):
        if not ((isinstance(top_n, int) and top_n >= 0) or top_n is None):
            raise ValueError("top_n must be a positive integer or None.")
        self._top_n = top_n
        self._total_positives = 0
        self._pred, self._act = [], []

    @property
    def heap_size(self):
        return int(sum(len(p) for p in self._pred))

    @property
    def num_accumulated_positives(self):
        return self._total_positives

    def accumulate(self, predictions, actuals, num_positives=None):
        predictions, actuals = np.asarray(predictions, dtype=np.float64), np.asarray(actuals, dtype=np.float64)
        if len(predictions) != len(actuals):
            raise ValueError("the shape of predictions and actuals does not match.")
        if num_positives is not None:
            if not isinstance(num_positives, numbers.Number) or num_positives < 0:
                raise ValueError("'num_positives' was provided but it wan't a nonzero number.")
            self._total_positives += num_positives
        else:
            self._total_positives += int(np.sum(actuals > 0))
        self._pred.append(predictions)
        self._act.append(actuals)
        if self._top_n is not None and self.heap_size > 4 * max(self._top_n, 1024):
            self._compact()

    def _compact(self):
        p, a = np.concatenate(self._pred), np.concatenate(self._act)
        if self._top_n is not None and len(p) > self._top_n:
            keep = np.argsort(-p, kind="stable")[: self._top_n]
            keep.sort()
            p, a = p[keep], a[keep]
        self._pred, self._act = [p], [a]

    def clear(self):
        self._pred, self._act, self._total_positives = [], [], 0

    def peek_ap_at_n(self):
        if self.heap_size <= 0:
            return 0
        self._compact()
        return self.ap_at_n(self._pred[0], self._act[0], n=self._top_n, total_num_positives=self._total_positives)

    @staticmethod
    def ap(predictions, actuals):
        return AveragePrecisionCalculator.ap_at_n(predictions, actuals, n=None)

    @staticmethod
    def ap_at_n(predictions, actuals, n=20, total_num_positives=None):
        if len(predictions) != len(actuals):
            raise ValueError("the shape of predictions and actuals does not match.")
        if n is not None and (not isinstance(n, int) or n <= 0):
            raise ValueError("n must be 'None' or a positive integer. It was '%s'." % n)
        p, a = np.asarray(predictions, dtype=np.float64), np.asarray(actuals, dtype=np.float64)
        order = np.argsort(-p, kind="stable")
        numpos = float(np.sum(a > 0)) if total_num_positives is None else float(total_num_positives)
        if numpos == 0:
            return 0
        if n is not None:
            numpos = min(numpos, n)
        r = len(order) if n is None else min(len(order), n)
        hits = a[order[:r]] > 0
        poscount = np.cumsum(hits)
        prec = poscount / np.arange(1, r + 1)
        return float(np.sum(prec[hits]) / numpos)


class MeanAveragePrecisionCalculator(object):
    """code_student_uniform/mean_average_precision_calculator.py:31-99."""

    def __init__(self, num_class):
        if not isinstance(num_class, int) or num_class <= 1:
            raise ValueError("num_class must be a positive integer.")
        self._ap_calculators = [AveragePrecisionCalculator() for _ in range(num_class)]
        self._num_class = num_class

    def accumulate(self, predictions, actuals, num_positives=None):
        if not num_positives:
            num_positives = [None for _ in predictions]
        for i in range(len(predictions)):
            if len(predictions[i]) or num_positives[i]:
                self._ap_calculators[i].accumulate(predictions[i], actuals[i], num_positives[i])

    def clear(self):
        for c in self._ap_calculators:
            c.clear()

    def is_empty(self):
        return ([c.heap_size for c in self._ap_calculators] == [0 for _ in range(self._num_class)])

    def peek_map_at_n(self):
        return [c.peek_ap_at_n() for c in self._ap_calculators]


class SparseMeanAveragePrecisionCalculator(object):
    """MeanAveragePrecisionCalculator fed with sparse (class, prediction, label) triplets instead of one python
    list per class: the same per-class AP (same stable tie order: arrival order inside a class), computed
    with one sort over all triplets instead of a python loop over 4716 calculators per batch."""

    def __init__(self, num_class):
        if not isinstance(num_class, int) or num_class <= 1:
            raise ValueError("num_class must be a positive integer.")
        self._num_class = num_class
        self.clear()

    def clear(self):
        self._cls, self._pred, self._act = [], [], []
        self._total_positives = np.zeros(self._num_class, dtype=np.float64)

    def is_empty(self):
        return sum(len(c) for c in self._cls) == 0

    def accumulate(self, classes, predictions, actuals, num_positives):
        """classes/predictions/actuals: flat arrays of equal length (arrival order); num_positives [num_class]."""
        classes = np.asarray(classes).reshape(-1).astype(np.int64)
        predictions = np.asarray(predictions, dtype=np.float64).reshape(-1)
        actuals = np.asarray(actuals, dtype=np.float64).reshape(-1)
        if not (len(classes) == len(predictions) == len(actuals)):
            raise ValueError("the shape of predictions and actuals does not match.")
        num_positives = np.asarray(num_positives, dtype=np.float64).reshape(-1)
        if len(num_positives) != self._num_class or np.any(num_positives < 0):
            raise ValueError("'num_positives' must hold one non-negative count per class.")
        self._cls.append(classes)
        self._pred.append(predictions)
        self._act.append(actuals)
        self._total_positives += num_positives

    def peek_map_at_n(self):
        aps = np.zeros(self._num_class, dtype=np.float64)
        if self.is_empty():
            return aps.tolist()
        c, p, a = np.concatenate(self._cls), np.concatenate(self._pred), np.concatenate(self._act)
        self._cls, self._pred, self._act = [c], [p], [a]
        order = np.lexsort((-p, c))                      # by class, then prediction descending; stable
        c, hits = c[order], a[order] > 0
        start = np.flatnonzero(np.r_[True, c[1:] != c[:-1]])          # first triplet of every class present
        seg = np.cumsum(np.r_[True, c[1:] != c[:-1]]) - 1
        rank = np.arange(len(c)) - start[seg] + 1
        cum = np.cumsum(hits)
        poscount = cum - (cum[start] - hits[start])[seg]              # positives so far inside the class
        contrib = np.where(hits, poscount / rank, 0.0)
        sums = np.bincount(c, weights=contrib, minlength=self._num_class)
        ok = self._total_positives > 0
        aps[ok] = sums[ok] / self._total_positives[ok]
        return aps.tolist()
