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
