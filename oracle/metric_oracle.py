"""CPU restatement of the metric half of the path (utils/metric.py).  TEST INFRASTRUCTURE.

The reference calls scikit-learn (un-vendored third party; installed 1.9.0, the
reference pins no version) and NumPy 2.3.5.  The algorithms are restated here in
plain NumPy + small Python loops, following sklearn/metrics/_ranking.py
(``_sort_inputs_and_compute_classification_thresholds`` 878-932,
``confusion_matrix_at_thresholds`` 934-1047, ``roc_curve`` 1317-1372,
``precision_recall_curve`` 1160-1205, ``auc`` 95-115) and numpy's pairwise
``add.reduce``.  Pinned bit-for-bit against outputs of the reference's own wrappers
(``get_auc_roc`` ...) stored in ``tests/golden/metrics_golden.json``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


# ----------------------------------------------------------------------------
# numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum)
# ----------------------------------------------------------------------------
def pairwise_sum(a: np.ndarray) -> float:
    """Order-exact restatement of ``np.add.reduce`` over a contiguous 1-D float array
    (fp64 or fp32; the dtype of ``a`` is the accumulation dtype)."""
    a = np.ascontiguousarray(a)
    n = a.shape[0]
    dt = a.dtype.type
    if n < 8:
        res = dt(0.0)
        for i in range(n):
            res = dt(res + a[i])
        return res
    if n <= 128:
        r = a[0:8].copy()
        m = n - (n % 8)
        for i in range(8, m, 8):
            r = r + a[i:i + 8]
        res = dt(dt(dt(r[0] + r[1]) + dt(r[2] + r[3])) + dt(dt(r[4] + r[5]) + dt(r[6] + r[7])))
        for i in range(m, n):
            res = dt(res + a[i])
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return dt(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def trapezoid(y: np.ndarray, x: np.ndarray) -> float:
    """np.trapezoid: sum(d * (y[1:] + y[:-1]) / 2.0) with numpy's reduction order."""
    d = np.diff(x)
    terms = d * (y[1:] + y[:-1]) / 2.0
    return float(pairwise_sum(terms))


# ----------------------------------------------------------------------------
# sklearn curves
# ----------------------------------------------------------------------------
def _thresholds(label: np.ndarray, score: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fps, tps (fp64) at each distinct-score index, scores sorted descending (stable)."""
    score = np.asarray(score).ravel()
    y = (np.asarray(label).ravel() == 1).astype(np.float64)
    if not (np.all(np.isfinite(score))):
        raise ValueError("Input contains NaN or infinity")
    n = score.shape[0]
    order = np.argsort(-score, kind="stable") if score.dtype.kind == "f" else np.argsort(score, kind="stable")[::-1]
    s = score[order]
    y = y[order]
    idx = np.concatenate([np.nonzero(np.diff(s))[0], [n - 1]])
    tps = np.cumsum(y, dtype=np.float64)[idx]
    fps = 1 + idx.astype(np.float64) - tps
    return fps, tps


def roc_auc(score: np.ndarray, label: np.ndarray) -> float:
    """utils/metric.py:29-44 ``get_auc_roc`` = roc_curve(drop_intermediate=True) + auc;
    any exception -> 0.0 (bare except)."""
    try:
        fps, tps = _thresholds(label, score)
        if fps.shape[0] > 2:
            keep = np.concatenate([[True], np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), [True]])
            fps, tps = fps[keep], tps[keep]
        tps = np.concatenate([[0.0], tps])
        fps = np.concatenate([[0.0], fps])
        fpr = np.full(fps.shape, np.nan) if fps[-1] <= 0 else fps / fps[-1]
        tpr = np.full(tps.shape, np.nan) if tps[-1] <= 0 else tps / tps[-1]
        return _auc(fpr, tpr)
    except Exception:
        return .0


def pr_auc(score: np.ndarray, label: np.ndarray) -> float:
    """utils/metric.py:97-116 ``get_auc_prc`` = precision_recall_curve (drop_intermediate
    False) + auc(recalls, precisions); any exception -> 0.0."""
    try:
        fps, tps = _thresholds(label, score)
        ps = tps + fps
        precision = np.where(ps != 0, np.divide(tps, np.where(ps != 0, ps, 1.0)), 0.0)
        recall = np.full(tps.shape, 1.0) if tps[-1] == 0 else tps / tps[-1]
        precision = np.concatenate([precision[::-1], [1.0]])
        recall = np.concatenate([recall[::-1], [0.0]])
        return _auc(recall, precision)
    except Exception:
        return .0


def _auc(x: np.ndarray, y: np.ndarray) -> float:
    if x.shape[0] < 2:
        raise ValueError("At least 2 points are needed")
    direction = 1
    dx = np.diff(x)
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1
        else:
            raise ValueError("x is neither increasing nor decreasing")
    return float(direction * trapezoid(y, x))


# ----------------------------------------------------------------------------
# threshold metrics
# ----------------------------------------------------------------------------
def quantile_f32(valid_score: np.ndarray, q: float = 0.90) -> np.float32:
    """np.quantile(fp32 array, python float) under NumPy 2.x: all arithmetic in fp32,
    linear interpolation with the lerp that switches formula at g >= 0.5
    (numpy/lib/_function_base_impl.py ``_lerp``)."""
    s = np.sort(np.asarray(valid_score, dtype=np.float32))
    n = s.shape[0]
    vi = np.float32(n - 1) * np.float32(q)
    lo = int(math.floor(float(vi)))
    g = np.float32(vi - np.float32(lo))
    a = s[lo]
    b = s[min(lo + 1, n - 1)]
    diff = np.float32(b - a)
    if g >= np.float32(0.5):
        return np.float32(b - np.float32(diff * np.float32(np.float32(1) - g)))
    return np.float32(a + np.float32(diff * g))


def f1_score(valid_score, test_score, test_label) -> Tuple[float, np.float32]:
    """utils/metric.py:118-130 -- threshold = q0.90 of valid scores (argument overwritten,
    line 120); predictions = test > thr; p, r, f1 (0/0 -> nan like numpy)."""
    thr = quantile_f32(valid_score, 0.90)
    pred = np.asarray(test_score) > thr
    lab = np.asarray(test_label).astype(bool)
    tp = float((pred & lab).sum())
    with np.errstate(all="ignore"):
        p = np.float64(tp) / np.float64(pred.sum())
        r = np.float64(tp) / np.float64(lab.sum())
        f1 = p * r * 2 / (p + r)
    return float(f1), thr


def confusion_precision_recall(score, test_label, threshold) -> Tuple[float, float]:
    """utils/metric.py:83-95 -- pred = score >= thr (note >=), tn/fp/fn/tp, p and r."""
    pred = np.asarray(score) >= threshold
    lab = np.asarray(test_label).astype(bool)
    tp = int((pred & lab).sum())
    fp = int((pred & ~lab).sum())
    fn = int((~pred & lab).sum())
    with np.errstate(all="ignore"):
        precision = np.float64(tp) / np.float64(tp + fp)
        recall = np.float64(tp) / np.float64(tp + fn)
    return float(precision), float(recall)
