"""Evaluation metrics of the reference's eval path (trainers/mvlpt.py:1046-1067 calls `self.dm._metric[task](y_true, y_pred)`
with the functions of trainers/vision_benchmark/datasets/metrics.py:1254-1294).  NumPy restatements of exactly the three
the ELEVATER tasks use — top-1 accuracy, mean-per-class accuracy and 11-point interpolated mAP — plus roc_auc through
scikit-learn when it is installed.  Host-side code (the reference computes them on the host too); pinned against the
reference's own functions by tests/golden/metrics.pt (oracle/gen_golden_metrics.py)."""
from __future__ import annotations

import numpy as np


def _targets_to_mat(targets: np.ndarray, n_class: int) -> np.ndarray:
    """metrics.py:122-130 — class indices (N,) -> one-hot (N, C); a matrix is returned as is."""
    targets = np.asarray(targets)
    if targets.ndim == 1:
        mat = np.zeros((len(targets), n_class), dtype=int)
        mat[np.arange(len(targets)), targets.astype(int)] = 1
        return mat
    return targets


def _drop_empty_classes(targets: np.ndarray, predictions: np.ndarray):
    """metrics.py:233-239 — classes without a single positive target are removed from targets AND predictions."""
    keep = np.where(~np.all(targets == 0, axis=0))[0]
    return targets[:, keep], predictions[:, keep]


def accuracy(y_label, y_pred) -> float:
    """metrics.py:1254-1262 (TopKAccuracyEvaluator(1), :256-290): share of samples whose arg-max class is the label."""
    y_label, y_pred = np.asarray(y_label), np.asarray(y_pred)
    assert y_label.ndim == 1 and len(y_label) == len(y_pred)
    if len(y_label) == 0:
        return 0.0
    return float(np.sum(np.argmax(y_pred, axis=1) == y_label)) / len(y_label)


def balanced_accuracy_score(y_label, y_pred) -> float:
    """metrics.py:1271-1274 (BalancedAccuracyScoreEvaluator, :839-850): mean over the classes that occur of their recall,
    arg-max taken over the occurring classes only."""
    y_pred = np.asarray(y_pred)
    if y_pred.size == 0:
        return 0.0
    tar, pred = _drop_empty_classes(_targets_to_mat(y_label, y_pred.shape[1]), y_pred)
    if tar.size == 0 or tar.shape[1] == 0:
        return 0.0
    t, p = np.argmax(tar, axis=1), np.argmax(pred, axis=1)
    recalls = [float(np.mean(p[t == c] == c)) for c in np.unique(t)]
    return float(np.mean(recalls))


def _precision_recall_curve(y_true: np.ndarray, score: np.ndarray):
    """sklearn.metrics.precision_recall_curve (the reference's `sm.precision_recall_curve`, metrics.py:872): one point per
    distinct score, recall decreasing, closed with (precision 1, recall 0)."""
    y_true = (np.asarray(y_true) == 1)  # pos_label = 1 for targets in {0,1} or {-1,1}
    order = np.argsort(score, kind="mergesort")[::-1]
    score, y_true = score[order], y_true[order]
    idx = np.r_[np.where(np.diff(score))[0], y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=ps != 0)
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    return np.hstack((precision[::-1], 1.0)), np.hstack((recall[::-1], 0.0))


def map_11_points(y_label, y_pred_proba, n_points: int = 11) -> float:
    """metrics.py:1265-1268 (MeanAveragePrecisionNPointsEvaluator, :853-895): per class, the maximum precision at recall
    >= 1.0, 0.9, .., 0.0 (walking the curve from high to low recall), averaged over thresholds, then over classes."""
    pred = np.asarray(y_pred_proba)
    if pred.size == 0:
        return 0.0
    tar, pred = _drop_empty_classes(_targets_to_mat(y_label, pred.shape[1]), pred)
    if tar.size == 0 or tar.shape[1] == 0:
        return 0.0
    thresholds = np.linspace(1, 0, n_points, endpoint=True).tolist()
    per_class = []
    for i in range(pred.shape[1]):
        precision, recall = _precision_recall_curve(tar[:, i], pred[:, i])
        interp, k, best = np.empty(len(thresholds)), 0, 0.0
        for j, th in enumerate(thresholds):
            while k < len(recall) and th <= recall[k]:
                best = max(best, precision[k])
                k += 1
            interp[j] = best
        per_class.append(np.mean(interp))
    return float(np.mean(per_class))


def roc_auc(y_true, y_score) -> float:
    """metrics.py:1276-1280 — scikit-learn's roc_auc_score, as in the reference."""
    from sklearn.metrics import roc_auc_score
    return float(roc_auc_score(y_true, y_score))


def get_metric(metric_name: str):
    """metrics.py:1283-1294."""
    table = {"accuracy": accuracy, "mean-per-class": balanced_accuracy_score, "11point_mAP": map_11_points, "roc_auc": roc_auc}
    if metric_name not in table:
        raise KeyError(f"Undefined metric {metric_name!r}")
    return table[metric_name]
