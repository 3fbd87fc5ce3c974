"""Eval metrics against golden values produced by the reference's own functions
(oracle/gen_golden_metrics.py -> tests/golden/metrics.pt): accuracy, mean-per-class, 11-point mAP."""
from pathlib import Path

import numpy as np
import pytest
import torch

from mvlpt_b200.trainers import metrics as M

GOLD = Path(__file__).resolve().parent / "golden" / "metrics.pt"


@pytest.mark.parametrize("i", range(5))
def test_metrics_match_reference_golden(i):
    case = torch.load(GOLD)[i]
    pred, label = case["pred"].numpy(), case["label"].numpy()
    for name in ("accuracy", "mean-per-class", "11point_mAP"):
        if name in case:
            got = M.get_metric(name)(label, pred)
            assert abs(got - case[name]) < 1e-12, (case["kind"], name, got, case[name])


def test_metric_edge_cases():
    assert M.accuracy(np.zeros(0, dtype=int), np.zeros((0, 3))) == 0.0
    assert M.map_11_points(np.zeros(0, dtype=int), np.zeros((0, 3))) == 0.0
    # a perfect ranking has precision 1 at every recall level
    y = np.array([0, 1, 2, 1])
    p = np.eye(3)[y] * 5.0
    assert M.map_11_points(y, p) == 1.0 and M.balanced_accuracy_score(y, p) == 1.0 and M.accuracy(y, p) == 1.0
    with pytest.raises(KeyError):
        M.get_metric("nope")
