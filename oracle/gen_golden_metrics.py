"""Golden vectors for the eval metrics: seeded predictions / labels through the REFERENCE's own functions
(/root/reference/trainers/vision_benchmark/datasets/metrics.py:1254-1294, loaded by file path; it only needs numpy and
scikit-learn).  Run in the build container:  python oracle/gen_golden_metrics.py  ->  tests/golden/metrics.pt"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/trainers/vision_benchmark/datasets/metrics.py")
spec = importlib.util.spec_from_file_location("ref_metrics", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(0)
cases = []
for (N, C, kind) in [(200, 10, "single"), (57, 37, "single_missing"), (300, 20, "multi"), (64, 5, "ties"), (1, 3, "one")]:
    logits = rng.normal(size=(N, C)).astype(np.float32)
    if kind == "ties":
        logits = np.round(logits * 2) / 2  # many equal scores: exercises the distinct-threshold logic
    if kind == "multi":
        y = (rng.random((N, C)) < 0.15).astype(int)
        y[:, 3] = 0  # a class without positives
    else:
        hi = C if kind != "single_missing" else C - 5  # the last 5 classes never occur
        y = rng.integers(0, hi, size=N)
        logits[np.arange(N), y] += 1.0
    out = {"N": N, "C": C, "kind": kind, "pred": torch.from_numpy(logits), "label": torch.from_numpy(np.asarray(y))}
    if kind != "multi":
        out["accuracy"] = float(ref.accuracy(y, logits))
        out["mean-per-class"] = float(ref.balanced_accuracy_score(y, logits))
    out["11point_mAP"] = float(ref.map_11_points(y, logits))
    cases.append(out)
    print({k: v for k, v in out.items() if not torch.is_tensor(v)})
dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "metrics.pt"
torch.save(cases, dst)
print("wrote", dst)
