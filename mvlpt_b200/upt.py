"""UPT shared-prompt projection (trainers/mvlpt.py:376-414) — forward and backward incl. weight gradients."""
from __future__ import annotations


class UptProjection:
    def __init__(self, prompt_learner):
        raise NotImplementedError("PROJECT_METHOD='transformer' kernels are not built yet")
