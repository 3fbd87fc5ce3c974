"""ctypes struct/argtype declarations mirroring include/mvlpt_sm100.h (everything except mvlpt_gemm)."""
from __future__ import annotations


def declare(L) -> None:
    pass
