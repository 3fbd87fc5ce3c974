"""ctypes binding of libmvlpt_sm100.so (the C ABI declared in include/mvlpt_sm100.h).

There is no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_uint64, c_void_p
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmvlpt_sm100.so"


class MvlptError(RuntimeError):
    pass


class GemmDesc(Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int),
                ("lda", c_int), ("ldw", c_int), ("ld_out", c_int), ("ld_aux", c_int),
                ("act", c_int), ("out_f32", c_int), ("alpha", c_float)]


class LnCarry(Structure):
    """mvlpt_ln_carry of include/mvlpt_sm100.h."""
    _fields_ = [("rec_in", c_void_p), ("rec_out", c_void_p), ("gamma", c_void_p), ("xt", c_void_p),
                ("rec", c_void_p), ("sg", c_void_p), ("width", c_int), ("eps", c_float)]


LN_REC = 20  # MVLPT_LN_REC

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("MVLPT_AUTOBUILD", "0") == "1":
            from .build import build_lib
            build_lib()
        else:
            raise MvlptError(
                f"{LIB_PATH} is missing: build it with `python -m mvlpt_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
    L = ctypes.CDLL(str(LIB_PATH))
    _declare_ops(L)
    _lib = L
    return L


def _declare_ops(L):
    """argtypes/restype for every entry point, parsed from include/mvlpt_sm100.h."""
    from . import _abi
    _abi.declare(L)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mvlpt_last_error().decode(errors="replace")
        raise MvlptError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().mvlpt_launch_count())
