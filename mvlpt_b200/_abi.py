"""argtypes for every entry point, derived from include/mvlpt_sm100.h (the header is the single source of truth)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

HEADER = Path(__file__).resolve().parent.parent / "include" / "mvlpt_sm100.h"

_PROTO = re.compile(r"^(int|uint64_t|size_t|const char\*)\s+(mvlpt_\w+)\s*\(([^;{]*)\)\s*;", re.M | re.S)


def _ctype(decl: str):
    decl = decl.strip()
    if decl in ("void", ""):
        return None
    if "*" in decl or "mvlpt_stream_t" in decl:
        return ctypes.c_void_p
    base = decl.rsplit(" ", 1)[0].strip() if " " in decl else decl
    return {"int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
            "uint64_t": ctypes.c_uint64}[base]


def prototypes() -> dict:
    """{name: (restype_str, [arg decls])} for every function declared in the header."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in _PROTO.finditer(text):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        out[m.group(2)] = (m.group(1), [a for a in args if a and a != "void"])
    return out


def declare(L) -> None:
    for name, (ret, args) in prototypes().items():
        fn = getattr(L, name)
        fn.argtypes = [_ctype(a) for a in args]
        fn.restype = {"int": ctypes.c_int, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t,
                      "const char*": ctypes.c_char_p}[ret]
