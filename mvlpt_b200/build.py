"""Build libmvlpt_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Usage:  python -m mvlpt_b200.build [--force] [--verbose]
The shared object lands next to this file so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libmvlpt_sm100.so"
OBJ = HERE / "csrc" / "_obj"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-DMVLPT_BUILDING_LIB",
] + os.environ.get("MVLPT_NVCC_EXTRA", "").split()  # e.g. -DMVLPT_FMHA_DBG for the in-kernel trace (tools/ only)


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "mvlpt_sm100.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    stamp = HERE / "csrc" / "_obj" / "stamp"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    OBJ.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        cmd = [NVCC, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, _sources()):
            if verbose or r.returncode:
                sys.stderr.write(f"--- {src.name}\n{r.stdout}{r.stderr}\n")
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src.name}")
            objs.append(str(obj))
    # exported symbols: everything declared extern "C" in include/mvlpt_sm100.h
    cmd = [NVCC, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
