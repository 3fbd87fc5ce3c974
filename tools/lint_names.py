"""Undefined-name check without third-party linters (none are installed in this image): every name a function or class body
reads as a global must be bound at module level or be a builtin.  `python tools/lint_names.py [files...]` exits 1 on a hit.
The GPU arms of bench.py cannot execute in a container without a GPU; this catches the NameError class of mistakes there."""
import ast
import builtins
import symtable
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def module_names(src):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__spec__", "__builtins__"}
    tree = ast.parse(src)
    for n in tree.body:
        for m in ast.walk(n) if not isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)) else [n]:
            if isinstance(m, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                names.add(m.name)
            elif isinstance(m, (ast.Import, ast.ImportFrom)):
                for al in m.names:
                    names.add((al.asname or al.name).split(".")[0])
            elif isinstance(m, ast.Name) and isinstance(m.ctx, ast.Store):
                names.add(m.id)
            elif isinstance(m, ast.ExceptHandler) and m.name:
                names.add(m.name)
    # `global x` assignments inside functions bind module names too
    for m in ast.walk(tree):
        if isinstance(m, ast.Global):
            names.update(m.names)
    return names


def check(path):
    src = Path(path).read_text()
    mod = module_names(src)
    star = any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(ast.parse(src)))
    bad = []

    def walk(tab):
        for child in tab.get_children():
            for s in child.get_symbols():
                if s.is_referenced() and s.is_global() and s.get_name() not in mod and not star:
                    bad.append((path, child.get_name(), child.get_lineno(), s.get_name()))
            walk(child)

    walk(symtable.symtable(src, str(path), "exec"))
    return bad


def main(argv):
    files = argv or [p for p in REPO.rglob("*.py") if not any(x in p.parts for x in ("_ref", "gpurun_out", ".git", "build"))]
    bad = []
    for f in files:
        bad += check(f)
    for path, scope, line, name in bad:
        print(f"{path}: in {scope} (line {line}): name {name!r} is not defined")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
