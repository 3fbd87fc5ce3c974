"""Build recipe of `oracle/_ref/`: the UNMODIFIED reference hot path, byte-compiled where its sources lie.

    python oracle/build_ref.py            # /root/reference -> oracle/_ref/{clip,trainers}/*.refpyc (+ the BPE vocabulary)

The reference is pure Python (no native code, no build system), so "compiling" it means `py_compile` of
the five modules the path imports — clip/{__init__,clip,model,simple_tokenizer}.py and trainers/mvlpt.py — read from
/root/reference and written ONLY into oracle/_ref/ as sourceless byte-code files (`*.refpyc`), next to the tokenizer's vocabulary data
file it opens relative to its own location (clip/simple_tokenizer.py:10-12).  oracle/_ref/ is git-ignored (no reference
source enters the history) but not gpurun-ignored, so the compiled reference travels to the GPU box, where
/root/reference does not exist.  `__graft_entry__.build()` runs this when /root/reference is present.

Users: `bench.py --impl reference` (the reference's own CustomCLIP + F.cross_entropy + torch.optim.SGD step on the host
cores, and the same modules in fp16 PyTorch-eager on cuda:0 as the secondary baseline of SURVEY.md §8d) and
`oracle/gen_golden.py` (which imports the sources directly).  Test infrastructure / baseline only: nothing under
mvlpt_b200/ imports it.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("MVLPT_REFERENCE", "/root/reference"))

MODULES = ["clip/__init__.py", "clip/clip.py", "clip/model.py", "clip/simple_tokenizer.py", "trainers/mvlpt.py"]
DATA = ["clip/bpe_simple_vocab_16e6.txt.gz"]
# compiled modules carry this suffix instead of ".pyc": snapshot tools (gpurun among them) drop *.pyc files
SUFFIX = ".refpyc"


def build_ref(ref: Path = REF, out: Path = OUT) -> Path:
    if not ref.exists():
        raise FileNotFoundError(f"{ref} is absent (GPU box?): oracle/_ref/ is built in the build container only")
    for rel in MODULES:
        dst = out / Path(rel).with_suffix(SUFFIX)
        dst.parent.mkdir(parents=True, exist_ok=True)
        py_compile.compile(str(ref / rel), cfile=str(dst), dfile=f"<reference>/{rel}", doraise=True)
    for rel in DATA:
        dst = out / rel
        if not dst.exists() or dst.stat().st_size != (ref / rel).stat().st_size:
            shutil.copyfile(ref / rel, dst)
    (out / "BUILT_FROM").write_text(f"{ref} via oracle/build_ref.py (python {sys.version.split()[0]})\n")
    return out


def available(out: Path = OUT) -> bool:
    return all((out / Path(rel).with_suffix(SUFFIX)).exists() for rel in MODULES) and all((out / r).exists() for r in DATA)


def install_stubs(root: Path) -> None:
    """The in-memory stand-ins SURVEY.md §8c lists for what the reference imports and this image lacks (dassl, ftfy, the
    ELEVATER toolkit's heavy dependencies), then `root` (the reference tree, or oracle/_ref) on sys.path.  Nothing of the
    reference is modified: the stubs only satisfy its import statements; the hot path never calls them."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("ftfy", fix_text=lambda s: s)

    class _Registry:
        def register(self):
            return lambda cls: cls

    mod("dassl")
    mod("dassl.engine", TRAINER_REGISTRY=_Registry(), TrainerX=type("TrainerX", (), {}))
    mod("dassl.metrics", compute_accuracy=None)
    mod("dassl.utils", load_pretrained_weights=None, load_checkpoint=None)
    mod("dassl.optim", build_optimizer=None, build_lr_scheduler=None)
    mod("dassl.data", DataManager=type("DataManager", (), {}))
    mod("dassl.data.data_manager", build_data_loader=None)
    mod("dassl.data.datasets", build_dataset=None)
    mod("dassl.data.samplers", build_sampler=None)
    mod("dassl.data.transforms", INTERPOLATION_MODES=None, build_transform=None)
    pkg = mod("trainers")
    pkg.__path__ = [str(Path(root) / "trainers")]
    vb = mod("trainers.vision_benchmark")
    vb.__path__ = []
    mod("trainers.vision_benchmark.evaluation", construct_dataloader=None, construct_multitask_dataset=None)
    mod("trainers.vision_benchmark.datasets", class_map_metric={}, get_metric=None)
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))


def _load_compiled(name: str, path: Path, package_dir: Path | None = None):
    import importlib.machinery
    import importlib.util
    loader = importlib.machinery.SourcelessFileLoader(name, str(path))
    spec = importlib.util.spec_from_file_location(name, str(path), loader=loader,
                                                  submodule_search_locations=[str(package_dir)] if package_dir else None)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    loader.exec_module(m)
    return m


def import_reference(root: Path = OUT):
    """-> (clip.model module, trainers.mvlpt module) of the reference under `root`: the source tree (/root/reference), or
    oracle/_ref, whose sourceless modules are loaded explicitly in dependency order (they are not named *.pyc)."""
    root = Path(root)
    install_stubs(root)
    import importlib
    if (root / "clip" / "model.py").exists():
        return importlib.import_module("clip.model"), importlib.import_module("trainers.mvlpt")
    pkg = types.ModuleType("clip")
    pkg.__path__ = [str(root / "clip")]
    pkg.__file__ = str(root / "clip" / ("__init__" + SUFFIX))
    sys.modules["clip"] = pkg
    model = _load_compiled("clip.model", root / "clip" / ("model" + SUFFIX))
    pkg.model = model
    pkg.simple_tokenizer = _load_compiled("clip.simple_tokenizer", root / "clip" / ("simple_tokenizer" + SUFFIX))
    pkg.clip = _load_compiled("clip.clip", root / "clip" / ("clip" + SUFFIX))
    for k in getattr(pkg.clip, "__all__", []):  # clip/__init__.py: `from .clip import *`
        setattr(pkg, k, getattr(pkg.clip, k))
    mv = _load_compiled("trainers.mvlpt", root / "trainers" / ("mvlpt" + SUFFIX))
    return model, mv


if __name__ == "__main__":
    print("built", build_ref())
