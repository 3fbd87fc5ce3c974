"""Shared helpers: rebuild a golden case's inputs from the fixture + the deterministic synth generator."""
from __future__ import annotations

import functools

import torch

from mvlpt_b200 import synth
from tests.conftest import load_golden


@functools.lru_cache(maxsize=4)
def clip_sd(arch: str):
    return synth.synth_clip_state_dict(arch, seed=0)


def case_inputs(name: str):
    fx = load_golden(name)
    case = fx["case"]
    arch = synth.ARCHS[case["arch"]]
    sd = clip_sd(case["arch"])
    fp = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(fp - fx["weights_fingerprint"]) <= 1e-9 * fx["weights_fingerprint"], "synthetic weights drifted"
    image = synth.synth_images(case["B"], arch["image_resolution"], seed=1)
    assert abs(float(image.double().abs().sum()) - fx["image_fingerprint"]) <= 1e-9 * fx["image_fingerprint"]
    upt = case.get("project_method") == "transformer"
    pp = synth.synth_prompt_params(case["arch"], case.get("coop_n_ctx", 0), case.get("vpt_n_ctx", 0),
                                   case.get("vpt_deep", False), csc_classes=case["C"] if case.get("csc") else 0,
                                   project_dim=case.get("project_dim", 0) if upt else 0, seed=0)
    return fx, case, arch, sd, image, pp, upt


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """normwise-max relative error  max|a-b| / max|b|  (SURVEY.md App. A)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
