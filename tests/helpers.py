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
                                   project_dim=case.get("project_dim", 0) if upt else 0, seed=0,
                                   cocoop_n_ctx=case.get("cocoop_n_ctx", 0), vpt_project=case.get("vpt_project", -1))
    return fx, case, arch, sd, image, pp, upt


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """normwise-max relative error  max|a-b| / max|b|  (SURVEY.md App. A)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def make_cfg(case, prec="fp16"):
    """Attribute-style cfg with the keys the trainer mirror reads (same keys as the reference)."""
    from types import SimpleNamespace as NS
    res = synth.ARCHS[case["arch"]]["image_resolution"]
    return NS(
        TRAINER=NS(
            MVLPT=NS(PREC=prec, PROJECT_METHOD=case.get("project_method", "identity"),
                     PROJECT_DIM=case.get("project_dim", 128),
                     VPT=NS(N_CTX=case.get("vpt_n_ctx", 0), CTX_INIT="", DROPOUT=case.get("vpt_dropout", 0.0),
                            PROJECT=case.get("vpt_project", -1),
                            DEEP=case.get("vpt_deep", False)),
                     COOP=NS(N_CTX=case.get("coop_n_ctx", 0), CTX_INIT="", CSC=case.get("csc", False),
                             CLASS_TOKEN_POSITION=case.get("position", "end")),
                     COCOOP=NS(N_CTX=case.get("cocoop_n_ctx", 0), CTX_INIT="", PREC="fp32" if prec == "fp32" else "fp16")),
            CUT_CONTEXTLEN=case.get("cut", False), ACT_CKPT=1),
        INPUT=NS(SIZE=(res, res)),
        DATASET=NS(MULTITASK_LABEL_PERTASK=case.get("task_mask", False)),
        MODEL=NS(BACKBONE=NS(NAME=case["arch"])),
    )


def build_custom_clip(name: str, prec: str = "fp32", device="cuda"):
    """Our CustomCLIP for a golden case, with the fixture's tokenisation and the synthetic prompt parameters."""
    from types import SimpleNamespace as NS
    from mvlpt_b200.clip_model import build_model
    from mvlpt_b200.trainers.mvlpt import CustomCLIP
    fx, case, arch, sd, image, pp, upt = case_inputs(name)
    clip_model = build_model(sd)
    if prec != "fp16":
        clip_model.float()
    dm = None
    if case.get("task_mask"):
        sizes = case["tasks"]
        tn = [f"t{i}" for i in range(len(sizes))]
        dm = NS(_num_classes=case["C"], _task_names=tn, _labelmap={t: list(range(s)) for t, s in zip(tn, sizes)})
    model = CustomCLIP(make_cfg(case, prec), fx["names"], clip_model, dm=dm,
                       tokenized_prompts=fx["tokenized_prompts"], name_lens=fx["name_lens"])
    missing, unexpected = model.prompt_learner.load_state_dict(pp, strict=False)
    assert not unexpected, unexpected
    model = model.to(device)
    return model, fx, case, sd, image, pp, upt


def oracle_kwargs(fx, case, sd, upt):
    emb = sd["token_embedding.weight"][fx["tokenized_prompts"]]
    return dict(embedding=emb, eot_index=fx["eot_index"], name_lens=fx["name_lens"], n_ctx=case.get("coop_n_ctx", 0),
                v=case.get("vpt_n_ctx", 0), position=case.get("position", "end"), upt=upt, task=fx["task"],
                task_ranges=fx["task_ranges"], cocoop_n_ctx=case.get("cocoop_n_ctx", 0),
                drop_p=case.get("vpt_dropout", 0.0), drop_keep=fx.get("drop_keep"))
