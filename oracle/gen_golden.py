"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on the CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/<case>.pt

The reference cannot be imported as-is offline: `dassl`, `ftfy` and the ELEVATER toolkit's heavy dependencies
are absent.  They are stubbed IN MEMORY (sys.modules) exactly as SURVEY.md §8c describes; no reference source
is copied or modified.  Weights come from mvlpt_b200.synth (name-keyed deterministic generator) so the GPU
box can rebuild the same parameters without the reference.

Each fixture stores: the case config, tokenised prompts + name lengths (from the reference BPE tokenizer),
reference logits / loss / prompt gradients (fp32 run), and, for the "tiny" architecture, intermediate
activations (image features, text features, assembled prompts, projected prompts).

It also stores the reference's OWN fp16 run of the same step (`ref16_logits`, `ref16_loss`, `ref16_grads`: the model
converted with the reference's `convert_weights`, TRAINER.MVLPT.PREC="fp16", fp16 image, exactly the default
configuration of train.py:131).  Its distance from the fp32 run is the yard-stick of the parity tests: a CUDA path that
computes with fp16 operands is held to 1e-3 (logits) / 5e-3 (gradients) of the fp32 reference, or, where fp16 operand
rounding makes that unreachable, to never being further from it than the reference's own fp16 path is.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path
from types import SimpleNamespace as NS

import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("MVLPT_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))

from mvlpt_b200 import synth  # noqa: E402


def install_stubs():
    """In-memory stand-ins for dassl / ftfy / the ELEVATER toolkit (oracle/build_ref.py), reference tree on sys.path."""
    from oracle.build_ref import install_stubs as _stubs
    _stubs(REF)


def make_cfg(case) -> NS:
    """Attribute-style cfg with the keys trainers/mvlpt.py reads."""
    res = synth.ARCHS[case["arch"]]["image_resolution"]
    return NS(
        TRAINER=NS(
            MVLPT=NS(PREC="fp32", PROJECT_METHOD=case.get("project_method", "identity"),
                     PROJECT_DIM=case.get("project_dim", 128),
                     VPT=NS(N_CTX=case.get("vpt_n_ctx", 0), CTX_INIT="", DROPOUT=case.get("vpt_dropout", 0.0),
                            PROJECT=case.get("vpt_project", -1),
                            DEEP=case.get("vpt_deep", False)),
                     COOP=NS(N_CTX=case.get("coop_n_ctx", 0), CTX_INIT="", CSC=case.get("csc", False),
                             CLASS_TOKEN_POSITION=case.get("position", "end")),
                     COCOOP=NS(N_CTX=case.get("cocoop_n_ctx", 0), CTX_INIT="", PREC="fp32")),
            CUT_CONTEXTLEN=case.get("cut", False), ACT_CKPT=1),
        INPUT=NS(SIZE=(res, res)),
        DATASET=NS(MULTITASK_LABEL_PERTASK=case.get("task_mask", False)),
    )


NAMES = ["cat", "golden retriever", "airplane", "sea lion", "annual crop land", "2012 bmw m3 coupe",
         "face", "tiger shark", "red-winged blackbird", "pizza", "bell pepper", "oxeye daisy",
         "sun 397 abbey", "boeing 737-800", "highway or road", "knitting", "baby crawling", "dotted texture",
         "forest", "yin yang"]

CASES = [
    dict(name="tiny_coop_end", arch="tiny", coop_n_ctx=4, B=3, C=5),
    dict(name="tiny_coop_middle_cut", arch="tiny", coop_n_ctx=4, position="middle", cut=True, B=3, C=6),
    dict(name="tiny_coop_front_csc", arch="tiny", coop_n_ctx=4, position="front", csc=True, B=2, C=5),
    dict(name="tiny_vpt_shallow", arch="tiny", vpt_n_ctx=3, B=3, C=5),
    dict(name="tiny_vpt_deep", arch="tiny", vpt_n_ctx=4, vpt_deep=True, B=3, C=5),
    dict(name="tiny_vpt_deep_taskmask_soft", arch="tiny", vpt_n_ctx=4, vpt_deep=True, B=4, C=8, task_mask=True,
         tasks=[3, 3, 2], soft_labels=True),
    dict(name="tiny_upt_identity", arch="tiny", coop_n_ctx=4, vpt_n_ctx=4, vpt_deep=True, B=3, C=5),
    dict(name="tiny_upt_transformer", arch="tiny", coop_n_ctx=4, vpt_n_ctx=4, vpt_deep=True, position="middle",
         cut=True, project_method="transformer", project_dim=32, B=3, C=5),
    dict(name="tiny_cocoop", arch="tiny", cocoop_n_ctx=4, B=3, C=5),
    dict(name="tiny_cocoop_vpt_deep", arch="tiny", cocoop_n_ctx=4, vpt_n_ctx=3, vpt_deep=True, B=2, C=6),
    dict(name="tiny_vpt_deep_project", arch="tiny", vpt_n_ctx=4, vpt_deep=True, vpt_project=24, B=3, C=5),
    dict(name="tiny_vpt_shallow_project_coop", arch="tiny", vpt_n_ctx=3, vpt_project=16, coop_n_ctx=4, B=2, C=5),
    dict(name="tiny_vpt_deep_dropout", arch="tiny", vpt_n_ctx=4, vpt_deep=True, vpt_dropout=0.25, B=3, C=5),
    dict(name="tiny_vpt_shallow_project_dropout", arch="tiny", vpt_n_ctx=3, vpt_project=16, vpt_dropout=0.5, B=4, C=5),
    dict(name="b16_coop_end", arch="ViT-B/16", coop_n_ctx=16, B=2, C=10),
    dict(name="b16_vpt_deep", arch="ViT-B/16", vpt_n_ctx=8, vpt_deep=True, B=2, C=10),
    dict(name="b16_upt_transformer", arch="ViT-B/16", coop_n_ctx=16, vpt_n_ctx=8, vpt_deep=True, position="middle",
         cut=True, project_method="transformer", project_dim=128, B=2, C=10),
    dict(name="b32_coop_cfg1", arch="ViT-B/32", coop_n_ctx=4, B=1, C=20),
    dict(name="b16_cocoop", arch="ViT-B/16", cocoop_n_ctx=4, B=2, C=4),
    dict(name="b16_vpt_deep_project", arch="ViT-B/16", vpt_n_ctx=8, vpt_deep=True, vpt_project=256, B=2, C=6),
    dict(name="l14_vpt_deep", arch="ViT-L/14", vpt_n_ctx=8, vpt_deep=True, B=1, C=4),
    # configs/trainers/MVLPT/vit_l14_336.yaml: 577 image tokens + prompts (the streaming attention kernels)
    dict(name="l14_336_vpt_deep", arch="ViT-L/14@336px", vpt_n_ctx=4, vpt_deep=True, B=1, C=3),
    dict(name="l14_coop_end", arch="ViT-L/14", coop_n_ctx=16, B=1, C=4),
    # BASELINE class counts.  configs[3]/[4] label space: the 1000 ImageNet names of scripts/classnames.txt, the scripts'
    # 'middle' position and context-length cut (scripts/mvlpt/main_mt_coopdata_cut.sh:41-47)
    dict(name="b16_coop_c1000_cut", arch="ViT-B/16", coop_n_ctx=16, position="middle", cut=True, B=2, C=1000,
         names_from="imagenet"),
    # configs[2]: the 11 CoOp-source tasks of scripts/mvlpt/main_mt_coopdata_cut.sh:21 (C = 2193), VPT-deep, one task
    # per sample, per-task logit mask (trainers/mvlpt.py:527-538,573-581), multi-hot soft labels (:914-916)
    dict(name="b16_vpt_deep_11task", arch="ViT-B/16", vpt_n_ctx=8, vpt_deep=True, B=4, C=2193, task_mask=True,
         tasks="11task", soft_labels=True, names_from="11task"),
]

# the 11 source tasks (scripts/mvlpt/main_mt_coopdata_cut.sh:21) under their ELEVATER names
# (trainers/vision_benchmark/datasets/prompts.py:3221-3247)
TASKS_11 = ["imagenet-1k", "caltech-101", "food-101", "stanford-cars", "oxford-iiit-pets", "oxford-flower-102",
            "fgvc-aircraft-2013b-variants102", "sun397", "dtd", "eurosat_clip", "ucf101"]


def label_space(case):
    """(class names, per-task sizes | None) of a case."""
    src = case.get("names_from")
    if src is None:
        return NAMES[:case["C"]], case.get("tasks")
    if src == "imagenet":
        lines = (REF / "scripts" / "classnames.txt").read_text().splitlines()
        names = [ln.split(" ", 1)[1] for ln in lines if ln.strip()]
        assert len(names) == case["C"]
        return names, None
    if src == "11task":
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "_ref_prompts", REF / "trainers" / "vision_benchmark" / "datasets" / "prompts.py")
        pm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(pm)
        names, sizes = [], []
        for t in TASKS_11:
            names += list(pm.class_map[t])
            sizes.append(len(pm.class_map[t]))
        assert len(names) == case["C"], len(names)
        return names, sizes
    raise ValueError(src)



def run_ref16(case, names, dm, pp, image, label_in, task):
    """The same step through the reference's default fp16 path (train.py:131 PREC="fp16"): weights converted by the
    reference's own `convert_weights` (clip/model.py:371-392), prompt parameters created in CLIP's dtype, fp16 image,
    F.cross_entropy on the fp16 logits, autograd backward.  CPU execution of the same modules."""
    from clip.model import CLIP, convert_weights
    import trainers.mvlpt as ref
    if case.get("vpt_dropout"):
        return {}  # the Bernoulli draws of a second run differ: no like-for-like fp16 yard-stick
    torch.manual_seed(1234)
    clip16 = CLIP(**synth.ARCHS[case["arch"]])
    clip16.load_state_dict(synth.synth_clip_state_dict(case["arch"], seed=0))
    clip16.eval()
    convert_weights(clip16)
    cfg = make_cfg(case)
    cfg.TRAINER.MVLPT.PREC = "fp16"
    cfg.TRAINER.MVLPT.COCOOP.PREC = "fp16"
    model = ref.CustomCLIP(cfg, names, clip16, dm=dm)
    pl = model.prompt_learner
    missing, unexpected = pl.load_state_dict(pp, strict=False)
    assert not unexpected, unexpected
    for n, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in n)
    try:
        logits = model(image.half(), task=task)
        lab = label_in
        if lab.dim() > 1:
            lab = lab.float()
            lab = (lab / lab.sum(dim=-1, keepdim=True)).to(logits.dtype)
        loss = torch.nn.functional.cross_entropy(logits, lab)
        loss.backward()
    except RuntimeError as e:  # an op without a CPU half kernel in this torch build
        print(f"  ref16 run unavailable for {case['name']}: {e}")
        return {}
    grads = {n: p.grad.detach().float().clone() for n, p in pl.named_parameters() if p.grad is not None}
    return dict(ref16_logits=logits.detach().float().clone(), ref16_loss=loss.detach().float().clone(), ref16_grads=grads)


def run_case(case, out_dir: Path):
    from clip.model import CLIP
    import trainers.mvlpt as ref

    torch.manual_seed(1234)
    arch = synth.ARCHS[case["arch"]]
    sd = synth.synth_clip_state_dict(case["arch"], seed=0)
    clip_model = CLIP(**arch)
    clip_model.load_state_dict(sd)
    clip_model.eval().float()

    C, B = case["C"], case["B"]
    names, sizes = label_space(case)
    case = dict(case)
    if sizes is not None:
        case["tasks"] = list(sizes)
    dm = None
    task_ranges = None
    if case.get("task_mask"):
        assert sum(sizes) == C
        tnames = [f"t{i}" for i in range(len(sizes))]
        labelmap = {t: list(range(s)) for t, s in zip(tnames, sizes)}
        # the reference sizes its per-task tables with dm._num_classes but indexes them by TASK id
        dm = NS(_num_classes=C, _task_names=tnames, _labelmap=labelmap)
        st, rng = 0, []
        for s in sizes:
            rng.append((st, st + s))
            st += s
        task_ranges = torch.tensor(rng)
    cfg = make_cfg(case)
    model = ref.CustomCLIP(cfg, names, clip_model, dm=dm)
    pl = model.prompt_learner
    pp = synth.synth_prompt_params(case["arch"], case.get("coop_n_ctx", 0), case.get("vpt_n_ctx", 0),
                                   case.get("vpt_deep", False), csc_classes=C if case.get("csc") else 0,
                                   project_dim=case.get("project_dim", 0) if case.get("project_method") == "transformer" else 0,
                                   seed=0, cocoop_n_ctx=case.get("cocoop_n_ctx", 0),
                                   vpt_project=case.get("vpt_project", -1))
    missing, unexpected = pl.load_state_dict(pp, strict=False)
    assert not unexpected, unexpected
    assert all(k in ("token_prefix", "token_suffix") for k in missing), missing
    for n, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in n)

    image = synth.synth_images(B, arch["image_resolution"], seed=1)
    g = torch.Generator().manual_seed(99)
    task = None
    if case.get("task_mask"):
        task = torch.randint(0, len(case["tasks"]), (B,), generator=g)
        label = torch.stack([torch.randint(int(task_ranges[t, 0]), int(task_ranges[t, 1]), (1,), generator=g)[0]
                             for t in task])
    else:
        label = torch.randint(0, C, (B,), generator=g)
    if case.get("soft_labels"):
        onehot = torch.zeros(B, C)
        onehot[torch.arange(B), label] = 1.0
        onehot[0, int(task_ranges[task[0], 0])] = 1.0  # one multi-hot row
        label_in = onehot
    else:
        label_in = label

    # vpt_dropout (training mode): record the Bernoulli draws of the reference's own nn.Dropout, in call order
    # (forward_vpt first, then one per deep layer), so that the oracle can be pinned with the masks given
    drop_keep = []
    if case.get("vpt_dropout"):
        assert pl.vpt_dropout.training
        pl.vpt_dropout.register_forward_hook(lambda m, inp, out: drop_keep.append((out != 0).detach().clone()))
    logits = model(image, task=task)
    lab = label_in
    if lab.dim() > 1:
        lab = lab.float()
        lab = lab / lab.sum(dim=-1, keepdim=True)
    loss = torch.nn.functional.cross_entropy(logits, lab)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in pl.named_parameters() if p.grad is not None}

    tok = pl.tokenized_prompts
    fix = dict(
        case=case, names=names, tokenized_prompts=tok.clone(), name_lens=list(pl.name_lens),
        eot_index=tok.argmax(dim=-1), label=label_in, task=task, task_ranges=task_ranges,
        logits=logits.detach().clone(), loss=loss.detach().clone(), grads=grads,
        weights_fingerprint=float(sum(v.double().abs().sum() for v in sd.values())),
        image_fingerprint=float(image.double().abs().sum()),
        torch_version=torch.__version__,
    )
    if drop_keep:
        fix["drop_keep"] = drop_keep
    if case["arch"] == "tiny" and not case.get("vpt_dropout"):
        with torch.no_grad():
            ctx, vpt, vpt_deep = pl.forward_mvlpt_proj(torch.float32)
            fix["proj_ctx"], fix["proj_vpt"], fix["proj_vpt_deep"] = ctx, vpt, vpt_deep
            fix["image_features"] = model.image_encoder(image, vpt, vpt_deep)
            if not case.get("cocoop_n_ctx"):
                prompts = pl.forward_coop(ctx)
                fix["prompts"] = prompts
                fix["text_features"] = model.text_encoder(prompts, tok)
            else:
                imf = fix["image_features"] / fix["image_features"].norm(dim=-1, keepdim=True)
                fix["cocoop_prompts"] = pl.forward_cocoop(imf)
    fix.update(run_ref16(case, names, dm, pp, image, label_in, task))
    # margins for the argmax check
    top2 = logits.detach().topk(2, dim=-1).values
    fix["top2_margin"] = (top2[:, 0] - top2[:, 1]).clone()
    torch.save(fix, out_dir / f"{case['name']}.pt")
    print(f"{case['name']}: loss={float(loss):.6f} logits[0,:3]={logits[0, :3].tolist()} "
          f"grads={ {k: float(v.abs().sum()) for k, v in grads.items()} }")


def main():
    install_stubs()
    out_dir = REPO / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])  # python oracle/gen_golden.py [case names]: regenerate only those
    for case in CASES:
        if only and case["name"] not in only:
            continue
        run_case(case, out_dir)


if __name__ == "__main__":
    main()
