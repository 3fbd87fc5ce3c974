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
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path
from types import SimpleNamespace as NS

import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("MVLPT_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))

from mvlpt_b200 import synth  # noqa: E402


def install_stubs():
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
    pkg.__path__ = [str(REF / "trainers")]
    vb = mod("trainers.vision_benchmark")
    vb.__path__ = []
    mod("trainers.vision_benchmark.evaluation", construct_dataloader=None, construct_multitask_dataset=None)
    mod("trainers.vision_benchmark.datasets", class_map_metric={}, get_metric=None)
    sys.path.insert(0, str(REF))


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
]


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
    names = NAMES[:C]
    dm = None
    task_ranges = None
    if case.get("task_mask"):
        sizes = case["tasks"]
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
