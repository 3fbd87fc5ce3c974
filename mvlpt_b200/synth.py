"""Deterministic synthetic CLIP weights, prompts and batches.

No pretrained CLIP checkpoint can be fetched offline (clip/clip.py:29-70 needs the network), so parity and
benchmarks run on random weights of the right architecture.  The generator is keyed per tensor NAME
(not by construction order), so the same state dict can be loaded into the reference's `clip.model.CLIP`
(oracle/gen_golden.py, in the build container) and into this package (on the GPU box) and the golden
outputs committed under tests/golden/ stay valid.

Key names and shapes follow the reference state dict (clip/model.py:202-217,239-293); magnitudes follow its
initialiser (clip/model.py:295-322).  Tensors the reference keeps in fp16 (`convert_weights`,
clip/model.py:371-392) are rounded to fp16-representable values so fp32 oracle and fp16 kernels see
identical parameters.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

# constructor arguments of the reference's CLIP(...) (clip/model.py:240-253) for the ViT backbones
ARCHS = {
    "ViT-B/16": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                     context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                     transformer_layers=12),
    "ViT-B/32": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32,
                     context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                     transformer_layers=12),
    "ViT-L/14": dict(embed_dim=768, image_resolution=224, vision_layers=24, vision_width=1024, vision_patch_size=14,
                     context_length=77, vocab_size=49408, transformer_width=768, transformer_heads=12,
                     transformer_layers=12),
    "ViT-L/14@336px": dict(embed_dim=768, image_resolution=336, vision_layers=24, vision_width=1024, vision_patch_size=14,
                           context_length=77, vocab_size=49408, transformer_width=768, transformer_heads=12,
                           transformer_layers=12),
    # small shapes for golden fixtures that carry their own activations (heads are still 64 wide)
    "tiny": dict(embed_dim=64, image_resolution=64, vision_layers=3, vision_width=128, vision_patch_size=16,
                 context_length=77, vocab_size=49408, transformer_width=128, transformer_heads=2,
                 transformer_layers=2),
}


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, seed, shape, std, mean=0.0, fp16=False):
    t = torch.randn(*shape, generator=_gen(name, seed), dtype=torch.float32) * std + mean
    if fp16:
        t = t.half().float()
    return t


def _block(sd, prefix, width, seed, attn_std, proj_std, fc_std):
    sd[prefix + "attn.in_proj_weight"] = _normal(prefix + "attn.in_proj_weight", seed, (3 * width, width), attn_std, fp16=True)
    sd[prefix + "attn.in_proj_bias"] = _normal(prefix + "attn.in_proj_bias", seed, (3 * width,), 0.02, fp16=True)
    sd[prefix + "attn.out_proj.weight"] = _normal(prefix + "attn.out_proj.weight", seed, (width, width), proj_std, fp16=True)
    sd[prefix + "attn.out_proj.bias"] = _normal(prefix + "attn.out_proj.bias", seed, (width,), 0.02, fp16=True)
    sd[prefix + "ln_1.weight"] = _normal(prefix + "ln_1.weight", seed, (width,), 0.05, mean=1.0)
    sd[prefix + "ln_1.bias"] = _normal(prefix + "ln_1.bias", seed, (width,), 0.02)
    sd[prefix + "mlp.c_fc.weight"] = _normal(prefix + "mlp.c_fc.weight", seed, (4 * width, width), fc_std, fp16=True)
    sd[prefix + "mlp.c_fc.bias"] = _normal(prefix + "mlp.c_fc.bias", seed, (4 * width,), 0.02, fp16=True)
    sd[prefix + "mlp.c_proj.weight"] = _normal(prefix + "mlp.c_proj.weight", seed, (width, 4 * width), proj_std, fp16=True)
    sd[prefix + "mlp.c_proj.bias"] = _normal(prefix + "mlp.c_proj.bias", seed, (width,), 0.02, fp16=True)
    sd[prefix + "ln_2.weight"] = _normal(prefix + "ln_2.weight", seed, (width,), 0.05, mean=1.0)
    sd[prefix + "ln_2.bias"] = _normal(prefix + "ln_2.bias", seed, (width,), 0.02)


def synth_clip_state_dict(arch: str | dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """fp32 state dict with the reference's key names; load with `CLIP(**a).load_state_dict(sd)`."""
    a = ARCHS[arch] if isinstance(arch, str) else arch
    vw, vl, p = a["vision_width"], a["vision_layers"], a["vision_patch_size"]
    tw, tl = a["transformer_width"], a["transformer_layers"]
    e = a["embed_dim"]
    grid = a["image_resolution"] // p
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    vs = vw ** -0.5
    sd["visual.class_embedding"] = _normal("visual.class_embedding", seed, (vw,), vs)
    sd["visual.positional_embedding"] = _normal("visual.positional_embedding", seed, (grid * grid + 1, vw), vs)
    sd["visual.proj"] = _normal("visual.proj", seed, (vw, e), vs, fp16=True)
    sd["visual.conv1.weight"] = _normal("visual.conv1.weight", seed, (vw, 3, p, p), (3 * p * p) ** -0.5, fp16=True)
    sd["visual.ln_pre.weight"] = _normal("visual.ln_pre.weight", seed, (vw,), 0.05, mean=1.0)
    sd["visual.ln_pre.bias"] = _normal("visual.ln_pre.bias", seed, (vw,), 0.02)
    v_proj = (vw ** -0.5) * ((2 * vl) ** -0.5)
    for i in range(vl):
        _block(sd, f"visual.transformer.resblocks.{i}.", vw, seed, vw ** -0.5, v_proj, (2 * vw) ** -0.5)
    sd["visual.ln_post.weight"] = _normal("visual.ln_post.weight", seed, (vw,), 0.05, mean=1.0)
    sd["visual.ln_post.bias"] = _normal("visual.ln_post.bias", seed, (vw,), 0.02)
    t_proj = (tw ** -0.5) * ((2 * tl) ** -0.5)
    for i in range(tl):
        _block(sd, f"transformer.resblocks.{i}.", tw, seed, tw ** -0.5, t_proj, (2 * tw) ** -0.5)
    sd["token_embedding.weight"] = _normal("token_embedding.weight", seed, (a["vocab_size"], tw), 0.02)
    sd["positional_embedding"] = _normal("positional_embedding", seed, (a["context_length"], tw), 0.01)
    sd["ln_final.weight"] = _normal("ln_final.weight", seed, (tw,), 0.05, mean=1.0)
    sd["ln_final.bias"] = _normal("ln_final.bias", seed, (tw,), 0.02)
    sd["text_projection"] = _normal("text_projection", seed, (tw, e), tw ** -0.5, fp16=True)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
    return sd


def synth_prompt_params(arch: str | dict, coop_n_ctx: int, vpt_n_ctx: int, vpt_deep: bool, csc_classes: int = 0,
                        project_dim: int = 0, seed: int = 0, cocoop_n_ctx: int = 0, vpt_project: int = -1) -> dict:
    """Trainable prompt tensors under the reference's prompt_learner key names (trainers/mvlpt.py:169-257).

    Values are fp16-representable (the reference creates them in CLIP's dtype, trainers/mvlpt.py:151,188-197,222-225).
    `project_dim` > 0 adds the UPT projection modules (PROJECT_METHOD='transformer'); `vpt_project` > -1 stores the
    visual prompts at that width and adds the vpt_proj Linear (:170-175).
    """
    a = ARCHS[arch] if isinstance(arch, str) else arch
    vw, vl, p, tw = a["vision_width"], a["vision_layers"], a["vision_patch_size"], a["transformer_width"]
    out = {}
    if vpt_n_ctx:
        pw = vpt_project if vpt_project > -1 else vw
        val = math.sqrt(6.0 / float(3 * p * p + pw))
        u = torch.rand(1, vpt_n_ctx, pw, generator=_gen("vpt_embeddings", seed)) * 2 - 1
        out["vpt_embeddings"] = (u * val).half().float()
        if vpt_deep:
            u = torch.rand(vl - 1, vpt_n_ctx, pw, generator=_gen("vpt_embeddings_deep", seed)) * 2 - 1
            out["vpt_embeddings_deep"] = (u * val).half().float()
        if vpt_project > -1:  # kaiming_normal_(fan_out): std = sqrt(2 / d)
            out["vpt_proj.weight"] = _normal("vpt_proj.weight", seed, (vw, pw), (2.0 / vw) ** 0.5, fp16=True)
            out["vpt_proj.bias"] = _normal("vpt_proj.bias", seed, (vw,), 0.02, fp16=True)
    if coop_n_ctx:
        shape = (csc_classes, coop_n_ctx, tw) if csc_classes else (coop_n_ctx, tw)
        out["ctx"] = _normal("ctx", seed, shape, 0.02, fp16=True)
    if cocoop_n_ctx:  # trainers/mvlpt.py:260-290: instance-conditioned context + meta network (e -> e/16 -> d_t)
        e = a["embed_dim"]
        out["cocoop_ctx"] = _normal("cocoop_ctx", seed, (cocoop_n_ctx, tw), 0.02, fp16=True)
        out["meta_net.linear1.weight"] = _normal("meta_net.linear1.weight", seed, (e // 16, e), e ** -0.5, fp16=True)
        out["meta_net.linear1.bias"] = _normal("meta_net.linear1.bias", seed, (e // 16,), 0.02, fp16=True)
        out["meta_net.linear2.weight"] = _normal("meta_net.linear2.weight", seed, (tw, e // 16), (e // 16) ** -0.5, fp16=True)
        out["meta_net.linear2.bias"] = _normal("meta_net.linear2.bias", seed, (tw,), 0.02, fp16=True)
    if project_dim and coop_n_ctx and vpt_n_ctx:
        pd = project_dim
        for nm, (i, o) in {"mvlpt_proj_ctx_coop_pre": (tw, pd), "mvlpt_proj_ctx_coop_post": (pd, tw),
                           "mvlpt_proj_ctx_vpt_pre": (vw, pd), "mvlpt_proj_ctx_vpt_post": (pd, vw)}.items():
            out[nm + ".weight"] = _normal(nm + ".weight", seed, (o, i), i ** -0.5, fp16=True)
            out[nm + ".bias"] = _normal(nm + ".bias", seed, (o,), 0.02, fp16=True)
        blk: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        _block(blk, "mvlpt_proj.resblocks.0.", pd, seed, pd ** -0.5, (pd ** -0.5) * (2 ** -0.5), (2 * pd) ** -0.5)
        out.update(blk)  # the projection block itself stays fp32 in the reference (trainers/mvlpt.py:256-259)
    return out


def synth_images(batch: int, resolution: int, seed: int = 1) -> torch.Tensor:
    """[B,3,H,W] fp16-representable fp32 images ~N(0,1) (normalised-pixel statistics)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    return torch.randn(batch, 3, resolution, resolution, generator=g).half().float()


def synth_token_ids(n_cls: int, n_ctx: int, context_length: int = 77, seed: int = 3, max_name_len: int = 6,
                    placeholder_id: int = 343, vocab: int = 49408):
    """Random class-name token sequences shaped like the reference tokenizer's output for "X X .. X name."

    Layout per row: [SOT=49406, n_ctx x placeholder('X'=343), name tokens, '.'=269, EOT=49407, 0 padding]
    (clip/clip.py:187-223, trainers/mvlpt.py:292-305).  Returns (tokens [C,77] int64, name_lens list).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 5)
    toks = torch.zeros(n_cls, context_length, dtype=torch.long)
    name_lens = []
    for c in range(n_cls):
        nl = int(torch.randint(1, max_name_len + 1, (1,), generator=g))
        name = torch.randint(1000, vocab - 1000, (nl,), generator=g)
        row = [49406] + [placeholder_id] * n_ctx + name.tolist() + [269, 49407]
        toks[c, :len(row)] = torch.tensor(row)
        name_lens.append(nl)
    return toks, name_lens


class SyntheticDataManager:
    """The attribute surface the MVLPT trainer reads from a Dassl DataManager (`dataset.classnames`,
    `lab2cname`, `num_classes`, `num_source_domains`, `train_loader_x`, `val_loader`, `test_loader`) over synthetic
    batches in the reference's CoOp-data format {'img', 'label', 'domain'} (trainers/mvlpt.py:953-968).  Loaders are
    lists of pinned host batches; the last test batch is ragged, as Dassl's `drop_last=False` test loaders are.
    Offline stand-in for the reference's dataset plumbing (out of scope): used by
    `python -m mvlpt_b200.train --synthetic`, tests and benchmarks."""

    def __init__(self, arch: str, n_classes: int, train_batches: int = 4, batch_size: int = 8, test_images: int = 20,
                 test_batch_size: int = 8, seed: int = 0, pin: bool = True, half: bool = True):
        import types
        res = ARCHS[arch]["image_resolution"]
        names = [f"class{c}" for c in range(n_classes)]
        self.dataset = types.SimpleNamespace(classnames=names)
        self.lab2cname = {i: n for i, n in enumerate(names)}
        self.num_classes = n_classes
        self.num_source_domains = 1

        def batch(n, s):
            img = synth_images(n, res, seed=s)
            img = img.half() if half else img
            g = torch.Generator().manual_seed(31 * s + 7)
            lab = torch.randint(0, n_classes, (n,), generator=g)
            if pin and torch.cuda.is_available():
                img, lab = img.pin_memory(), lab.pin_memory()
            return {"img": img, "label": lab, "domain": torch.zeros(n, dtype=torch.long)}

        def split(total, bs, s0):
            out, done = [], 0
            while done < total:
                n = min(bs, total - done)
                out.append(batch(n, s0 + len(out)))
                done += n
            return out

        self.train_loader_x = [batch(batch_size, 1000 * seed + i) for i in range(train_batches)]
        self.train_loader_u = None
        self.val_loader = split(test_images, test_batch_size, 1000 * seed + 500)
        self.test_loader = split(test_images, test_batch_size, 1000 * seed + 700)
