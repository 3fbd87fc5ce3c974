"""B200-native mirror of the reference's trainers/mvlpt.py hot path.

Same class names, constructor signatures, attribute names and state-dict keys as the reference
(MultitaskVLPromptLearner :138-515, ImageEncoder :45-93, TextEncoder :95-130, CustomCLIP :517-583), so a Dassl
config / checkpoint drops in unchanged — but no arithmetic happens in torch: forward and backward are sequenced
by mvlpt_b200.engine over the sm_100a kernels of libmvlpt_sm100.so.  There is no CPU or eager fallback: calling a
module with CPU tensors, or without the built library, raises.

What is deliberately different (result-identical):
  * prompts are never concatenated into activations — kernels read them from the prompt tables;
  * 'middle'/'front' class-token positions use a precomputed index map instead of a Python loop over classes;
  * text features are cached when no text-side parameter trains (VPT only);
  * the EOT index is precomputed on the device (no per-call CPU argmax);
  * activations keep an fp32 residual stream (the reference's is fp16), which only moves results closer to exact.
"""
from __future__ import annotations

import os

import math
from collections import OrderedDict
from functools import reduce
from operator import mul
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import engine as E
from .. import ops
from ..clip_model import FrozenCLIP, as_state_dict, build_model
from ..tokenizer import EOT, SOT, get_tokenizer, tokenize
from ..upt import UptProjection, VptProjection

__all__ = ["load_clip_to_cpu", "ImageEncoder", "TextEncoder", "MultitaskVLPromptLearner", "CustomCLIP", "MVLPT"]

GRAD_SCALE = 4096.0  # power of two: exact; keeps fp16 dgrad operands in range (see engine.py)


def load_clip_to_cpu(cfg, state_dict: Optional[Dict[str, torch.Tensor]] = None) -> FrozenCLIP:
    """trainers/mvlpt.py:28-43.  The reference downloads a checkpoint by backbone name; offline the caller supplies
    the state dict (or MODEL.BACKBONE.PATH points at a torch-saved one)."""
    if state_dict is None:
        path = getattr(cfg.MODEL.BACKBONE, "PATH", None)
        if not path:
            raise RuntimeError("no network: pass state_dict= or set cfg.MODEL.BACKBONE.PATH to a CLIP state dict")
        obj = torch.load(path, map_location="cpu")
        state_dict = obj.state_dict() if hasattr(obj, "state_dict") else obj
    return build_model(state_dict)


# ---------------------------------------------------------------------------------------------------------------
# parameter containers that reproduce the reference's state-dict key names for the UPT projection
# ---------------------------------------------------------------------------------------------------------------
class _Linear(nn.Module):
    def __init__(self, i: int, o: int, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i, dtype=dtype))
        self.bias = nn.Parameter(torch.empty(o, dtype=dtype))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(i)
        nn.init.uniform_(self.bias, -bound, bound)


class _LN(nn.Module):
    def __init__(self, w: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(w))
        self.bias = nn.Parameter(torch.zeros(w))


class _Attn(nn.Module):
    def __init__(self, w: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * w, w))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * w))
        self.out_proj = _Linear(w, w, torch.float32)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _ProjBlock(nn.Module):
    """Parameters of the width-`w` ResidualAttentionBlock the reference uses as mvlpt_proj (fp32, :256-259)."""

    def __init__(self, w: int):
        super().__init__()
        self.attn = _Attn(w)
        self.ln_1 = _LN(w)
        self.mlp = nn.ModuleDict(OrderedDict(c_fc=_Linear(w, 4 * w, torch.float32),
                                             c_proj=_Linear(4 * w, w, torch.float32)))
        self.ln_2 = _LN(w)


class _ProjTransformer(nn.Module):
    def __init__(self, w: int):
        super().__init__()
        self.width = w
        self.resblocks = nn.ModuleList([_ProjBlock(w)])


# ---------------------------------------------------------------------------------------------------------------
class MultitaskVLPromptLearner(nn.Module):
    """Owns every trainable tensor (trainers/mvlpt.py:139-325).  `tokenized_prompts` / `name_lens` may be supplied
    when no BPE tokenizer is available (fixtures, synthetic benchmarks)."""

    def __init__(self, cfg, classnames, clip_model, tokenized_prompts: Optional[torch.Tensor] = None,
                 name_lens: Optional[Sequence[int]] = None):
        super().__init__()
        T = cfg.TRAINER.MVLPT
        n_cls = len(classnames)
        coop_n_ctx, cocoop_n_ctx, vpt_n_ctx = T.COOP.N_CTX, T.COCOOP.N_CTX, T.VPT.N_CTX
        if cocoop_n_ctx != 0 and coop_n_ctx != 0:
            raise NotImplementedError("COOP.N_CTX and COCOOP.N_CTX together: the reference's own token_suffix then no "
                                      "longer matches forward_coop (trainers/mvlpt.py:312-315); not supported")
        dtype = clip_model.dtype
        coop_ctx_dim = clip_model.ln_final.weight.shape[0]
        vpt_ctx_dim = clip_model.visual.conv1.weight.shape[0]
        patch = clip_model.visual.conv1.weight.shape[-1]
        clip_imsize = clip_model.visual.input_resolution
        cfg_imsize = cfg.INPUT.SIZE[0]
        assert cfg_imsize == clip_imsize, f"cfg_imsize ({cfg_imsize}) must equal to clip_imsize ({clip_imsize})"

        self.vpt_dropout = nn.Dropout(T.VPT.DROPOUT)  # applied inside the tower kernels (drop_state below)
        self.drop_seed_override: Optional[int] = None
        self.vpt_deep = T.VPT.DEEP
        self.vpt_embeddings = None
        self.vpt_embeddings_deep = None
        prompt_prefix = None
        if vpt_n_ctx != 0:
            if T.VPT.PROJECT > -1:  # trainers/mvlpt.py:170-175: prompts stored at width PROJECT, Linear up to the tower's
                vpt_dim = T.VPT.PROJECT
                self.vpt_proj = nn.Linear(vpt_dim, vpt_ctx_dim).type(dtype)
                nn.init.kaiming_normal_(self.vpt_proj.weight, a=0, mode="fan_out")
            else:
                vpt_dim = vpt_ctx_dim
                self.vpt_proj = nn.Identity()
            if T.VPT.CTX_INIT:
                raise ValueError("CTX initiation scheme is not supported")
            val = math.sqrt(6.0 / float(3 * reduce(mul, (patch, patch), 1) + vpt_dim))
            self.vpt_embeddings = nn.Parameter(torch.zeros(1, vpt_n_ctx, vpt_dim, dtype=dtype))
            nn.init.uniform_(self.vpt_embeddings.data, -val, val)
            if self.vpt_deep:
                self.vision_layers = len([k for k in clip_model.state_dict().keys()
                                          if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
                self.vpt_embeddings_deep = nn.Parameter(
                    torch.zeros(self.vision_layers - 1, vpt_n_ctx, vpt_dim, dtype=dtype))
                nn.init.uniform_(self.vpt_embeddings_deep.data, -val, val)
            prompt_prefix = "a photo of a "

        self.ctx = None
        if coop_n_ctx != 0:
            if T.COOP.CTX_INIT:
                init = T.COOP.CTX_INIT.replace("_", " ")
                coop_n_ctx = len(init.split(" "))
                with torch.no_grad():
                    emb = clip_model.token_embedding(tokenize(init)).type(dtype)
                ctx_vectors = emb[0, 1:1 + coop_n_ctx, :].clone()
                prompt_prefix = init
            else:
                shape = (n_cls, coop_n_ctx, coop_ctx_dim) if T.COOP.CSC else (coop_n_ctx, coop_ctx_dim)
                ctx_vectors = torch.empty(*shape, dtype=dtype)
                nn.init.normal_(ctx_vectors, std=0.02)
                prompt_prefix = " ".join(["X"] * coop_n_ctx)
            self.ctx = nn.Parameter(ctx_vectors)

        self.mvlpt_proj = nn.Identity()
        if vpt_n_ctx != 0 and coop_n_ctx != 0:
            self.mvlpt_proj_ctx_dim = T.PROJECT_DIM
            method = T.PROJECT_METHOD
            if method == "identity":
                pass
            elif method == "transformer":
                if T.VPT.PROJECT > -1:
                    # the reference sizes mvlpt_proj_ctx_vpt_pre for the tower width (trainers/mvlpt.py:243-246) and then
                    # feeds it PROJECT-wide prompts: a shape error there, an explicit one here.  PROJECT == vision width
                    # would run there (UPT projection, then vpt_proj on its output); that chain has no backward here and
                    # no shipped config uses it, so it is refused as well instead of returning wrong gradients
                    raise NotImplementedError("VPT.PROJECT together with PROJECT_METHOD='transformer' is not supported "
                                              "(shape-inconsistent in the reference, trainers/mvlpt.py:243-246,389, "
                                              "unless PROJECT equals the vision width)")
                pd = T.PROJECT_DIM
                ident = lambda: nn.Identity()
                self.mvlpt_proj_ctx_vpt_pre, self.mvlpt_proj_ctx_vpt_post = ident(), ident()
                self.mvlpt_proj_ctx_coop_pre, self.mvlpt_proj_ctx_coop_post = ident(), ident()
                if coop_ctx_dim != pd:
                    self.mvlpt_proj_ctx_coop_pre = _Linear(coop_ctx_dim, pd, dtype)
                    self.mvlpt_proj_ctx_coop_post = _Linear(pd, coop_ctx_dim, dtype)
                if vpt_ctx_dim != pd:
                    self.mvlpt_proj_ctx_vpt_pre = _Linear(vpt_ctx_dim, pd, dtype)
                    self.mvlpt_proj_ctx_vpt_post = _Linear(pd, vpt_ctx_dim, dtype)
                self.mvlpt_proj = _ProjTransformer(pd)
            elif method == "mlp":
                # the reference crashes here too (nn.GeLU does not exist, trainers/mvlpt.py:252-253)
                raise AttributeError("module 'torch.nn' has no attribute 'GeLU'")
            else:
                raise ValueError(f"unknown PROJECT_METHOD {method!r}")
        # CoCoOp (trainers/mvlpt.py:260-290): instance-conditioned context = cocoop_ctx + meta_net(image feature)
        self.cocoop_ctx = None
        self.meta_net = None
        if cocoop_n_ctx != 0:
            if T.COCOOP.CTX_INIT:
                init = T.COCOOP.CTX_INIT.replace("_", " ")
                cocoop_n_ctx = len(init.split(" "))
                with torch.no_grad():
                    emb = clip_model.token_embedding(tokenize(init)).type(dtype)
                ctx_vectors = emb[0, 1:1 + cocoop_n_ctx, :].clone()
                prompt_prefix = init
            else:
                ctx_vectors = torch.empty(cocoop_n_ctx, coop_ctx_dim, dtype=dtype)
                nn.init.normal_(ctx_vectors, std=0.02)
                prompt_prefix = " ".join(["X"] * cocoop_n_ctx)
            self.cocoop_ctx = nn.Parameter(ctx_vectors)
            vis_dim = clip_model.visual.output_dim
            self.meta_net = nn.Sequential(OrderedDict([
                ("linear1", nn.Linear(vis_dim, vis_dim // 16)),
                ("relu", nn.ReLU(inplace=True)),
                ("linear2", nn.Linear(vis_dim // 16, coop_ctx_dim)),
            ]))
            if T.COCOOP.PREC == "fp16":
                self.meta_net.half()

        if prompt_prefix is None:
            raise ValueError("at least one of VPT.N_CTX / COOP.N_CTX must be non-zero")
        classnames = [name.replace("_", " ") for name in classnames]
        if tokenized_prompts is None:
            tok = get_tokenizer()
            name_lens = [len(tok.encode(name)) for name in classnames]
            prompts = [prompt_prefix + " " + name + "." for name in classnames]
            if cfg.TRAINER.CUT_CONTEXTLEN:
                max_length = min(clip_model.context_length, max(len(tok.encode(p)) + 2 for p in prompts))
            else:
                max_length = clip_model.context_length
            tokenized_prompts = torch.cat([tokenize(p, context_length=max_length, tokenizer=tok) for p in prompts])
        else:
            if name_lens is None:
                raise ValueError("name_lens must accompany tokenized_prompts")
            tokenized_prompts = tokenized_prompts.clone().long()
            if cfg.TRAINER.CUT_CONTEXTLEN:
                used = int((tokenized_prompts != 0).sum(dim=1).max())
                tokenized_prompts = tokenized_prompts[:, :min(clip_model.context_length, used)]
        with torch.no_grad():
            embedding = clip_model.token_embedding(tokenized_prompts).type(dtype)
        # rows 1..n_text of every prompt are the learned context (CoOp's, or CoCoOp's instance-shifted one, :312-315)
        n_text = cocoop_n_ctx if cocoop_n_ctx != 0 else coop_n_ctx
        self.text_n_ctx = n_text
        # saved with the checkpoint, ignored on load (trainers/mvlpt.py:309-316, 1117-1121)
        self.register_buffer("token_prefix", embedding[:, :1, :].clone())
        self.register_buffer("token_suffix", embedding[:, 1 + n_text:, :].clone())

        self.n_cls = n_cls
        self.vpt_n_ctx = vpt_n_ctx
        self.coop_n_ctx = coop_n_ctx
        self.cocoop_n_ctx = cocoop_n_ctx
        self.tokenized_prompts = tokenized_prompts
        self.name_lens = list(name_lens)
        self.class_token_position = T.COOP.CLASS_TOKEN_POSITION
        self.csc = bool(self.ctx is not None and self.ctx.dim() == 3)
        # kernel-side view of the prompt layout (built once; replaces the per-class cat loop of :472-510)
        Lt = tokenized_prompts.shape[1]
        full = torch.cat([embedding[:, :1], torch.zeros(n_cls, n_text, coop_ctx_dim, dtype=dtype),
                          embedding[:, 1 + n_text:]], dim=1).float()
        pos = self.class_token_position if coop_n_ctx else "end"  # forward_cocoop always appends the class name (:369)
        emb_full = E.rearrange_embedding(full, self.name_lens, n_text, pos)
        slot, ctx_pos = E.build_ctx_maps(self.name_lens, n_text, Lt, pos)
        eot = tokenized_prompts.argmax(dim=-1)
        # Causal cut.  The text tower is causal (clip/model.py:324-330) and only the EOT row of each class is read
        # (trainers/mvlpt.py:124-128): rows after the last EOT of the class set can influence neither the features nor
        # any gradient, so the kernels run on the first `kernel_len` rows only.  Bit-for-bit the same features as the
        # full 77 rows; it is what TRAINER.CUT_CONTEXTLEN (trainers/mvlpt.py:297-300) does to the tokens, applied to
        # the computation.  MVLPT_TEXT_CAUSAL_CUT=0 keeps every row.
        self.context_len = Lt
        self.kernel_len = int(eot.max()) + 1 if os.environ.get("MVLPT_TEXT_CAUSAL_CUT", "1") != "0" else Lt
        Lk = self.kernel_len
        self.register_buffer("_emb", emb_full[:, :Lk].contiguous(), persistent=False)
        self.register_buffer("_slot", slot[:, :Lk].contiguous(), persistent=False)
        self.register_buffer("_ctx_pos", ctx_pos, persistent=False)
        self.register_buffer("_eot_rows", (torch.arange(n_cls) * Lk + eot).to(torch.int32), persistent=False)
        self._upt: Optional[UptProjection] = None
        self._vproj: Optional[VptProjection] = None

    # ---- vpt_dropout -----------------------------------------------------------------------------------------
    def drop_state(self):
        """(p, seed) of this forward pass: nn.Dropout semantics — active iff the module is in training mode and p > 0.
        The seed of the counter-based mask comes from torch's CPU generator (so torch.manual_seed reproduces a run)."""
        p = float(self.vpt_dropout.p)
        if p <= 0.0 or not self.vpt_dropout.training or self.vpt_embeddings is None:
            return 0.0, 0
        if self.drop_seed_override is not None:
            return p, int(self.drop_seed_override)
        return p, int(torch.randint(0, 2 ** 62, (1,)).item())

    # ---- vpt_proj --------------------------------------------------------------------------------------------
    @property
    def uses_vpt_proj(self) -> bool:
        return self.vpt_embeddings is not None and isinstance(self.vpt_proj, nn.Linear)

    def vproj(self) -> "VptProjection":
        if self._vproj is None:
            self._vproj = VptProjection(self)
        return self._vproj

    # ---- UPT -------------------------------------------------------------------------------------------------
    @property
    def uses_projection(self) -> bool:
        return not (self.coop_n_ctx == 0 or isinstance(self.mvlpt_proj, nn.Identity) or self.vpt_n_ctx == 0)

    def upt(self) -> UptProjection:
        if self._upt is None:
            self._upt = UptProjection(self)
        return self._upt

    def forward_mvlpt_proj(self, dtype=torch.float):
        """trainers/mvlpt.py:376-414 -> (ctx', vpt', vpt_deep').  Identity pass-through unless both prompt kinds are on
        and PROJECT_METHOD='transformer'."""
        if not self.uses_projection:
            return self.ctx, self.vpt_embeddings, self.vpt_embeddings_deep
        return self.upt().forward()

    # ---- assembly helpers kept for API parity (the fused kernels do this inside the towers) --------------------
    def forward_vpt(self, x, vpt_embeddings=None):
        """trainers/mvlpt.py:416-437.  The hot path never calls this (embed_assemble writes the rows directly)."""
        if vpt_embeddings is None:
            if self.vpt_embeddings is None:
                return x
            vpt_embeddings = self.vpt_embeddings
        B = x.shape[0]
        if self.uses_vpt_proj:
            vpt_embeddings = self.vproj().forward(vpt_embeddings, None)[0]
        vpt_embeddings = self.vpt_dropout(vpt_embeddings.expand(B, -1, -1))
        return torch.cat([x[:, :1, :], vpt_embeddings.expand(B, -1, -1).to(x.dtype), x[:, 1:, :]], dim=1)

    def forward_coop(self, ctx=None):
        """trainers/mvlpt.py:439-515: assembled prompt embeddings [n_cls, L_t, d_t] (without positional embedding)."""
        if ctx is None:
            ctx = self.ctx
        if self.class_token_position not in ("end", "middle", "front"):
            raise ValueError
        C, Lk, d = self._emb.shape
        _require_cuda(self._emb, "forward_coop")
        out = torch.empty(C, Lk, d, device=self._emb.device, dtype=torch.float32)
        zero_pos = torch.zeros(Lk, d, device=self._emb.device, dtype=torch.float32)
        c = None if ctx is None else ctx.detach().contiguous()
        ops.text_assemble(self._emb, c, self._slot, zero_pos, out, C, Lk, self.coop_n_ctx, d, self.csc)
        out = out.to(self._emb.dtype if ctx is None else ctx.dtype)
        if Lk < self.context_len:  # the rows behind the causal cut are the untouched token embeddings (API parity)
            tail = self.token_suffix[:, Lk - 1 - self.text_n_ctx:, :].to(out.dtype)
            out = torch.cat([out, tail], dim=1)
        return out

    def meta_params(self):
        m = self.meta_net
        return m.linear1.weight, m.linear1.bias, m.linear2.weight, m.linear2.bias

    def forward_cocoop(self, im_features):
        """trainers/mvlpt.py:348-374: prompts [B, n_cls, L_t, d_t] with the context shifted by meta_net(im_features)
        (without positional embedding).  API parity; the fused path assembles straight into the text tower's input."""
        if self.cocoop_ctx is None:
            return torch.cat([self.token_prefix, self.token_suffix], dim=1)
        _require_cuda(self._emb, "forward_cocoop")
        dev = self._emb.device
        B, e = im_features.shape
        C, Lk, d = self._emb.shape
        W1, b1, W2, b2 = [t.detach().contiguous() for t in self.meta_params()]
        h1 = torch.empty(B, W1.shape[0], device=dev, dtype=torch.float32)
        bias = torch.empty(B, d, device=dev, dtype=torch.float32)
        ops.metanet_fwd(im_features.detach().float().contiguous(), W1, b1, W2, b2, h1, bias)
        out = torch.empty(B, C, Lk, d, device=dev, dtype=torch.float32)
        zero_pos = torch.zeros(Lk, d, device=dev, dtype=torch.float32)
        ops.cocoop_assemble(self._emb, self.cocoop_ctx.detach().contiguous(), bias, self._slot, zero_pos, out, B, C, Lk, d)
        out = out.to(self.cocoop_ctx.dtype)
        if Lk < self.context_len:
            tail = self.token_suffix[:, Lk - 1 - self.text_n_ctx:, :].to(out.dtype)
            out = torch.cat([out, tail.unsqueeze(0).expand(B, -1, -1, -1)], dim=2)
        return out

    def construct_prompts(self, ctx, prefix, suffix, label=None):
        """trainers/mvlpt.py:327-346 (API parity)."""
        if label is not None:
            prefix, suffix = prefix[label], suffix[label]
        return torch.cat([prefix, ctx, suffix], dim=1)


def _require_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise ops._lib.MvlptError(f"{who}: tensors must live on a CUDA device (sm_100a); there is no CPU path")


class ImageEncoder(nn.Module):
    """trainers/mvlpt.py:45-93: forward(x, vpt_embeddings, vpt_embeddings_deep) -> [B, embed_dim] (forward only when
    called directly; gradients flow through CustomCLIP)."""

    def __init__(self, clip_model, mvlpt_model):
        super().__init__()
        self._sd = as_state_dict(clip_model)
        self.__dict__["mvlpt_model"] = mvlpt_model  # not a submodule (avoid a registration cycle)
        self._tower: Optional[E.ImageTower] = None
        self._out_dtype = clip_model.dtype

    def tower(self, device) -> E.ImageTower:
        if self._tower is None:
            self._tower = E.ImageTower(self._sd, device)
        return self._tower

    @torch.no_grad()
    def forward(self, x: torch.Tensor, vpt_embeddings=None, vpt_embeddings_deep=None):
        _require_cuda(x, "ImageEncoder")
        m = self.mvlpt_model
        if vpt_embeddings is None:
            vpt_embeddings = m.vpt_embeddings
        deep = None
        if m.vpt_deep and (vpt_embeddings_deep is not None or m.vpt_embeddings_deep is not None):
            deep = vpt_embeddings_deep if vpt_embeddings_deep is not None else m.vpt_embeddings_deep
        x = x.contiguous()
        if x.dtype not in (torch.float16, torch.float32):
            x = x.float()
        cont = lambda t: None if t is None else t.detach().contiguous()
        vpt_embeddings, deep = cont(vpt_embeddings), cont(deep)
        if vpt_embeddings is not None and m.uses_vpt_proj:
            vpt_embeddings, deep = m.vproj().forward(vpt_embeddings, deep)
        drop_p, drop_seed = m.drop_state()
        feat = self.tower(x.device).forward(x, vpt_embeddings, deep, False, drop_p, drop_seed)
        return feat.to(self._out_dtype)


class TextEncoder(nn.Module):
    """trainers/mvlpt.py:95-130: forward(prompts, tokenized_prompts) -> [n_cls, embed_dim] (forward only when called
    directly).  CUT_CONTEXTLEN needs no mask surgery here (the causal mask is implicit) and ACT_CKPT has no B200
    counterpart (everything fits in HBM)."""

    def __init__(self, clip_model, cfg=None):
        super().__init__()
        self._sd = as_state_dict(clip_model)
        self.dtype = clip_model.dtype
        self.cfg = cfg
        self._tower: Optional[E.TextTower] = None

    def tower(self, device) -> E.TextTower:
        if self._tower is None:
            self._tower = E.TextTower(self._sd, device)
        return self._tower

    @torch.no_grad()
    def forward(self, prompts: torch.Tensor, tokenized_prompts: torch.Tensor):
        _require_cuda(prompts, "TextEncoder")
        C, Lt, _ = prompts.shape
        eot = tokenized_prompts.argmax(dim=-1).to(prompts.device)
        if os.environ.get("MVLPT_TEXT_CAUSAL_CUT", "1") != "0":
            Lt = int(eot.max()) + 1  # rows after the last EOT cannot reach any EOT row through the causal mask
            prompts = prompts[:, :Lt]
        rows = (torch.arange(C, device=prompts.device) * Lt + eot).to(torch.int32)
        feat = self.tower(prompts.device).forward(prompts.float().contiguous(), None, None, rows, 0, False, train=False)
        return feat.to(self.dtype)


class _CustomCLIPFn(torch.autograd.Function):
    """logits = f(image; prompt tensors): forward/backward are the hand-written engine passes."""

    @staticmethod
    def forward(ctx, model, image, task, train, *params):
        logits = model._forward_core(image, task, train)
        ctx.model, ctx.B, ctx.task, ctx.train = model, image.shape[0], task, train
        ctx.n = len(params)
        # a COPY: the engine's logits buffer is reused by the next call with the same (B, C)
        return logits[:, :model.prompt_learner.n_cls].to(model.dtype, copy=True)

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        if not ctx.train:
            raise RuntimeError("backward through a CustomCLIP forward that ran without gradient tracking")
        grads = model._backward_core(ctx.B, dlogits=dlogits, task=ctx.task)
        out = []
        for name, p in model._trainables():
            g = grads.get(name)
            # the engine's gradients live in a reused flat buffer: hand autograd its own copy
            out.append(None if g is None else g.reshape(p.shape).to(p.dtype, copy=True))
        return (None, None, None, None, *out)


class CustomCLIP(nn.Module):
    """trainers/mvlpt.py:517-583."""

    def __init__(self, cfg, classnames, clip_model, dm=None, tokenized_prompts=None, name_lens=None):
        super().__init__()
        self.prompt_learner = MultitaskVLPromptLearner(cfg, classnames, clip_model, tokenized_prompts, name_lens)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.image_encoder = ImageEncoder(clip_model, self.prompt_learner)
        self.text_encoder = TextEncoder(clip_model, cfg)
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.grad_scale = GRAD_SCALE
        self.cache_text_features = True
        self._txt_cache_valid = False
        # data-parallel group (set by the trainer).  With more than one rank the text tower is CLASS-SHARDED: rank r
        # encodes classes dp.shard(n_cls), features are all-gathered, their gradients reduce-scattered
        self.dp = None
        self.shard_text = os.environ.get("MVLPT_SHARD_TEXT", "1") != "0"
        self._shard_cache = {}
        self._head: Optional[E.LogitHead] = None
        self._embed_dim = clip_model.visual.output_dim

        self.multi_task_label_pertask = cfg.DATASET.MULTITASK_LABEL_PERTASK
        self._task_ranges = None
        if self.multi_task_label_pertask:
            # per-task [start, end) class ranges, in task order (trainers/mvlpt.py:527-538)
            rng, start = [], 0
            for task in dm._task_names:
                n = len(dm._labelmap[task])
                rng.append((start, start + n))
                start += n
            self.class_index_pertask_start = torch.tensor([r[0] for r in rng])
            self.class_index_pertask_end = torch.tensor([r[1] for r in rng])
            self.register_buffer("_ranges", torch.tensor(rng, dtype=torch.int32), persistent=False)

    # ---- plumbing ----------------------------------------------------------------------------------------------
    def _trainables(self):
        return [(n, p) for n, p in self.prompt_learner.named_parameters()]

    def head(self, device) -> E.LogitHead:
        if self._head is None:
            self._head = E.LogitHead(float(self.logit_scale), self._embed_dim, device)
        return self._head

    def _text_shard(self, device):
        """Class range of this rank and the exchange buffers of the class-sharded text tower (None when not sharded)."""
        dp = self.dp
        if dp is None or not self.shard_text or dp.world <= 1:
            return None
        pl = self.prompt_learner
        C, Lt, e = pl.n_cls, pl._emb.shape[1], self._embed_dim
        key = (C, Lt, dp.world, dp.rank, str(device))
        if key not in self._shard_cache:
            ranges = [dp.shard(C, r) for r in range(dp.world)]
            c0, c1 = ranges[dp.rank]
            cmax = max(b - a for a, b in ranges)
            z = lambda *sz, dt=torch.float32: torch.zeros(*sz, device=device, dtype=dt)
            self._shard_cache[key] = dict(
                range=(c0, c1), ranges=ranges, cmax=cmax,
                eot_rows=(pl._eot_rows[c0:c1] - c0 * Lt).to(torch.int32).contiguous(),
                send=z(cmax, e), gathered=z(dp.world, cmax, e), full=z(C, e),
                g_send=z(dp.world, cmax, e, dt=torch.float16), g_recv=z(cmax, e, dt=torch.float16))
        return self._shard_cache[key]

    def _task_dev(self, task, device):
        if task is None or not self.multi_task_label_pertask:
            return None, None
        return task.to(device=device, dtype=torch.int32).contiguous(), self._ranges

    # ---- engine passes -------------------------------------------------------------------------------------------
    def _forward_core(self, image: torch.Tensor, task, train: bool) -> torch.Tensor:
        """Runs projection -> image tower -> text tower -> logits (+ task mask); returns fp32 logits [B, ldc]."""
        _require_cuda(image, "CustomCLIP")
        pl = self.prompt_learner
        dev = image.device
        # a batch staged by the trainer on its copy stream carries the event that marks the end of its host->device copy
        ready = getattr(image, "_mvlpt_ready", None)
        B, C = image.shape[0], pl.n_cls
        ctx, vpt, deep = pl.forward_mvlpt_proj(self.dtype)
        if not pl.vpt_deep:
            deep = None
        cont = lambda t: None if t is None else t.detach().contiguous()
        ctx, vpt, deep = cont(ctx), cont(vpt), cont(deep)
        if vpt is not None and pl.uses_vpt_proj:
            vpt, deep = pl.vproj().forward(vpt, deep)
        self._img_train = bool(train and vpt is not None)
        self._txt_train = bool(train and (ctx is not None or pl.cocoop_ctx is not None))
        self._shapes = dict(B=B, C=C, v=0 if vpt is None else vpt.shape[1],
                            n_deep=None if deep is None else deep.shape[0], Lt=pl._emb.shape[1])
        head = self.head(dev)
        # the text tower does not depend on the images: it runs first, while the batch is still crossing PCIe.  (Issuing it
        # on a second stream, to fill the tails of the image tower's persistent kernels, was measured: no gain.)
        if pl.cocoop_ctx is not None:
            return self._forward_cocoop(image, ready, vpt, deep, task, train)
        held = getattr(self, "_txt_hold", False) and not train and getattr(self, "_txt_held_for", None) == C
        # the decision depends on the label space only (never on this rank's batch size): under the class-sharded text
        # tower every rank must enter the all-gather of _text_features together
        if not held and (ctx is not None or not (self.cache_text_features and self._txt_cache_valid == C)):
            self._text_features(dev, ctx, C)
            if getattr(self, "_txt_hold", False) and not train:
                self._txt_held_for = C
        if ready is not None:
            torch.cuda.current_stream(dev).wait_event(ready)
        image = image.contiguous()
        if image.dtype not in (torch.float16, torch.float32):
            image = image.float()
        drop_p, drop_seed = pl.drop_state()
        img_feat = self.image_encoder.tower(dev).forward(image, vpt, deep, self._img_train, drop_p, drop_seed)
        logits = head.logits(img_feat, C)
        t_dev, ranges = self._task_dev(task, dev)
        if t_dev is not None:
            ops.task_mask(logits, head.buffers(B, C)["ldc"], t_dev, ranges, B, C)
        return logits

    # ---- CoCoOp branch (trainers/mvlpt.py:556-571) -------------------------------------------------------------------
    def _cocoop_buffers(self, dev, B: int, C: int):
        pl = self.prompt_learner
        e, d, n = self._embed_dim, pl._emb.shape[2], pl.text_n_ctx
        H = pl.meta_net.linear1.weight.shape[0]
        key = (B, C, str(dev))
        if getattr(self, "_cc_key", None) != key:
            z = lambda *sz, dt=torch.float32: torch.zeros(*sz, device=dev, dtype=dt)
            Lk = pl._emb.shape[1]
            bc = torch.arange(B * C, device=dev, dtype=torch.int64)
            eot = (pl._eot_rows.to(torch.int64) - torch.arange(C, device=dev, dtype=torch.int64) * Lk)  # eot(c)
            self._cc = dict(h1=z(B, H), bias=z(B, d), t16=z(B * C, e, dt=torch.float16), t32=z(B * C, e), t_inv=z(B * C),
                            dt32=z(B * C, e), dtfeat16=z(B * C, e, dt=torch.float16), d_bias=z(B, d), d_h1=z(B, H),
                            part=z(B, n, d), eot_rows=(bc * Lk + eot.repeat(B)).to(torch.int32).contiguous())
            self._cc_key = key
        return self._cc

    def _forward_cocoop(self, image, ready, vpt, deep, task, train: bool) -> torch.Tensor:
        """Image tower -> normalised features -> meta_net -> B*C instance-conditioned prompts -> text tower on B*C
        sequences -> logits[b,c] = exp(logit_scale) <img_b, txt_{b,c}>."""
        pl = self.prompt_learner
        dev = image.device
        B, C = image.shape[0], pl.n_cls
        Lk, d = pl._emb.shape[1], pl._emb.shape[2]
        head = self.head(dev)
        hb = head.buffers(B, C)
        cc = self._cocoop_buffers(dev, B, C)
        if ready is not None:
            torch.cuda.current_stream(dev).wait_event(ready)
        image = image.contiguous()
        if image.dtype not in (torch.float16, torch.float32):
            image = image.float()
        drop_p, drop_seed = pl.drop_state()
        img_feat = self.image_encoder.tower(dev).forward(image, vpt, deep, self._img_train, drop_p, drop_seed)
        ops.l2norm_fwd(img_feat, hb["i16"], hb["i32"], hb["i_inv"], B, self._embed_dim)
        W1, b1, W2, b2 = [t.detach().contiguous() for t in pl.meta_params()]
        ops.metanet_fwd(hb["i32"], W1, b1, W2, b2, cc["h1"], cc["bias"])
        ctx = pl.cocoop_ctx.detach().contiguous()
        tt = self.text_encoder.tower(dev)
        txt = tt.forward_assembled(
            B * C, Lk, cc["eot_rows"], self._txt_train,
            lambda x0: ops.cocoop_assemble(pl._emb, ctx, cc["bias"], pl._slot, tt.pos, x0, B, C, Lk, d))
        ops.l2norm_fwd(txt, cc["t16"], cc["t32"], cc["t_inv"], B * C, self._embed_dim)
        ops.pair_logits_fwd(hb["i32"], cc["t32"], head.s, hb["logits"], hb["ldc"], B, C, self._embed_dim)
        logits = hb["logits"]
        t_dev, ranges = self._task_dev(task, dev)
        if t_dev is not None:
            ops.task_mask(logits, hb["ldc"], t_dev, ranges, B, C)
        return logits

    def _backward_cocoop(self, B: int) -> Dict[str, torch.Tensor]:
        """dz16 (scaled, from the cross-entropy kernel) -> gradients of cocoop_ctx, meta_net.* (and the visual prompts)."""
        pl = self.prompt_learner
        sh = self._shapes
        C, v, n_deep = sh["C"], sh["v"], sh["n_deep"]
        dev = pl._emb.device
        head = self.head(dev)
        hb = head.buffers(B, C)
        cc = self._cc
        e, Lk, d, n = self._embed_dim, pl._emb.shape[1], pl._emb.shape[2], pl.text_n_ctx
        inv = 1.0 / self.grad_scale
        views = self.grad_views()
        ops.pair_logits_bwd(hb["dz16"], hb["ldc"], hb["i32"], cc["t32"], head.s, cc["dt32"], hb["di32"], B, C, e)
        ops.l2norm_bwd(cc["dt32"], cc["t32"], cc["t_inv"], cc["dtfeat16"], B * C, e)
        tt = self.text_encoder.tower(dev)
        dx0 = tt.backward_to_input(cc["dtfeat16"], B * C, Lk, cc["eot_rows"])
        if dx0.dtype != torch.float16:
            raise ops._lib.MvlptError("the CoCoOp branch needs the fp16 gradient stream (unset MVLPT_GRAD_STREAM_F32)")
        ops.cocoop_ctx_grad(dx0, pl._ctx_pos, cc["part"], views["cocoop_ctx"], cc["d_bias"], B, C, Lk, n, d, inv)
        W1, b1, W2, b2 = [t.detach().contiguous() for t in pl.meta_params()]
        # meta_net gradients in true scale; its gradient w.r.t. the normalised image features joins the (scaled) one of
        # the logits before both go back through the L2 normalisation
        ops.metanet_bwd(cc["d_bias"], hb["i32"], cc["h1"], W1, W2, cc["d_h1"], views["meta_net.linear1.weight"],
                        views["meta_net.linear1.bias"], views["meta_net.linear2.weight"], views["meta_net.linear2.bias"],
                        hb["di32"], self.grad_scale)
        if self._img_train:
            ops.l2norm_bwd(hb["di32"], hb["i32"], hb["i_inv"], hb["difeat16"], B, e)
            it = self.image_encoder.tower(dev)
            if pl.uses_vpt_proj:
                d_vpt, d_deep = pl.vproj().grad_input_views(n_deep)
            else:
                d_vpt = views["vpt_embeddings"].view(v, it.d)
                d_deep = views["vpt_embeddings_deep"] if n_deep is not None else None
            it.backward(hb["difeat16"], B, v, n_deep, d_vpt, d_deep, inv)
            if pl.uses_vpt_proj:
                pl.vproj().backward(views, n_deep)
            if d_deep is None and "vpt_embeddings_deep" in views:
                ops.zero(views["vpt_embeddings_deep"])
        return dict(views)

    def hold_text_features(self, on: bool):
        """While held, the text tower runs at most once more: its features are constant as long as no prompt parameter
        changes (evaluation, trainers/mvlpt.py:989-1088 recomputes them for every batch)."""
        self._txt_hold = bool(on)
        self._txt_held_for = None

    def _text_features(self, dev, ctx, C: int):
        """Text tower forward (class-sharded under data parallelism) + L2 normalisation into the head's buffers, on the
        current stream."""
        pl = self.prompt_learner
        head = self.head(dev)
        sh = self._text_shard(dev)
        tt = self.text_encoder.tower(dev)
        if sh is None:
            txt_feat = tt.forward(pl._emb, ctx, pl._slot, pl._eot_rows, pl.coop_n_ctx, pl.csc, train=self._txt_train)
        else:
            c0, c1 = sh["range"]
            ctx_l = ctx[c0:c1] if (ctx is not None and pl.csc) else ctx
            if c1 > c0:
                local = tt.forward(pl._emb[c0:c1], ctx_l, pl._slot[c0:c1], sh["eot_rows"], pl.coop_n_ctx, pl.csc,
                                   train=self._txt_train)
                sh["send"][:c1 - c0].copy_(local)
            self.dp.all_gather_into(sh["gathered"], sh["send"])
            txt_feat = sh["full"]
            for r, (a, b) in enumerate(sh["ranges"]):  # drop the padding rows of each rank's slab
                if b > a:
                    txt_feat[a:b].copy_(sh["gathered"][r, :b - a])
        head.normalize_text(txt_feat)  # into buffers shared by every batch size (keyed by C)
        self._txt_cache_valid = C if ctx is None else False

    def grad_buffer(self) -> torch.Tensor:
        """Flat fp32 gradient buffer over every trainable prompt tensor, in named_parameters() order.  The engine writes
        prompt gradients straight into views of it, so data-parallel ranks all-reduce ONE tensor."""
        dev = self.prompt_learner._emb.device
        if getattr(self, "_grad_flat", None) is None or self._grad_flat.device != dev:
            names = self._trainables()
            total = sum(p.numel() for _, p in names)
            self._grad_flat = torch.zeros(total, device=dev, dtype=torch.float32)
            self._grad_views, off = {}, 0
            for n, p in names:
                self._grad_views[n] = self._grad_flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
        return self._grad_flat

    def grad_views(self) -> Dict[str, torch.Tensor]:
        self.grad_buffer()
        return self._grad_views

    def _backward_core(self, B: int, dlogits: Optional[torch.Tensor] = None, task=None) -> Dict[str, torch.Tensor]:
        """Head -> towers -> prompt gradients (fp32, unscaled), written into views of grad_buffer().  With dlogits=None,
        dz16 was already produced by the fused cross-entropy kernel."""
        pl = self.prompt_learner
        sh = self._shapes
        C, v, n_deep, Lt = sh["C"], sh["v"], sh["n_deep"], sh["Lt"]
        dev = pl._emb.device
        head = self.head(dev)
        bf = head.buffers(B, C)
        if dlogits is not None:
            t_dev, ranges = self._task_dev(task, dev)
            dl = dlogits.float().contiguous()
            ops.dlogits_prepare(dl, dl.stride(0), t_dev, ranges, bf["dz16"], bf["ldc"], B, C, self.grad_scale)
        if pl.cocoop_ctx is not None:
            return self._backward_cocoop(B)
        head.backward(B, C, need_img=self._img_train, need_txt=self._txt_train)
        inv = 1.0 / self.grad_scale
        views = self.grad_views()
        proj = pl.uses_projection
        grads: Dict[str, torch.Tensor] = {}
        d_ctx = d_vpt = d_deep = None
        if self._img_train:
            it = self.image_encoder.tower(dev)
            if pl.uses_vpt_proj:
                d_vpt, d_deep = pl.vproj().grad_input_views(n_deep)
            elif proj:
                _, d_vpt, d_deep = pl.upt().grad_input_views()
            else:
                d_vpt = views["vpt_embeddings"].view(v, it.d)
                d_deep = views["vpt_embeddings_deep"] if n_deep is not None else None
            it.backward(bf["difeat16"], B, v, n_deep, d_vpt, d_deep, inv)
            if pl.uses_vpt_proj:
                pl.vproj().backward(views, n_deep)
        if self._txt_train:
            tt = self.text_encoder.tower(dev)
            if proj:
                d_ctx = pl.upt().grad_input_views()[0]
            else:
                d_ctx = views["ctx"]
            sh = self._text_shard(dev)
            if sh is None:
                tt.backward(bf["dtfeat16"], C, Lt, pl._eot_rows, pl._ctx_pos, pl.coop_n_ctx, pl.csc, d_ctx, inv)
            else:
                # every rank holds d(loss)/d(text features) of ITS images for ALL classes: sum over ranks, keep my classes
                for r, (a, b) in enumerate(sh["ranges"]):
                    if b > a:
                        sh["g_send"][r, :b - a].copy_(bf["dtfeat16"][a:b])
                self.dp.reduce_scatter_sum(sh["g_recv"], sh["g_send"])
                c0, c1 = sh["range"]
                ops.zero(d_ctx)  # class-specific contexts: the other ranks' rows stay zero for the final all-reduce
                if c1 > c0:
                    tt.backward(sh["g_recv"][:c1 - c0], c1 - c0, Lt, sh["eot_rows"], pl._ctx_pos[c0:c1], pl.coop_n_ctx,
                                pl.csc, d_ctx[c0:c1] if pl.csc else d_ctx, inv)
        if proj:
            pl.upt().backward(views)
            grads = dict(views)
        else:
            for k in ("ctx", "vpt_embeddings", "vpt_embeddings_deep", "vpt_proj.weight", "vpt_proj.bias"):
                if k in views:
                    grads[k] = views[k]
            if d_deep is None and "vpt_embeddings_deep" in grads:
                ops.zero(grads["vpt_embeddings_deep"])
        return grads

    # ---- public API ------------------------------------------------------------------------------------------------
    def forward(self, image, task=None):
        """trainers/mvlpt.py:540-583 -> logits [B, n_cls] (differentiable w.r.t. the prompt tensors)."""
        params = [p for _, p in self._trainables()]
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _CustomCLIPFn.apply(self, image, task, train, *params)

    def loss_and_grads(self, image, label, task=None, global_batch: Optional[int] = None):
        """Fused train-step core (forward + cross-entropy + backward) without autograd: returns
        (loss_rows fp32 [B], pred int32 [B], grads {name: fp32 tensor}).  `global_batch` is the divisor of the mean
        (data-parallel ranks pass the global batch so that gradient all-reduce is a plain sum)."""
        B = image.shape[0]
        pl = self.prompt_learner
        logits = self._forward_core(image, task, train=True)
        dev = image.device
        head = self.head(dev)
        bf = head.buffers(B, pl.n_cls)
        t_dev, ranges = self._task_dev(task, dev)
        lab = soft = None
        if label.dim() > 1 and label.shape[-1] > 1:
            soft = label.to(device=dev, dtype=torch.float32).contiguous()
        else:
            lab = label.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        coef = self.grad_scale / float(global_batch or B)
        ops.ce_fwd_bwd(logits, bf["ldc"], lab, soft, t_dev, ranges, bf["loss_rows"], bf["pred"], bf["dz16"], B, pl.n_cls,
                       coef, hit=bf["hit"])
        ops.step_metrics(bf["loss_rows"], bf["hit"], B, 1.0 / B, bf["metrics"])
        grads = self._backward_core(B)
        return bf["loss_rows"], bf["pred"], grads

    def last_metrics(self, B: int) -> torch.Tensor:
        """Device fp32 [2] = (batch-mean loss, top-1 accuracy %) of the most recent loss_and_grads call."""
        return self.head(self.prompt_learner._emb.device).buffers(B, self.prompt_learner.n_cls)["metrics"]

    def last_logits(self, B: int) -> torch.Tensor:
        """fp32 logits [B, n_cls] of the most recent pass (a view into the engine buffer)."""
        C = self.prompt_learner.n_cls
        return self.head(self.prompt_learner._emb.device).buffers(B, C)["logits"][:, :C]


class LossSummary(dict):
    """`forward_backward`'s return value: {"loss", "acc"[, "num_tasks"]} (trainers/mvlpt.py:940-946).  A dict whose loss
    and accuracy are filled from the pinned host slot on first use (any read synchronises with the copy of ITS step,
    not with the steps enqueued since), so a training loop that looks at the summary every few steps never stalls
    the host in between.  The slot is reused after 16 further steps: read a summary before then (or keep `.resolve()`)."""

    def __init__(self, host, event, extra):
        super().__init__(loss=None, acc=None, **extra)
        self._host, self._event = host, event

    def resolve(self) -> "LossSummary":
        if self._event is not None:
            self._event.synchronize()
            dict.__setitem__(self, "loss", float(self._host[0]))
            dict.__setitem__(self, "acc", float(self._host[1]))
            self._event = None
        return self

    def __getitem__(self, k):
        return dict.__getitem__(self.resolve(), k)

    def get(self, k, default=None):
        return dict.get(self.resolve(), k, default)

    def items(self):
        return dict.items(self.resolve())

    def keys(self):
        return dict.keys(self.resolve())

    def __iter__(self):  # also takes dict(summary) / {**summary} off CPython's raw-table fast path, which would copy None
        return dict.__iter__(self.resolve())

    def copy(self):
        return dict(dict.items(self.resolve()))

    def values(self):
        return dict.values(self.resolve())

    def __repr__(self):
        return dict.__repr__(self.resolve())

    def __eq__(self, other):
        return dict.__eq__(self.resolve(), other)

    __hash__ = None


# ---------------------------------------------------------------------------------------------------------------
# Trainer (trainers/mvlpt.py:827-1125).  The reference subclasses Dassl's TrainerX; Dassl is not installable here,
# so the handful of inherited members it uses are provided by this class itself.
# ---------------------------------------------------------------------------------------------------------------
from .runtime import TRAINER_REGISTRY  # noqa: E402


@TRAINER_REGISTRY.register()
class MVLPT:
    """Drop-in for the reference's `MVLPT(TrainerX)`: same cfg keys, same method names, same batch formats, same
    `forward_backward -> {"loss", "acc"[, "num_tasks"]}` contract.  One process per GPU; with torch.distributed
    initialised the prompt gradients are SUM-all-reduced over NCCL (replaces nn.DataParallel, :877-880)."""

    def __init__(self, cfg, dm=None, clip_state_dict=None, device=None, tokenized_prompts=None, name_lens=None,
                 dp=None):
        from . import runtime as R
        self.check_cfg(cfg)
        self.cfg = cfg
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise ops._lib.MvlptError("MVLPT trainer needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is not None \
                and self.device.index != torch.cuda.current_device():
            # the library launches on the CURRENT device, on its current stream (ops._stream): one process drives one GPU
            torch.cuda.set_device(self.device)
        self.dp = dp if dp is not None else R.DataParallelGroup()
        self._models, self._optims, self._scheds = OrderedDict(), OrderedDict(), OrderedDict()
        self.epoch, self.start_epoch = 0, 0
        self.max_epoch = cfg.OPTIM.MAX_EPOCH
        self.output_dir = getattr(cfg, "OUTPUT_DIR", "")
        self.best_result = -float("inf")
        self.batch_idx, self.num_batches = 0, 1
        self._clip_state_dict = clip_state_dict
        self._tok = (tokenized_prompts, name_lens)
        self.dm = dm
        self.build_data_loader()
        self.build_model()

    # ---- reference API -------------------------------------------------------------------------------------------
    def check_cfg(self, cfg):
        assert cfg.TRAINER.MVLPT.PREC in ["fp16", "fp32", "amp"]

    def build_data_loader(self):
        """trainers/mvlpt.py:883-908.  The reference builds one of three Dassl/ELEVATER data managers from files on
        disk (out of scope); here the caller passes any object exposing the same attributes."""
        dm = self.dm
        self.multi_task = self.cfg.DATASET.MULTITASK
        self.multi_task_label_pertask = self.cfg.DATASET.MULTITASK_LABEL_PERTASK
        if dm is None:
            raise ValueError("pass dm= (an object with the DataManager attributes: dataset.classnames | lab2cname, "
                             "num_classes, train_loader_x, val_loader, test_loader)")
        self.train_loader_x = getattr(dm, "train_loader_x", None)
        self.train_loader_u = getattr(dm, "train_loader_u", None)
        self.val_loader = getattr(dm, "val_loader", None)
        self.test_loader = getattr(dm, "test_loader", None)
        self.num_classes = dm.num_classes
        self.num_source_domains = getattr(dm, "num_source_domains", 1)
        self.lab2cname = dm.lab2cname
        if self.train_loader_x is not None and hasattr(self.train_loader_x, "__len__"):
            self.num_batches = len(self.train_loader_x)

    def build_model(self):
        """trainers/mvlpt.py:838-880."""
        from . import runtime as R
        cfg = self.cfg
        classnames = self.dm.dataset.classnames if cfg.DATASET.COOP else list(self.dm.lab2cname.values())
        clip_model = load_clip_to_cpu(cfg, self._clip_state_dict)
        if cfg.TRAINER.MVLPT.PREC in ("fp32", "amp"):
            clip_model.float()
        self.model = CustomCLIP(cfg, classnames, clip_model, dm=self.dm, tokenized_prompts=self._tok[0],
                                name_lens=self._tok[1])
        # only prompt_learner.* is trainable; the CLIP towers are frozen kernel inputs, not nn.Parameters (:856-858)
        for name, param in self.model.named_parameters():
            param.requires_grad_("prompt_learner" in name)
        if getattr(cfg.MODEL, "INIT_WEIGHTS", ""):
            sd = torch.load(cfg.MODEL.INIT_WEIGHTS, map_location="cpu")
            self.model.prompt_learner.load_state_dict(sd.get("state_dict", sd), strict=False)
        self.model.to(self.device)
        self.model.dp = self.dp
        self.optim = R.build_optimizer(self.model._trainables(), cfg.OPTIM)
        self.sched = R.build_lr_scheduler(self.optim, cfg.OPTIM)
        self.register_model("prompt_learner", self.model.prompt_learner, self.optim, self.sched)
        self.scaler = None  # amp: the kernels already scale gradients by a fixed power of two (engine.py)
        # loss / accuracy of a step travel to a ring of pinned host slots by asynchronous copies; the host only waits for
        # one when somebody reads the summary (LossSummary), so it can enqueue the next steps meanwhile
        self._metrics_host = torch.zeros(16, 2, dtype=torch.float32).pin_memory()
        self._metrics_events = [torch.cuda.Event() for _ in range(16)]
        self._metrics_slot = 0

    def forward_backward(self, batch):
        """trainers/mvlpt.py:910-951: forward, cross-entropy, backward, SGD step; returns the loss summary."""
        image, label, tasks_ = self.parse_batch_train(batch)
        B = image.shape[0]
        model = self.model
        # soft (multi-hot) labels are row-normalised inside the cross-entropy kernel (:914-916)
        model.loss_and_grads(image, label, tasks_, global_batch=B * self.dp.world)
        flat = model.grad_buffer()
        self.dp.all_reduce_sum(flat)
        self.optim.step(flat)
        slot = self._metrics_slot
        self._metrics_slot = (slot + 1) % len(self._metrics_events)
        host = self._metrics_host[slot]
        host.copy_(model.last_metrics(B), non_blocking=True)  # the device -> host read of the step's result
        self._metrics_events[slot].record(torch.cuda.current_stream(self.device))
        extra = {"num_tasks": len(set(tasks_.tolist()))} if tasks_ is not None else {}
        if (self.batch_idx + 1) == self.num_batches:
            self.update_lr()
        # the reference's two .item() calls (:939-942) stall the host every step; here the summary waits for the copy when
        # (and only when) it is read — Dassl prints it every TRAIN.PRINT_FREQ steps
        return LossSummary(host, self._metrics_events[slot], extra)

    def _parse(self, batch):
        if self.cfg.DATASET.COOP:
            inp_key, lab_key, task_key = "img", "label", "domain"
        else:
            inp_key, lab_key, task_key = 0, 1, 3
        input, label = batch[inp_key], batch[lab_key]
        tasks = batch[task_key] if self.multi_task else None
        input = self._stage_input(input)
        label = label.to(self.device, non_blocking=True)
        return input, label, tasks

    def _stage_input(self, input: torch.Tensor, label: Optional[torch.Tensor] = None):
        """Host->device copy of the image batch.  A pinned host batch is copied on a dedicated copy stream into one of two
        persistent device buffers and tagged with the event that ends the copy: CustomCLIP runs the (image-independent)
        text tower first and only then waits for it, so the PCIe transfer hides behind compute; with `stage_batch` called
        one step ahead (run_epoch does) it hides behind the whole previous step.  Anything else takes the reference's plain
        `.to(device)` route (trainers/mvlpt.py:959-960)."""
        u8 = input.dtype == torch.uint8
        if input.device.type != "cpu" or not input.is_pinned():
            input = input.to(self.device, non_blocking=True)
            return self._normalize_u8(input) if u8 else input
        key = (tuple(input.shape), input.dtype)
        if getattr(self, "_stage_key", None) != key:
            self._stage_key = key
            self._stage_bufs = [torch.empty(input.shape, dtype=input.dtype, device=self.device) for _ in range(2)]
            self._stage_norm = [torch.empty(input.shape, dtype=self._image_dtype(), device=self.device) for _ in range(2)] \
                if u8 else None
            self._stage_slot = 0
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        slot = self._stage_slot
        buf = self._stage_bufs[slot]
        self._stage_slot ^= 1
        # everything enqueued so far — in particular the step that last read this buffer — is done before it is overwritten
        cs.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cs):
            buf.copy_(input, non_blocking=True)
            if u8:  # one byte per value crossed PCIe; ToTensor + Normalize happen here, still on the copy stream
                buf = self._normalize_u8(buf, self._stage_norm[slot])
            if label is not None and label.device.type == "cpu" and label.is_pinned():
                # into a persistent slot like the images: a fresh allocation per step on the copy stream makes the caching
                # allocator grow (record_stream delays reuse) and every cudaMalloc it then needs stalls the device
                lkey = (tuple(label.shape), label.dtype)
                if getattr(self, "_stage_lkey", None) != lkey:
                    self._stage_lkey = lkey
                    self._stage_labels = [torch.empty(label.shape, dtype=label.dtype, device=self.device) for _ in range(2)]
                label = self._stage_labels[slot].copy_(label, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        buf._mvlpt_ready = ev
        return buf if label is None else (buf, label)

    def _image_dtype(self):
        return torch.float16 if self.cfg.TRAINER.MVLPT.PREC == "fp16" else torch.float32

    def _normalize_u8(self, img_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [B,3,H,W] batches (images already at INPUT.SIZE: the loader kept them as bytes) -> ToTensor + Normalize with
        INPUT.PIXEL_MEAN / PIXEL_STD on the device (`mvlpt_normalize_u8`, bit-identical to torchvision), on the current
        stream.  The reference ships fp32 tensors from its CPU transform stack (trainers/mvlpt.py:959-960)."""
        if out is None:
            out = torch.empty(img_u8.shape, dtype=self._image_dtype(), device=img_u8.device)
        return ops.normalize_u8(img_u8.contiguous(), out, self.cfg.INPUT.PIXEL_MEAN, self.cfg.INPUT.PIXEL_STD)

    def stage_batch(self, batch):
        """Start the host->device copies of a batch AHEAD of its step (pinned host tensors; anything else is returned as
        is).  Returns a batch of the same structure whose image (and label) live on the device; forward_backward /
        parse_batch_* accept it like any other batch.  At most one batch may be staged ahead of the one being computed."""
        if self.cfg.DATASET.COOP:
            inp_key, lab_key = "img", "label"
        else:
            inp_key, lab_key = 0, 1
        input, label = batch[inp_key], batch[lab_key]
        if not (isinstance(input, torch.Tensor) and input.device.type == "cpu" and input.is_pinned()):
            return batch
        staged = self._stage_input(input, label)
        out = dict(batch) if isinstance(batch, dict) else list(batch)
        out[inp_key], out[lab_key] = staged
        return out if isinstance(batch, dict) else tuple(out)

    def parse_batch_train(self, batch):
        return self._parse(batch)

    def parse_batch_test(self, batch):
        return self._parse(batch)

    @torch.no_grad()
    def model_inference(self, input, task=None):
        return self.model(input, task=task)

    @torch.no_grad()
    def test(self, split=None):
        """trainers/mvlpt.py:989-1088.  Same flow and the same numbers; what differs is where the work happens: the text
        features are computed once for the whole evaluation (they are constant while no parameter moves), logits stay on
        the device until the loader is exhausted, and the reference's per-sample Python scatter into per-task evaluators
        (:1030-1042) is one boolean mask per task.  CoOp-style datasets (cfg.DATASET.COOP) report Dassl's classification
        accuracy in percent; ELEVATER-style ones the task's own metric (accuracy / mean-per-class / 11point_mAP / roc_auc,
        `dm._metric_name`, `dm._metric`, else trainers/metrics.py).  Returns the first value of the final `results` dict
        like the reference; the dicts are kept in `self.last_test_results`."""
        import numpy as np
        from . import metrics as MT
        self.set_model_mode("eval")
        split = split or self.cfg.TEST.SPLIT
        if split == "val" and self.val_loader is not None:
            loader = self.val_loader
        else:
            split, loader = "test", self.test_loader
        coop = bool(self.cfg.DATASET.COOP)
        outs, labs, tasks = [], [], []
        self.model.hold_text_features(True)  # constant during evaluation: one text-tower pass for the whole loader
        try:
            for batch in loader:
                image, label, tasks_ = self.parse_batch_test(batch)
                outs.append(self.model_inference(image, task=tasks_).float())
                labs.append(label)
                if tasks_ is not None:
                    tasks.append(torch.as_tensor(tasks_).reshape(-1).cpu())
        finally:
            self.model.hold_text_features(False)
        out = torch.cat(outs).cpu() if outs else torch.zeros(0, self.num_classes)
        lab = torch.cat([l.cpu() for l in labs]) if labs else torch.zeros(0, dtype=torch.long)
        task_ids = torch.cat(tasks) if tasks else None

        def classification_accuracy(o, l):  # Dassl's Classification evaluator: percent of arg-max hits
            if l.dim() > 1 and l.shape[-1] > 1:
                l = l.argmax(dim=1)
            return 100.0 * float((o.argmax(dim=1) == l).float().mean()) if len(l) else 0.0

        def metric_of(name, fn, y_true, y_pred):
            fn = fn if fn is not None else MT.get_metric(name)
            if name == "accuracy" and y_true.ndim > 1:
                y_true = np.argmax(y_true, axis=-1)
            return float(fn(y_true, y_pred))

        dm = self.dm
        per_task = {}
        if self.multi_task and task_ids is not None:
            for tid in sorted(set(task_ids.tolist())):
                name = dm._id2task[tid] if hasattr(dm, "_id2task") else tid
                sel = task_ids == tid
                o, l = out[sel], lab[sel]
                if hasattr(dm, "_task_class_idx"):  # evaluate on the task's own classes (:1036-1040, :1052-1054)
                    c0, c1 = dm._task_class_idx[name]
                    o = o[:, c0:c1]
                    l = l[:, c0:c1] if l.dim() > 1 else l - c0
                if coop:
                    per_task[name] = classification_accuracy(o, l)
                else:
                    mname = dm._metric_name[name]
                    per_task[name] = metric_of(mname, getattr(dm, "_metric", {}).get(name), l.numpy(), o.numpy())
        if per_task:
            key = self.cfg.DATASET.MULTITASK_EVALKEY
            if key == "average":
                results = {"average": sum(per_task.values()) / len(per_task)}
            else:
                assert key in per_task, key
                results = {key: per_task[key]}
        elif coop or not hasattr(dm, "_metric_name"):
            results = {"accuracy": classification_accuracy(out, lab)}
        else:
            results = {dm._metric_name: metric_of(dm._metric_name, getattr(dm, "_metric", None), lab.numpy(), out.numpy())}
        self.last_test_results = dict(split=split, results=results, per_task=per_task)
        return list(results.values())[0]

    def load_model(self, directory, epoch=None):
        """trainers/mvlpt.py:1090-1125: <dir>/prompt_learner/model-best.pth.tar | model.pth.tar-<epoch>; renames
        upt_proj -> mvlpt_proj, drops token_prefix/token_suffix, loads non-strictly."""
        import os.path as osp
        if not directory:
            print("Note that load_model() is skipped as no pretrained model is given")
            return
        model_file = "model-best.pth.tar" if epoch is None else "model.pth.tar-" + str(epoch)
        for name in self.get_model_names():
            model_path = osp.join(directory, name, model_file)
            if not osp.exists(model_path):
                raise FileNotFoundError('Model not found at "{}"'.format(model_path))
            checkpoint = torch.load(model_path, map_location="cpu", weights_only=False)
            state_dict = {k.replace("upt_proj", "mvlpt_proj"): v for k, v in checkpoint["state_dict"].items()}
            for k in ("token_prefix", "token_suffix"):
                state_dict.pop(k, None)
            print('Loading weights to {} from "{}" (epoch = {})'.format(name, model_path, checkpoint.get("epoch")))
            self._models[name].load_state_dict(state_dict, strict=False)
            self.model._txt_cache_valid = False

    # ---- members the reference inherits from Dassl's TrainerX ------------------------------------
    def register_model(self, name="model", model=None, optim=None, sched=None):
        self._models[name], self._optims[name], self._scheds[name] = model, optim, sched

    def get_model_names(self, names=None):
        return list(self._models.keys()) if names is None else list(names)

    def set_model_mode(self, mode="train", names=None):
        for name in self.get_model_names(names):
            self._models[name].train(mode == "train")

    def update_lr(self, names=None):
        for name in self.get_model_names(names):
            if self._scheds[name] is not None:
                self._scheds[name].step()

    def get_current_lr(self, names=None):
        return self._optims[self.get_model_names(names)[0]].param_groups[0]["lr"]

    def save_model(self, epoch, directory, is_best=False, val_result=None, model_name=""):
        """Dassl's TrainerBase.save_model -> save_checkpoint (upstream): `epoch` is the 0-based epoch just finished; the file
        is <dir>/<name>/model.pth.tar-<epoch+1> (or `model_name`), the dict stores epoch + 1, and a text file `checkpoint`
        next to it names the newest one (what --resume reads).  Keys: state_dict / epoch / optimizer / scheduler /
        val_result (scripts/avg_ckpt.py:21-66 reads the same)."""
        import os
        import os.path as osp
        for name in self.get_model_names():
            sched = self._scheds[name]
            ckpt = {"state_dict": {k: v.detach().cpu() for k, v in self._models[name].state_dict().items()},
                    "epoch": epoch + 1, "optimizer": self._optims[name].state_dict(),
                    "scheduler": None if sched is None else sched.state_dict(), "val_result": val_result}
            save_dir = osp.join(directory, name)
            os.makedirs(save_dir, exist_ok=True)
            fname = model_name or f"model.pth.tar-{epoch + 1}"
            torch.save(ckpt, osp.join(save_dir, fname))
            with open(osp.join(save_dir, "checkpoint"), "w") as f:
                f.write(fname + "\n")
            if is_best:
                torch.save(ckpt, osp.join(save_dir, "model-best.pth.tar"))

    def resume_model_if_exist(self, directory) -> int:
        """Dassl's TrainerBase.resume_model_if_exist (upstream; reached through --resume, train.py:55-56): if every
        registered model has a `checkpoint` file under <directory>/<name>/, restore parameters, optimiser and scheduler
        from the checkpoint it names and return the epoch to continue from; else 0."""
        import os.path as osp
        names = self.get_model_names()
        if not directory or not all(osp.exists(osp.join(directory, n, "checkpoint")) for n in names):
            if directory:
                print("No checkpoint found, train from scratch")
            return 0
        start = 0
        for name in names:
            with open(osp.join(directory, name, "checkpoint")) as f:
                fname = f.readline().strip()
            path = osp.join(directory, name, fname)
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            print('Loading checkpoint from "{}"'.format(path))
            self._models[name].load_state_dict(ckpt["state_dict"], strict=False)
            self._optims[name].load_state_dict(ckpt["optimizer"])
            for k, v in self._optims[name].bufs.items():
                self._optims[name].bufs[k] = v.to(self.device)
            if ckpt.get("scheduler") is not None and self._scheds[name] is not None:
                self._scheds[name].load_state_dict(ckpt["scheduler"])
            start = int(ckpt["epoch"])
            print("Previous epoch: {}".format(start))
        self.model._txt_cache_valid = False
        return start

    def run_epoch(self):
        self.set_model_mode("train")
        self.num_batches = len(self.train_loader_x)
        last = None
        it = iter(self.train_loader_x)
        nxt = next(it, None)
        nxt = None if nxt is None else self.stage_batch(nxt)
        self.batch_idx = -1
        while nxt is not None:
            batch, self.batch_idx = nxt, self.batch_idx + 1
            nxt = next(it, None)
            if nxt is not None:
                nxt = self.stage_batch(nxt)  # the next batch crosses PCIe while this one is computed
            last = self.forward_backward(batch)
        return last

    # ---- Dassl's epoch loop (SimpleTrainer.before_train / after_epoch / after_train, upstream; train.py:219) -----------
    def before_train(self):
        directory = getattr(self.cfg, "RESUME", "") or self.output_dir
        self.start_epoch = self.resume_model_if_exist(directory)

    def after_epoch(self):
        last_epoch = (self.epoch + 1) == self.max_epoch
        do_test = not getattr(self.cfg.TEST, "NO_TEST", False)
        freq = getattr(self.cfg.TRAIN, "CHECKPOINT_FREQ", 0)
        meet_checkpoint_freq = (self.epoch + 1) % freq == 0 if freq > 0 else False
        if do_test and self.cfg.TEST.FINAL_MODEL == "best_val" and self.val_loader is not None:
            curr_result = self.test(split="val")
            if curr_result > self.best_result:
                self.best_result = curr_result
                if self.output_dir:
                    self.save_model(self.epoch, self.output_dir, val_result=curr_result, model_name="model-best.pth.tar")
        if (meet_checkpoint_freq or last_epoch) and self.output_dir:
            self.save_model(self.epoch, self.output_dir)

    def after_train(self):
        print("Finish training")
        if getattr(self.cfg.TEST, "NO_TEST", False) or self.test_loader is None:
            return None
        if self.cfg.TEST.FINAL_MODEL == "best_val" and self.output_dir and self.val_loader is not None:
            print("Deploy the model with the best val performance")
            self.load_model(self.output_dir)
        else:
            print("Deploy the last-epoch model")
        return self.test()

    def train(self):
        """Dassl's TrainerBase.train (upstream): before_train, then per epoch run_epoch + after_epoch, then after_train.
        Returns the last step's loss summary."""
        last = None
        self.before_train()
        for self.epoch in range(self.start_epoch, self.max_epoch):
            last = self.run_epoch()
            self.after_epoch()
        self.final_result = self.after_train()
        return last


# When Dassl is importable (an MVLPT checkout with its dependencies), put this trainer into ITS registry as well, under
# the reference's name (`@TRAINER_REGISTRY.register() class MVLPT(TrainerX)`, trainers/mvlpt.py:827-828), so that
# `--trainer MVLPT` in the reference's own train.py resolves here.  If the reference's trainers/mvlpt.py was imported
# first the name is taken: leave it.
try:  # pragma: no cover - dassl is not installable in this image
    from dassl.engine import TRAINER_REGISTRY as _DASSL_REGISTRY
    try:
        _DASSL_REGISTRY.register()(MVLPT)
    except (KeyError, AssertionError):
        pass
except ImportError:
    pass
