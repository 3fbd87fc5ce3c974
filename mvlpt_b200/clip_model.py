"""Frozen CLIP parameter container with the reference's state-dict layout.

The reference builds an nn.Module tree (clip/model.py:239-293, build_model :395-432) whose forward does the math
in torch.  Here the CLIP towers are frozen inputs to CUDA kernels, so the container only has to (a) accept the
same state dict, (b) answer the handful of attribute look-ups trainers/mvlpt.py makes on `clip_model`
(dtype, visual.conv1.weight, visual.input_resolution, visual.output_dim, ln_final.weight, token_embedding,
context_length, logit_scale, state_dict()).  The device copies in kernel layout are made by mvlpt_b200.engine.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict

import torch


class FrozenCLIP:
    def __init__(self, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float16):
        if "visual.proj" not in state_dict:
            raise ValueError("only the ViT backbones are supported (the reference's ImageEncoder assumes a vision "
                             "transformer, trainers/mvlpt.py:48)")
        sd = OrderedDict((k, v.detach()) for k, v in state_dict.items()
                         if k not in ("input_resolution", "context_length", "vocab_size"))
        self._sd = sd
        self._dtype = dtype
        conv = sd["visual.conv1.weight"]
        grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        self.visual = SimpleNamespace(
            conv1=SimpleNamespace(weight=conv),
            input_resolution=conv.shape[-1] * grid,
            output_dim=sd["visual.proj"].shape[1],
            proj=sd["visual.proj"],
        )
        self.ln_final = SimpleNamespace(weight=sd["ln_final.weight"], bias=sd["ln_final.bias"])
        self.context_length = sd["positional_embedding"].shape[0]
        self.vocab_size = sd["token_embedding.weight"].shape[0]
        self.logit_scale = sd["logit_scale"]
        self.positional_embedding = sd["positional_embedding"]
        self.text_projection = sd["text_projection"]

    # -- the look-ups trainers/mvlpt.py performs on clip_model -------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    def float(self) -> "FrozenCLIP":
        """`clip_model.float()` for PREC=fp32/amp (trainers/mvlpt.py:848-850): prompt tensors become fp32; the frozen
        towers still run the fp16-operand / fp32-accumulate kernels."""
        self._dtype = torch.float32
        return self

    def eval(self) -> "FrozenCLIP":
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self._sd

    def token_embedding(self, token_ids: torch.Tensor) -> torch.Tensor:
        """nn.Embedding lookup (clip/model.py:286), init-time only."""
        return self._sd["token_embedding.weight"][token_ids.cpu()]

    @property
    def vision_layers(self) -> int:
        return len([k for k in self._sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])


def build_model(state_dict: Dict[str, torch.Tensor]) -> FrozenCLIP:
    """Same call as clip.build_model (clip/model.py:395-432): fp16 model from a state dict."""
    return FrozenCLIP(state_dict, dtype=torch.float16)


def as_state_dict(clip_model) -> Dict[str, torch.Tensor]:
    """Accept either a FrozenCLIP or the reference's own clip.model.CLIP instance."""
    return clip_model.state_dict()
