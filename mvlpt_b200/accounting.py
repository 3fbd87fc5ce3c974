"""Algorithmic FLOP / byte accounting of the hot path — used by bench.py and DESIGN.md's roofline."""
from __future__ import annotations


def tower_fwd_flops(L: int, d: int, layers: int) -> float:
    """Per sequence: four linears (24·L·d²) + QKᵀ and PV (4·L²·d) per block; full L² counted for causal too."""
    return float(layers * (24 * L * d * d + 4 * L * L * d))


def tower_bwd_flops(L: int, d: int, layers: int) -> float:
    """dgrad only (frozen weights, no wgrad): four dgrads (24·L·d²) + dV, dP, dQ, dK (8·L²·d); recompute not counted."""
    return float(layers * (24 * L * d * d + 8 * L * L * d))


def text_flops(arch: dict, C: int, L_t: int, backward: bool) -> float:
    """Text tower over C class prompts of L_t rows: forward (+ projection), plus the dgrad-only backward if asked."""
    e, dt, lt = arch["embed_dim"], arch["transformer_width"], arch["transformer_layers"]
    f = C * (tower_fwd_flops(L_t, dt, lt) + 2 * dt * e)
    if backward:
        f += C * tower_bwd_flops(L_t, dt, lt)
    return float(f)


def flops_step(arch: dict, B: int, C: int, L_t: int, v: int, n_ctx: int, text_passes: int = 1,
               head_classes: int | None = None) -> float:
    """One train step: image fwd (+bwd when visual prompts train), text fwd (+bwd when context trains), head.
    `text_passes` = B for the CoCoOp branch (every image has its own class prompts), else 1.  `C` is the number of class
    prompts the text tower encodes (0: features held; C/world under class sharding), `head_classes` the width of the logit
    head (defaults to C)."""
    d, ly, p = arch["vision_width"], arch["vision_layers"], arch["vision_patch_size"]
    e = arch["embed_dim"]
    n_p = (arch["image_resolution"] // p) ** 2
    L = 1 + v + n_p
    Ch = C if head_classes is None else head_classes
    f = B * (tower_fwd_flops(L, d, ly) + 2 * n_p * d * 3 * p * p + 2 * d * e)
    f += text_passes * text_flops(arch, C, L_t, n_ctx > 0) + 2 * B * Ch * e
    if v > 0:
        f += B * tower_bwd_flops(L, d, ly) + 2 * B * Ch * e
    if n_ctx > 0:
        f += 2 * B * Ch * e
    return float(f)


def flops_inference(arch: dict, B: int, C: int, v: int) -> float:
    """One evaluation batch with the text features held (MVLPT.test): image tower forward + logit head."""
    d, ly, p = arch["vision_width"], arch["vision_layers"], arch["vision_patch_size"]
    e = arch["embed_dim"]
    n_p = (arch["image_resolution"] // p) ** 2
    L = 1 + v + n_p
    return float(B * (tower_fwd_flops(L, d, ly) + 2 * n_p * d * 3 * p * p + 2 * d * e) + 2 * B * C * e)
