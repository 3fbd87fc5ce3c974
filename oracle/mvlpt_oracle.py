"""CPU ORACLE for the MVLPT prompt-tuning hot path — TEST INFRASTRUCTURE ONLY.

This is a from-scratch, functional restatement (plain torch tensor ops on the CPU, fp32 by default) of the
arithmetic the reference performs in clip/model.py + trainers/mvlpt.py.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may import it, and only as the checker or the CPU
baseline — the product path (mvlpt_b200/) never does, and raises if its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the REFERENCE ITSELF: oracle/gen_golden.py imports /root/reference (with in-memory stubs for the
absent dassl/ftfy packages), runs its CustomCLIP forward/backward on deterministic synthetic weights
(mvlpt_b200/synth.py) and commits the outputs under tests/golden/.  tests/test_oracle_golden.py checks every
function below against those fixtures.

Every function cites the reference lines it restates (paths relative to the reference root).
Activations are batch-first [N, L, d]; the reference's [L, N, d] permutes (clip/model.py:227,229) are layout
only and change no value.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------- primitives
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """clip/model.py:153-159 — LayerNorm evaluated in fp32 whatever the input dtype, biased variance."""
    xf = x.to(torch.float32) if x.dtype == torch.float16 else x
    mu = xf.mean(dim=-1, keepdim=True)
    var = ((xf - mu) ** 2).mean(dim=-1, keepdim=True)
    y = (xf - mu) * torch.rsqrt(var + eps) * w.to(xf.dtype) + b.to(xf.dtype)
    return y.to(x.dtype)


def quick_gelu(t: Tensor) -> Tensor:
    """clip/model.py:162-164."""
    return t * torch.sigmoid(1.702 * t)


def attention(h: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, heads: int, causal: bool) -> Tensor:
    """clip/model.py:171,181-183 -> torch F.multi_head_attention_forward (upstream): packed in_proj
    (rows [0,d)=Q, [d,2d)=K, [2d,3d)=V), per-head width d/heads, q scaled by hd^-1/2, additive -inf mask strictly
    above the diagonal for the text tower (clip/model.py:324-330), softmax over keys, out_proj."""
    n, L, d = h.shape
    hd = d // heads
    qkv = h @ w_in.t() + b_in
    q, k, v = qkv.split(d, dim=-1)
    q = q.reshape(n, L, heads, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.reshape(n, L, heads, hd).transpose(1, 2)
    v = v.reshape(n, L, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if causal:
        mask = torch.full((L, L), float("-inf"), dtype=s.dtype).triu(1)
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(n, L, d)
    return o @ w_out.t() + b_out


def residual_block(x: Tensor, sd: Dict[str, Tensor], prefix: str, heads: int, causal: bool) -> Tensor:
    """clip/model.py:185-188 — x += attn(ln_1(x)); x += c_proj(quickgelu(c_fc(ln_2(x))))."""
    g = lambda k: sd[prefix + k].to(x.dtype)
    h = layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"])
    x = x + attention(h, g("attn.in_proj_weight"), g("attn.in_proj_bias"), g("attn.out_proj.weight"),
                      g("attn.out_proj.bias"), heads, causal)
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    t = h @ g("mlp.c_fc.weight").t() + g("mlp.c_fc.bias")
    x = x + quick_gelu(t) @ g("mlp.c_proj.weight").t() + g("mlp.c_proj.bias")
    return x


def _n_layers(sd: Dict[str, Tensor], prefix: str) -> int:
    return len([k for k in sd if k.startswith(prefix) and k.endswith(".attn.in_proj_weight")])


# ----------------------------------------------------------------------------------------------- image tower
def patch_embed(image: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """trainers/mvlpt.py:53-58 (= clip/model.py:220-225): stride-p conv without bias == per-patch linear over
    (c, ky, kx); prepend class embedding; add positional embedding; ln_pre.  Returns [B, 1+g*g, d]."""
    w = sd["visual.conv1.weight"].to(image.dtype)  # [d, 3, p, p]
    d, _, p, _ = w.shape
    B, C, H, W = image.shape
    g = H // p
    patches = image.reshape(B, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * p * p)
    x = patches @ w.reshape(d, -1).t()
    cls = sd["visual.class_embedding"].to(x.dtype).expand(B, 1, d)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"].to(x.dtype)
    return layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])


def image_tower(image: Tensor, sd: Dict[str, Tensor], vpt: Optional[Tensor] = None,
                vpt_deep: Optional[Tensor] = None, vpt_proj: Optional[Sequence[Tensor]] = None,
                drop_p: float = 0.0, drop_keep: Optional[Sequence[Tensor]] = None) -> Tensor:
    """trainers/mvlpt.py:52-93 with forward_vpt (:416-437).  `vpt` [1,v,p]: rows inserted after the class token
    (no positional embedding, no ln_pre).  `vpt_deep` [layers-1,v,p]: before block l>=1 rows 1..v are overwritten
    with vpt_deep[l-1]; blocks with l > vpt_deep.shape[0] are SKIPPED (reference quirk, :73).
    `vpt_proj` = (weight [d,p], bias [d]) of the Linear the reference builds when VPT.PROJECT > -1 (:170-175), applied
    to each slab before the batch expansion (:76-77, :425); None = Identity (the default).
    `drop_p`, `drop_keep`: vpt_dropout (:165) in training mode with the keep masks GIVEN — drop_keep[l] is a bool
    [B,v,d] tensor for the slab entering block l (l=0: the shallow prompts): rows = slab * keep / (1 - p), exactly what
    nn.Dropout computes once its Bernoulli draw is fixed."""
    heads = sd["visual.conv1.weight"].shape[0] // 64
    x = patch_embed(image, sd)
    B = x.shape[0]
    v = 0

    def rows(slab: Tensor, l: int) -> Tensor:
        if vpt_proj is not None:
            slab = slab @ vpt_proj[0].to(slab.dtype).t() + vpt_proj[1].to(slab.dtype)
        slab = slab.to(x.dtype).expand(B, -1, -1)
        if drop_p > 0.0 and drop_keep is not None:
            slab = slab * drop_keep[l].to(x.dtype) / (1.0 - drop_p)
        return slab

    if vpt is not None:
        v = vpt.shape[1]
        x = torch.cat([x[:, :1], rows(vpt.reshape(v, -1), 0), x[:, 1:]], dim=1)
    layers = _n_layers(sd, "visual.")
    for l in range(layers):
        if vpt_deep is not None and l >= 1:
            if l > vpt_deep.shape[0]:
                continue
            x = torch.cat([x[:, :1], rows(vpt_deep[l - 1], l), x[:, 1 + v:]], dim=1)
        x = residual_block(x, sd, f"visual.transformer.resblocks.{l}.", heads, causal=False)
    c = layer_norm(x[:, 0], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return c @ sd["visual.proj"].to(x.dtype)


# ----------------------------------------------------------------------------------------------- text tower
def coop_prompts(embedding: Tensor, ctx: Optional[Tensor], name_lens: Sequence[int], n_ctx: int,
                 position: str = "end") -> Tensor:
    """trainers/mvlpt.py:439-515.  `embedding` [C, L_t, d_t] is token_embedding("X"*n_ctx + " name.") — i.e.
    cat(token_prefix, placeholder rows, token_suffix) (:307-316).  Rows 1..n_ctx are placeholders replaced by ctx:
    end:    [SOS, ctx, name . EOS pad]
    middle: [SOS, ctx[:n/2], name, ctx[n/2:], . EOS pad]
    front:  [SOS, name, ctx, . EOS pad]
    ctx is [n,d_t] (shared) or [C,n,d_t] (class-specific, CSC)."""
    if ctx is None:
        return embedding
    C = embedding.shape[0]
    ctx = ctx.to(embedding.dtype)
    if ctx.dim() == 2:
        ctx = ctx.unsqueeze(0).expand(C, -1, -1)
    prefix, suffix = embedding[:, :1], embedding[:, 1 + n_ctx:]
    if position == "end":
        return torch.cat([prefix, ctx, suffix], dim=1)
    rows = []
    half = n_ctx // 2
    for c in range(C):
        nl = name_lens[c]
        name, rest = suffix[c, :nl], suffix[c, nl:]
        if position == "middle":
            rows.append(torch.cat([prefix[c], ctx[c, :half], name, ctx[c, half:], rest], dim=0))
        elif position == "front":
            rows.append(torch.cat([prefix[c], name, ctx[c], rest], dim=0))
        else:
            raise ValueError(position)
    return torch.stack(rows)


def text_tower(prompts: Tensor, eot_index: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """trainers/mvlpt.py:105-130: + positional_embedding[:L_t]; causal blocks; ln_final; take the row at the EOT
    token (argmax of the token ids, computed by the caller); @ text_projection."""
    heads = sd["ln_final.weight"].shape[0] // 64
    L = prompts.shape[1]
    x = prompts + sd["positional_embedding"].to(prompts.dtype)[:L]
    for l in range(_n_layers(sd, "transformer.")):
        x = residual_block(x, sd, f"transformer.resblocks.{l}.", heads, causal=True)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"])
    x = x[torch.arange(x.shape[0]), eot_index]
    return x @ sd["text_projection"].to(x.dtype)


# ----------------------------------------------------------------------------------------------- UPT projection
def upt_project(pp: Dict[str, Tensor], n_ctx: int, v: int, dtype=torch.float32):
    """trainers/mvlpt.py:376-414, PROJECT_METHOD='transformer'.  pp holds ctx, vpt_embeddings[, vpt_embeddings_deep]
    and the projection modules.  The 1-layer/1-head block sees a (1, n_tok, p) tensor in (L, N, E) layout, i.e.
    sequence length 1: softmax == 1 and attention reduces to out_proj(v_proj(ln_1(x))) per token."""
    ctx = pp["ctx"]
    vpt = pp["vpt_embeddings"]
    if pp.get("vpt_embeddings_deep") is not None:
        vpt = torch.cat([vpt, pp["vpt_embeddings_deep"]], dim=0)
    dv, dt = vpt.shape[-1], ctx.shape[-1]
    lin = lambda x, nm: x @ pp[nm + ".weight"].to(x.dtype).t() + pp[nm + ".bias"].to(x.dtype)
    c = lin(ctx.reshape(-1, dt).to(dtype), "mvlpt_proj_ctx_coop_pre")
    p = lin(vpt.reshape(-1, dv).to(dtype), "mvlpt_proj_ctx_vpt_pre")
    n_c = c.shape[0]
    x = torch.cat([c, p], dim=0).float()
    pre = "mvlpt_proj.resblocks.0."
    pd = x.shape[-1]
    h = layer_norm(x, pp[pre + "ln_1.weight"], pp[pre + "ln_1.bias"])
    w_in, b_in = pp[pre + "attn.in_proj_weight"].float(), pp[pre + "attn.in_proj_bias"].float()
    val = h @ w_in[2 * pd:].t() + b_in[2 * pd:]
    x = x + val @ pp[pre + "attn.out_proj.weight"].float().t() + pp[pre + "attn.out_proj.bias"].float()
    h = layer_norm(x, pp[pre + "ln_2.weight"], pp[pre + "ln_2.bias"])
    t = h @ pp[pre + "mlp.c_fc.weight"].float().t() + pp[pre + "mlp.c_fc.bias"].float()
    x = x + quick_gelu(t) @ pp[pre + "mlp.c_proj.weight"].float().t() + pp[pre + "mlp.c_proj.bias"].float()
    x = x.to(dtype)
    c2 = lin(x[:n_c], "mvlpt_proj_ctx_coop_post").reshape(-1, n_ctx, dt)
    c2 = c2.squeeze(0)
    p2 = lin(x[n_c:], "mvlpt_proj_ctx_vpt_post").reshape(-1, v, dv)
    return c2, p2[:1], (p2[1:] if p2.shape[0] > 1 else None)


# ----------------------------------------------------------------------------------------------- head + loss
def logit_head(img_feat: Tensor, txt_feat: Tensor, logit_scale: Tensor, task: Optional[Tensor] = None,
               task_ranges: Optional[Tensor] = None) -> Tensor:
    """trainers/mvlpt.py:550-554 (+ :573-581): L2-normalise both, logits = exp(logit_scale) * i @ t^T; optional
    per-task selection multiplies foreign-task logits by 0.  task [B] int; task_ranges [T,2] = (start,end)."""
    i = img_feat / img_feat.norm(dim=-1, keepdim=True)
    t = txt_feat / txt_feat.norm(dim=-1, keepdim=True)
    logits = logit_scale.to(i.dtype).exp() * i @ t.t()
    if task is not None and task_ranges is not None:
        idx = torch.arange(logits.shape[1]).unsqueeze(0)
        lo = task_ranges[task, 0].unsqueeze(-1)
        hi = task_ranges[task, 1].unsqueeze(-1)
        logits = logits * ((idx >= lo).float() * (idx < hi).float()).to(logits.dtype)
    return logits


def cross_entropy(logits: Tensor, label: Tensor) -> Tensor:
    """trainers/mvlpt.py:914-916,922/931: integer labels, or multi-hot rows normalised to sum 1 used as soft
    targets; mean over the batch."""
    lsm = torch.log_softmax(logits.float(), dim=-1)
    if label.dim() == 1:
        return -lsm[torch.arange(logits.shape[0]), label].mean()
    y = label.float()
    y = y / y.sum(dim=-1, keepdim=True)
    return -(y * lsm).sum(dim=-1).mean()


def sgd_step(params: List[Tensor], grads: List[Tensor], bufs: List[Optional[Tensor]], lr: float,
             momentum: float = 0.9, weight_decay: float = 5e-4) -> List[Tensor]:
    """torch.optim.SGD as Dassl builds it (dampening 0, no Nesterov): g += wd*p;
    buf = mu*buf + g (buf = g on the first step); p -= lr*buf.  Updates in place, returns the new buffers."""
    out = []
    for p, g, b in zip(params, grads, bufs):
        g = g + weight_decay * p
        b = g.clone() if b is None else momentum * b + g
        p.sub_(lr * b)
        out.append(b)
    return out


# ----------------------------------------------------------------------------------------------- whole model
def cocoop_logits(img_f: Tensor, sd: Dict[str, Tensor], pp: Dict[str, Tensor], embedding: Tensor, eot_index: Tensor,
                  n_ctx: int) -> Tensor:
    """CoCoOp branch, trainers/mvlpt.py:348-374 (forward_cocoop) + :556-571: the meta network maps each normalised image
    feature to a bias that shifts every context vector; every image gets its own set of class prompts, hence its own
    text features; logits[b, c] = exp(logit_scale) * <img_b, txt_{b,c}>."""
    img_n = img_f / img_f.norm(dim=-1, keepdim=True)
    h1 = torch.relu(img_n @ pp["meta_net.linear1.weight"].t().to(img_n.dtype) + pp["meta_net.linear1.bias"].to(img_n.dtype))
    bias = h1 @ pp["meta_net.linear2.weight"].t().to(img_n.dtype) + pp["meta_net.linear2.bias"].to(img_n.dtype)
    ctx_shifted = pp["cocoop_ctx"].to(img_n.dtype).unsqueeze(0) + bias.unsqueeze(1)  # [B, n, d_t]
    scale = sd["logit_scale"].exp()
    rows = []
    for b in range(img_f.shape[0]):
        prompts = coop_prompts(embedding, ctx_shifted[b], (), n_ctx, "end")  # construct_prompts(ctx_i, prefix, suffix)
        txt = text_tower(prompts, eot_index, sd)
        txt = txt / txt.norm(dim=-1, keepdim=True)
        rows.append(scale * img_n[b] @ txt.t())
    return torch.stack(rows)


def custom_clip_forward(image: Tensor, sd: Dict[str, Tensor], pp: Dict[str, Tensor], embedding: Tensor,
                        eot_index: Tensor, name_lens: Sequence[int], n_ctx: int, v: int, position: str = "end",
                        upt: bool = False, task: Optional[Tensor] = None,
                        task_ranges: Optional[Tensor] = None, cocoop_n_ctx: int = 0, drop_p: float = 0.0,
                        drop_keep: Optional[Sequence[Tensor]] = None) -> Tensor:
    """trainers/mvlpt.py:540-583: projection -> image tower -> prompt assembly -> text tower -> cosine logits
    (or, with COCOOP.N_CTX > 0, the instance-conditioned branch)."""
    ctx, vpt, vpt_deep = pp.get("ctx"), pp.get("vpt_embeddings"), pp.get("vpt_embeddings_deep")
    if upt and ctx is not None and vpt is not None:
        ctx, vpt, vpt_deep = upt_project(pp, n_ctx, v, dtype=image.dtype)
    proj = (pp["vpt_proj.weight"], pp["vpt_proj.bias"]) if "vpt_proj.weight" in pp else None
    img_f = image_tower(image, sd, vpt, vpt_deep, proj, drop_p, drop_keep)
    if cocoop_n_ctx:
        logits = cocoop_logits(img_f, sd, pp, embedding.to(image.dtype), eot_index, cocoop_n_ctx)
        if task is not None and task_ranges is not None:  # :573-581
            idx = torch.arange(logits.shape[1])
            lo, hi = task_ranges[task, 0].unsqueeze(-1), task_ranges[task, 1].unsqueeze(-1)
            logits = logits * ((idx >= lo).float() * (idx < hi).float())
        return logits
    prompts = coop_prompts(embedding.to(image.dtype), ctx, name_lens, n_ctx, position)
    txt_f = text_tower(prompts, eot_index, sd)
    return logit_head(img_f, txt_f, sd["logit_scale"], task, task_ranges)


def train_step(image: Tensor, label: Tensor, sd: Dict[str, Tensor], pp: Dict[str, Tensor], **kw):
    """One reference train step up to the gradients (trainers/mvlpt.py:910-951): returns (logits, loss, grads)
    with grads keyed like pp, for the tensors in pp that are floating point (all trainable prompt tensors)."""
    leaves = {k: t.detach().clone().requires_grad_(True) for k, t in pp.items()}
    logits = custom_clip_forward(image, sd, leaves, **kw)
    loss = cross_entropy(logits, label)
    keys = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    return logits.detach(), loss.detach(), {k: g for k, g in zip(keys, gs) if g is not None}


def flops_step(arch: dict, B: int, C: int, L_t: int, v: int, n_ctx: int) -> float:
    """Algorithmic FLOPs of one train step (SURVEY.md §8d / BASELINE.md §3)."""
    d, ly, p = arch["vision_width"], arch["vision_layers"], arch["vision_patch_size"]
    e, dt, lt = arch["embed_dim"], arch["transformer_width"], arch["transformer_layers"]
    n_p = (arch["image_resolution"] // p) ** 2
    L = 1 + v + n_p
    fwd = lambda L_, d_, ly_: ly_ * (24 * L_ * d_ * d_ + 4 * L_ * L_ * d_)
    bwd = lambda L_, d_, ly_: ly_ * (24 * L_ * d_ * d_ + 8 * L_ * L_ * d_)
    f = B * (fwd(L, d, ly) + 2 * n_p * d * 3 * p * p + 2 * d * e) + C * (fwd(L_t, dt, lt) + 2 * dt * e) + 2 * B * C * e
    if v > 0:
        f += B * bwd(L, d, ly) + 2 * B * C * e
    if n_ctx > 0:
        f += C * bwd(L_t, dt, lt) + 2 * B * C * e
    return float(f)
