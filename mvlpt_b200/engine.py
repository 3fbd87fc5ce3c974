"""Explicit forward/backward engine of the MVLPT hot path over the C-ABI kernels.

Everything the reference does through autograd over torch ops (clip/model.py:185-188, trainers/mvlpt.py:52-130,
540-583, 910-951) is sequenced here by hand: forward saves exactly what the dgrad-only backward needs
 into preallocated HBM buffers, backward walks the blocks in reverse and reduces activation
gradients into prompt gradients.  Frozen CLIP weights never receive a gradient write.

HBM layout (N sequences of L tokens, M = N*L, width d):
  residual stream      fp32 [M, d]    one buffer per block boundary when training (they double as the saved
                                      LayerNorm inputs), two ping-pong... none when evaluating (updated in place)
  LN output h          fp16 [M, d]    workspace, consumed at once by the next GEMM
  qkv / attention out  fp16 [M, 3d] / [M, d], lse fp32 [N, heads, L]     saved per block when training
  FC1 pre-activation t fp16 [M, 4d]   saved per block when training; QuickGELU output g is a workspace
  gradient stream      fp16 [M, d] (GEMM operand and running sum; optional fp32 running sum), dt [M,4d], dh [M,d],
                       dqkv [M,3d] fp16
Gradients are carried multiplied by `grad_scale` (power of two) so the fp16 operands of the dgrad GEMMs stay in
range; the prompt-gradient reductions divide it out in fp32.

torch is used for device memory, streams and (optionally) CUDA-graph capture only.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import ops

F16, F32, I32 = torch.float16, torch.float32, torch.int32

# Precision of the running residual GRADIENT between blocks.  fp16 (default) is the reference's own (its activations and
# their gradients are fp16 tensors) and costs 10 instead of 16 bytes per element in every LayerNorm backward;
# MVLPT_GRAD_STREAM_F32=1 keeps an fp32 running sum next to the fp16 GEMM operand.  Either way every sum inside a kernel
# is fp32 and the values are scaled by `grad_scale`.
GRAD_STREAM_F32 = os.environ.get("MVLPT_GRAD_STREAM_F32", "0") == "1"


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class BlockWeights:
    """One frozen ResidualAttentionBlock (clip/model.py:167-188) in kernel layout."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, device):
        h = lambda k: sd[prefix + k].detach().to(device=device, dtype=F16).contiguous()
        f = lambda k: sd[prefix + k].detach().to(device=device, dtype=F32).contiguous()
        self.ln1_g, self.ln1_b = f("ln_1.weight"), f("ln_1.bias")
        self.ln2_g, self.ln2_b = f("ln_2.weight"), f("ln_2.bias")
        self.w_qkv, self.b_qkv = h("attn.in_proj_weight"), h("attn.in_proj_bias")
        self.w_o, self.b_o = h("attn.out_proj.weight"), h("attn.out_proj.bias")
        self.w_fc, self.b_fc = h("mlp.c_fc.weight"), h("mlp.c_fc.bias")
        self.w_pr, self.b_pr = h("mlp.c_proj.weight"), h("mlp.c_proj.bias")
        self.w_qkv_t = self.w_o_t = self.w_fc_t = self.w_pr_t = None
        # LayerNorm carried through the linears (csrc/gemm_sm100.cuh): LN(x).W^T + b = rstd*(xt.W^T - (mu-c)*sg) + bp with
        # sg[n] = sum_k gamma[k] W[n,k], bp[n] = b[n] + sum_k beta[k] W[n,k] — summed in fp32, stored like the biases (fp16), once
        # per layer at load time
        self.sg_qkv = (self.w_qkv.float() @ self.ln1_g).to(F16).contiguous()
        self.bp_qkv = (self.b_qkv.float() + self.w_qkv.float() @ self.ln1_b).to(F16).contiguous()
        self.sg_fc = (self.w_fc.float() @ self.ln2_g).to(F16).contiguous()
        self.bp_fc = (self.b_fc.float() + self.w_fc.float() @ self.ln2_b).to(F16).contiguous()

    def ensure_transposed(self):
        """[in, out] copies: the W operand of dA = dY . W (dgrad); only towers that train need them."""
        if self.w_qkv_t is None:
            self.w_qkv_t = self.w_qkv.t().contiguous()
            self.w_o_t = self.w_o.t().contiguous()
            self.w_fc_t = self.w_fc.t().contiguous()
            self.w_pr_t = self.w_pr.t().contiguous()


class TowerBuffers:
    """Activation / gradient buffers of one tower for a fixed (N, L)."""

    def __init__(self, N: int, L: int, d: int, heads: int, layers: int, train: bool, device):
        self.N, self.L, self.d, self.heads, self.layers, self.train = N, L, d, heads, layers, train
        M = N * L
        self.M = M
        e = lambda *s, dt=F16: torch.empty(*s, device=device, dtype=dt)
        nsave = layers if train else 1
        self.x = [e(M, d, dt=F32) for _ in range(layers + 1 if train else 1)]
        self.xmid = [e(M, d, dt=F32) for _ in range(nsave)] if train else self.x
        self.h = e(M, d)  # LayerNorm output, or xt = (x - c) * gamma when the linears carry the LayerNorm
        self.rec = [e(M, ops.LN_REC, dt=F32) for _ in range(2)]  # row records of the block input / of x after attention
        self.qkv = [e(M, 3 * d) for _ in range(nsave)]
        self.o = [e(M, d) for _ in range(nsave)]
        self.lse = [e(N, heads, L, dt=F32) for _ in range(nsave)]
        self.t = [e(M, 4 * d) for _ in range(nsave)] if train else None
        self.g = e(M, 4 * d)
        if train:
            self.dx = e(M, d, dt=F32) if GRAD_STREAM_F32 else None  # fp32 running gradient (optional, see above)
            self.dx16 = e(M, d)
            self.dt = e(M, 4 * d)
            self.dh = e(M, d)
            self.dqkv = e(M, 3 * d)

    def idx(self, l: int) -> int:
        return l if self.train else 0

    def x_in(self, l: int) -> torch.Tensor:
        return self.x[l] if self.train else self.x[0]

    def x_out(self, l: int) -> torch.Tensor:
        return self.x[l + 1] if self.train else self.x[0]

    def x_mid(self, l: int) -> torch.Tensor:
        return self.xmid[l] if self.train else self.x[0]


def block_forward(w: BlockWeights, a: TowerBuffers, l: int, causal: bool, h_ready: bool = False, next_ln=None) -> bool:
    """x += attn(ln_1(x)); x += mlp(ln_2(x))   — clip/model.py:185-188.

    When the shape allows it (ops.gemm_ln_supported) no LayerNorm kernel runs: the residual-stream linears (out-proj, FC2)
    emit xt = (x - c) * gamma and the row records, the QKV / FC1 linears finish the LayerNorm in their epilogue — five
    launches per block: QKV, attention, out-proj, FC1 (+ QuickGELU), FC2.  `next_ln` = gamma of the LayerNorm that will
    read this block's output (the next block's ln_1); `h_ready`: a.h / a.rec[0] already hold xt / records of the input
    (written by the previous block's FC2).  Returns whether they hold those of the output on exit."""
    M, d, i = a.M, a.d, a.idx(l)
    xin, xmid, xout = a.x_in(l), a.x_mid(l), a.x_out(l)
    aux = a.t[i] if a.train else None
    if not ops.gemm_ln_supported(M, d):
        ops.ln_fwd(xin, w.ln1_g, w.ln1_b, a.h, M, d)
        ops.gemm(a.h, w.w_qkv, a.qkv[i], bias=w.b_qkv)
        ops.fmha_fwd(a.qkv[i], a.o[i], a.lse[i], a.N, a.L, d, a.heads, causal)
        ops.gemm(a.o[i], w.w_o, xmid, bias=w.b_o, resid=xin)
        ops.ln_fwd(xmid, w.ln2_g, w.ln2_b, a.h, M, d)
        ops.gemm(a.h, w.w_fc, a.g, bias=w.b_fc, act=ops.ACT_QUICKGELU, aux_out=aux)
        ops.gemm(a.g, w.w_pr, xout, bias=w.b_pr, resid=xmid)
        return False
    if not h_ready:
        ops.ln_prep(xin, w.ln1_g, a.h, a.rec[0], M, d)
    ops.gemm(a.h, w.w_qkv, a.qkv[i], ln_cons=(a.rec[0], w.sg_qkv, w.bp_qkv))
    ops.fmha_fwd(a.qkv[i], a.o[i], a.lse[i], a.N, a.L, d, a.heads, causal)
    ops.gemm(a.o[i], w.w_o, xmid, bias=w.b_o, resid=xin, ln_prod=(a.rec[0], a.rec[1], w.ln2_g, a.h))
    ops.gemm(a.h, w.w_fc, a.g, act=ops.ACT_QUICKGELU, aux_out=aux, ln_cons=(a.rec[1], w.sg_fc, w.bp_fc))
    if next_ln is not None:
        ops.gemm(a.g, w.w_pr, xout, bias=w.b_pr, resid=xmid, ln_prod=(a.rec[1], a.rec[0], next_ln, a.h))
        return True
    ops.gemm(a.g, w.w_pr, xout, bias=w.b_pr, resid=xmid)
    return False


def block_backward(w: BlockWeights, a: TowerBuffers, l: int, causal: bool):
    """dgrad-only backward of one block; a.dx / a.dx16 hold d(loss)/d(x_out) on entry and
    d(loss)/d(x_in) on exit."""
    M, d = a.M, a.d
    ops.gemm(a.dx16, w.w_pr_t, a.dt, act=ops.ACT_MUL_DQUICKGELU, aux_in=a.t[l])
    ops.gemm(a.dt, w.w_fc_t, a.dh)
    ops.ln_bwd(a.dh, a.xmid[l], w.ln2_g, a.dx, a.dx16, M, d, accumulate=True)
    ops.gemm(a.dx16, w.w_o_t, a.dh)
    ops.fmha_bwd(a.qkv[l], a.o[l], a.dh, a.lse[l], a.dqkv, a.N, a.L, d, a.heads, causal)
    ops.gemm(a.dqkv, w.w_qkv_t, a.dh)
    ops.ln_bwd(a.dh, a.x[l], w.ln1_g, a.dx, a.dx16, M, d, accumulate=True)


class ImageTower:
    """VisionTransformer with visual-prompt injection — trainers/mvlpt.py:45-93 over clip/model.py:202-236."""

    def __init__(self, sd: Dict[str, torch.Tensor], device):
        cw = sd["visual.conv1.weight"]
        self.d, _, self.p, _ = cw.shape
        self.heads = self.d // 64
        self.layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        self.G = sd["visual.positional_embedding"].shape[0] - 1
        self.res = int(round(math.sqrt(self.G))) * self.p
        self.e = sd["visual.proj"].shape[1]
        K = 3 * self.p * self.p
        self.Kp = _round_up(K, 8)
        conv = torch.zeros(self.d, self.Kp, dtype=F16, device=device)
        conv[:, :K] = cw.detach().reshape(self.d, K).to(device=device, dtype=F16)
        self.conv_w = conv
        f = lambda k: sd[k].detach().to(device=device, dtype=F32).contiguous()
        self.cls, self.pos = f("visual.class_embedding"), f("visual.positional_embedding")
        self.ln_pre_g, self.ln_pre_b = f("visual.ln_pre.weight"), f("visual.ln_pre.bias")
        self.ln_post_g, self.ln_post_b = f("visual.ln_post.weight"), f("visual.ln_post.bias")
        self.proj = sd["visual.proj"].detach().to(device=device, dtype=F16).contiguous()  # [d, e]
        self.proj_t = self.proj.t().contiguous()  # [e, d]
        self.proj_t2 = torch.cat([self.proj_t, self.proj_t], dim=1).contiguous()  # [e, 2d]: against pooled [hi | lo]
        self.blocks = [BlockWeights(sd, f"visual.transformer.resblocks.{i}.", device) for i in range(self.layers)]
        self.device = device
        self._bufs: Dict[tuple, dict] = {}

    def buffers(self, B: int, v: int, train: bool) -> dict:
        key = (B, v, train)
        if key not in self._bufs:
            dev, d = self.device, self.d
            L = 1 + v + self.G
            if train:
                for b in self.blocks:
                    b.ensure_transposed()
            self._bufs[key] = dict(
                act=TowerBuffers(B, L, d, self.heads, self.layers, train, dev),
                patches=torch.empty(B * self.G, self.Kp, device=dev, dtype=F16),
                pe=torch.empty(B * self.G, d, device=dev, dtype=F16),
                cls_idx=(torch.arange(B, device=dev, dtype=I32) * L).contiguous(),
                pooled=torch.empty(B, 2 * d, device=dev, dtype=F16),
                feat=torch.empty(B, self.e, device=dev, dtype=F32),
                dpool=torch.empty(B, d, device=dev, dtype=F16) if train else None,
            )
        return self._bufs[key]

    def executed_layers(self, n_deep: Optional[int]) -> List[int]:
        """Blocks that actually run: the reference skips block l when l > vpt_embeddings_deep.shape[0]
        (trainers/mvlpt.py:71-83, no else branch)."""
        if n_deep is None:
            return list(range(self.layers))
        return [l for l in range(self.layers) if l == 0 or l <= n_deep]

    def forward(self, image: torch.Tensor, vpt: Optional[torch.Tensor], vpt_deep: Optional[torch.Tensor],
                train: bool, drop_p: float = 0.0, drop_seed: int = 0) -> torch.Tensor:
        """image [B,3,H,W] fp16/fp32; vpt [1,v,d] or None; vpt_deep [n_deep,v,d] or None -> features fp32 [B,e].
        drop_p > 0: vpt_dropout in training mode (trainers/mvlpt.py:76,425) with the counter-based mask of `drop_seed`;
        the backward of the same buffers replays it."""
        B = image.shape[0]
        v = 0 if vpt is None else vpt.shape[1]
        bf = self.buffers(B, v, train)
        a: TowerBuffers = bf["act"]
        ops.im2col(image, bf["patches"], B, image.shape[2], image.shape[3], self.p, self.Kp)
        ops.gemm(bf["patches"], self.conv_w, bf["pe"])
        ops.embed_assemble(bf["pe"], self.cls, self.pos, self.ln_pre_g, self.ln_pre_b, vpt, a.x_in(0), B, self.G, v,
                           self.d)
        if v == 0:
            drop_p = 0.0
        if drop_p > 0.0:
            ops.set_prompt_rows(a.x_in(0), vpt, B, a.L, v, self.d, drop_p, drop_seed, 0)
        bf["drop"] = (drop_p, drop_seed)
        run = self.executed_layers(None if vpt_deep is None else vpt_deep.shape[0])
        last = a.x_in(0)
        h_ready = False
        for k, l in enumerate(run):
            blk = self.blocks[l]
            if vpt_deep is not None and l >= 1:
                # rows 1..v of the previous output are discarded and replaced (trainers/mvlpt.py:74-82)
                src = a.x_in(l)
                if a.train and last is not src:
                    src.copy_(last)  # only reachable through the skipped-layer quirk
                ops.set_prompt_rows(src, vpt_deep[l - 1], B, a.L, v, self.d, drop_p, drop_seed, l,
                                    ln=(a.h, a.rec[0], blk.ln1_g) if h_ready else None)
            nxt = self.blocks[run[k + 1]] if k + 1 < len(run) else None
            h_ready = block_forward(blk, a, l, causal=False, h_ready=h_ready, next_ln=None if nxt is None else nxt.ln1_g)
            last = a.x_out(l)
        bf["final"] = last
        bf["run"] = run
        # the pooled row goes into the projection as an fp16 pair hi + lo (K = 2d): the feature is the last thing the
        # tower computes, nothing downstream averages the rounding of its input away
        ops.ln_fwd(last, self.ln_post_g, self.ln_post_b, bf["pooled"], B, self.d, row_index=bf["cls_idx"], hilo=True)
        ops.gemm(bf["pooled"], self.proj_t2, bf["feat"])
        return bf["feat"]

    def backward(self, dfeat16: torch.Tensor, B: int, v: int, n_deep: Optional[int], grad_vpt: torch.Tensor,
                 grad_deep: Optional[torch.Tensor], inv_scale: float):
        """dfeat16 fp16 [B,e] (scaled) -> grad_vpt fp32 [v,d], grad_deep fp32 [n_deep,v,d] (unscaled)."""
        bf = self.buffers(B, v, True)
        a: TowerBuffers = bf["act"]
        drop_p, drop_seed = bf.get("drop", (0.0, 0))
        ops.gemm(dfeat16, self.proj, bf["dpool"])
        if a.dx is not None:
            ops.zero(a.dx)
        ops.zero(a.dx16)
        ops.ln_bwd(bf["dpool"], bf["final"], self.ln_post_g, a.dx, a.dx16, B, self.d, accumulate=False,
                   row_index=bf["cls_idx"])
        for l in reversed(bf["run"]):
            block_backward(self.blocks[l], a, l, causal=False)
            if n_deep is not None and l >= 1:
                ops.prompt_grad(a.dx, a.dx16, grad_deep[l - 1], B, a.L, v, self.d, inv_scale, True, drop_p, drop_seed, l)
        ops.prompt_grad(a.dx, None if a.dx is not None else a.dx16, grad_vpt, B, a.L, v, self.d, inv_scale, False,
                        drop_p, drop_seed, 0)


class TextTower:
    """Text Transformer over assembled prompts — trainers/mvlpt.py:95-130 with forward_coop (:439-515)."""

    def __init__(self, sd: Dict[str, torch.Tensor], device):
        self.d = sd["ln_final.weight"].shape[0]
        self.heads = self.d // 64
        self.layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
        self.e = sd["text_projection"].shape[1]
        f = lambda k: sd[k].detach().to(device=device, dtype=F32).contiguous()
        self.pos = f("positional_embedding")
        self.ln_g, self.ln_b = f("ln_final.weight"), f("ln_final.bias")
        self.proj = sd["text_projection"].detach().to(device=device, dtype=F16).contiguous()  # [d_t, e]
        self.proj_t = self.proj.t().contiguous()
        self.proj_t2 = torch.cat([self.proj_t, self.proj_t], dim=1).contiguous()  # [e, 2d]: against pooled [hi | lo]
        self.blocks = [BlockWeights(sd, f"transformer.resblocks.{i}.", device) for i in range(self.layers)]
        self.device = device
        self._bufs: Dict[tuple, dict] = {}

    def buffers(self, C: int, Lt: int, train: bool) -> dict:
        key = (C, Lt, train)
        if key not in self._bufs:
            dev, d = self.device, self.d
            if train:
                for b in self.blocks:
                    b.ensure_transposed()
            self._bufs[key] = dict(
                act=TowerBuffers(C, Lt, d, self.heads, self.layers, train, dev),
                pooled=torch.empty(C, 2 * d, device=dev, dtype=F16),
                feat=torch.empty(C, self.e, device=dev, dtype=F32),
                dpool=torch.empty(C, d, device=dev, dtype=F16) if train else None,
            )
        return self._bufs[key]

    def forward(self, emb: torch.Tensor, ctx: Optional[torch.Tensor], slot: Optional[torch.Tensor],
                eot_rows: torch.Tensor, n_ctx: int, csc: bool, train: bool) -> torch.Tensor:
        """emb fp32 [C,Lt,d]; ctx [n,d] | [C,n,d] | None; slot int32 [C,Lt]; eot_rows int32 [C] = c*Lt + eot(c)."""
        C, Lt, d = emb.shape
        return self.forward_assembled(C, Lt, eot_rows, train,
                                      lambda x0: ops.text_assemble(emb, ctx, slot, self.pos, x0, C, Lt, n_ctx, d, csc))

    def forward_assembled(self, N: int, Lt: int, eot_rows: torch.Tensor, train: bool, assemble) -> torch.Tensor:
        """N sequences of Lt rows; `assemble(x0)` writes the tower input (prompt rows + positional embedding) into the
        fp32 buffer x0 [N*Lt, d] (CoOp: one sequence per class; CoCoOp: one per (image, class))."""
        d = self.d
        bf = self.buffers(N, Lt, train)
        a: TowerBuffers = bf["act"]
        assemble(a.x_in(0))
        h_ready = False
        for l in range(self.layers):
            nxt = self.blocks[l + 1] if l + 1 < self.layers else None
            h_ready = block_forward(self.blocks[l], a, l, causal=True, h_ready=h_ready,
                                    next_ln=None if nxt is None else nxt.ln1_g)
        bf["final"] = a.x_out(self.layers - 1)
        ops.ln_fwd(bf["final"], self.ln_g, self.ln_b, bf["pooled"], N, d, row_index=eot_rows, hilo=True)
        ops.gemm(bf["pooled"], self.proj_t2, bf["feat"])
        return bf["feat"]

    def backward_to_input(self, dfeat16: torch.Tensor, N: int, Lt: int, eot_rows: torch.Tensor) -> torch.Tensor:
        """dfeat16 fp16 [N,e] (scaled) -> the gradient w.r.t. the tower input, fp16 (or fp32) [N*Lt, d], scaled."""
        bf = self.buffers(N, Lt, True)
        a: TowerBuffers = bf["act"]
        ops.gemm(dfeat16, self.proj, bf["dpool"])
        if a.dx is not None:
            ops.zero(a.dx)
        ops.zero(a.dx16)
        ops.ln_bwd(bf["dpool"], bf["final"], self.ln_g, a.dx, a.dx16, N, self.d, accumulate=False, row_index=eot_rows)
        for l in reversed(range(self.layers)):
            block_backward(self.blocks[l], a, l, causal=True)
        return a.dx if a.dx is not None else a.dx16

    def backward(self, dfeat16: torch.Tensor, C: int, Lt: int, eot_rows: torch.Tensor, ctx_pos: torch.Tensor,
                 n_ctx: int, csc: bool, grad_ctx: torch.Tensor, inv_scale: float):
        dx0 = self.backward_to_input(dfeat16, C, Lt, eot_rows)
        ops.ctx_grad(dx0, ctx_pos, grad_ctx, C, Lt, n_ctx, self.d, csc, inv_scale)


class LogitHead:
    """Cosine-similarity logits + cross-entropy — trainers/mvlpt.py:550-554, 573-581, 914-931."""

    def __init__(self, logit_scale: float, e: int, device):
        self.s = float(math.exp(logit_scale))
        self.e = e
        self.device = device
        self._bufs: Dict[tuple, dict] = {}
        self._tbufs: Dict[int, dict] = {}

    def text_buffers(self, C: int) -> dict:
        """Normalised text features of a C-class label space.  They depend on the classes only, never on the batch: every
        batch size reads the same tensors (a ragged last batch of an evaluation loader must not see stale ones)."""
        if C not in self._tbufs:
            dev, e = self.device, self.e
            z = lambda *s, dt=F16: torch.zeros(*s, device=dev, dtype=dt)
            self._tbufs[C] = dict(t16=z(C, e), t32=z(C, e, dt=F32), t_inv=z(C, dt=F32), t16_t=z(e, _round_up(C, 8)),
                                  t16x3=z(C, 3 * e), dt32=z(C, e, dt=F32), dtfeat16=z(C, e))
        return self._tbufs[C]

    def buffers(self, B: int, C: int) -> dict:
        key = (B, C)
        if key not in self._bufs:
            dev, e = self.device, self.e
            ldc, ldb = _round_up(C, 8), _round_up(B, 8)
            z = lambda *s, dt=F16: torch.zeros(*s, device=dev, dtype=dt)
            self._bufs[key] = dict(
                ldc=ldc, ldb=ldb,
                i16=z(B, e), i32=z(B, e, dt=F32), i_inv=z(B, dt=F32), i16x3=z(B, 3 * e),
                logits=z(B, ldc, dt=F32), dz16=z(B, ldc), loss_rows=z(B, dt=F32), pred=z(B, dt=I32),
                hit=z(B, dt=I32), metrics=z(2, dt=F32),
                dz16_t=z(C, ldb), i16_t=z(e, ldb),
                di32=z(B, e, dt=F32), difeat16=z(B, e),
                **self.text_buffers(C),
            )
        return self._bufs[key]

    def normalize_text(self, txt_feat: torch.Tensor):
        C = txt_feat.shape[0]
        bf = self.text_buffers(C)
        ops.l2norm_fwd(txt_feat, bf["t16"], bf["t32"], bf["t_inv"], C, self.e, y16x3=bf["t16x3"], pattern=1)

    def logits(self, img_feat: torch.Tensor, C: int) -> torch.Tensor:
        B = img_feat.shape[0]
        bf = self.buffers(B, C)
        # logits = s * i.t with both normalised features split into fp16 pairs: [hi|hi|lo] . [hi|lo|hi] over K = 3e is
        # hi.hi + hi.lo + lo.hi, the fp32 product to 2^-22 — the cosine itself is not rounded to fp16 operands
        ops.l2norm_fwd(img_feat, bf["i16"], bf["i32"], bf["i_inv"], B, self.e, y16x3=bf["i16x3"], pattern=0)
        ops.gemm(bf["i16x3"], bf["t16x3"], bf["logits"], alpha=self.s, N=C)
        return bf["logits"]

    def backward(self, B: int, C: int, need_img: bool, need_txt: bool):
        """From dz16 (already scaled) to feature gradients difeat16 [B,e] / dtfeat16 [C,e] (fp16, scaled)."""
        bf = self.buffers(B, C)
        e, ldc, ldb = self.e, bf["ldc"], bf["ldb"]
        if need_img:
            ops.transpose_f16(bf["t16"], bf["t16_t"], C, e, e, ldc)
            ops.gemm(bf["dz16"], bf["t16_t"], bf["di32"], alpha=self.s, K=ldc)
            ops.l2norm_bwd(bf["di32"], bf["i32"], bf["i_inv"], bf["difeat16"], B, e)
        if need_txt:
            ops.transpose_f16(bf["dz16"], bf["dz16_t"], B, C, ldc, ldb)
            ops.transpose_f16(bf["i16"], bf["i16_t"], B, e, e, ldb)
            ops.gemm(bf["dz16_t"], bf["i16_t"], bf["dt32"], alpha=self.s, K=ldb)
            ops.l2norm_bwd(bf["dt32"], bf["t32"], bf["t_inv"], bf["dtfeat16"], C, e)


def build_ctx_maps(name_lens: Sequence[int], n_ctx: int, Lt: int, position: str):
    """slot[c,t] (context index living at token t, or -1) and ctx_pos[c,j] (token index of context j) for the three
    CLASS_TOKEN_POSITION layouts of forward_coop (trainers/mvlpt.py:455-510)."""
    C = len(name_lens)
    slot = torch.full((C, Lt), -1, dtype=torch.int32)
    pos = torch.zeros((C, max(n_ctx, 1)), dtype=torch.int32)
    half = n_ctx // 2
    for c, nl in enumerate(name_lens):
        for j in range(n_ctx):
            if position == "end":
                t = 1 + j
            elif position == "middle":
                t = 1 + j if j < half else 1 + nl + j
            elif position == "front":
                t = 1 + nl + j
            else:
                raise ValueError(position)
            slot[c, t] = j
            pos[c, j] = t
    return slot, pos


def rearrange_embedding(emb: torch.Tensor, name_lens: Sequence[int], n_ctx: int, position: str) -> torch.Tensor:
    """Token embedding rows in FINAL order for 'middle'/'front': the reference moves the class-name rows in front
    of (part of) the context (trainers/mvlpt.py:472-510); context slots hold don't-care rows."""
    if position == "end" or n_ctx == 0:
        return emb
    out = emb.clone()
    half = n_ctx // 2
    for c, nl in enumerate(name_lens):
        name = emb[c, 1 + n_ctx:1 + n_ctx + nl]
        if position == "middle":
            out[c, 1 + half:1 + half + nl] = name
        else:
            out[c, 1:1 + nl] = name
    return out
