"""UPT shared-prompt projection (trainers/mvlpt.py:376-414) — forward and backward incl. weight gradients, over the
mvlpt_upt_fwd / mvlpt_upt_bwd entry points (csrc/upt.cu)."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib

# parameter order of include/mvlpt_sm100.h (enum MVLPT_UPT_*), as prompt_learner state-dict keys
_BLK = "mvlpt_proj.resblocks.0."
PARAM_NAMES = [
    "ctx", "vpt_embeddings", "vpt_embeddings_deep",
    "mvlpt_proj_ctx_coop_pre.weight", "mvlpt_proj_ctx_coop_pre.bias",
    "mvlpt_proj_ctx_coop_post.weight", "mvlpt_proj_ctx_coop_post.bias",
    "mvlpt_proj_ctx_vpt_pre.weight", "mvlpt_proj_ctx_vpt_pre.bias",
    "mvlpt_proj_ctx_vpt_post.weight", "mvlpt_proj_ctx_vpt_post.bias",
    _BLK + "ln_1.weight", _BLK + "ln_1.bias", _BLK + "attn.in_proj_weight", _BLK + "attn.in_proj_bias",
    _BLK + "attn.out_proj.weight", _BLK + "attn.out_proj.bias", _BLK + "ln_2.weight", _BLK + "ln_2.bias",
    _BLK + "mlp.c_fc.weight", _BLK + "mlp.c_fc.bias", _BLK + "mlp.c_proj.weight", _BLK + "mlp.c_proj.bias",
]


class UptDesc(ctypes.Structure):
    _fields_ = [("n_ctx", ctypes.c_int), ("v", ctypes.c_int), ("n_deep", ctypes.c_int), ("dt", ctypes.c_int),
                ("dv", ctypes.c_int), ("pd", ctypes.c_int), ("param_f16", ctypes.c_int)]


class UptProjection:
    def __init__(self, prompt_learner):
        pl = prompt_learner
        if pl.csc:
            raise NotImplementedError("UPT projection over class-specific contexts is not implemented")
        params = dict(pl.named_parameters())
        missing = [n for n in PARAM_NAMES if n not in params and n != "vpt_embeddings_deep"]
        if missing:
            raise NotImplementedError(f"UPT projection needs pre/post Linears on both sides (PROJECT_DIM must differ "
                                      f"from both prompt widths); missing {missing}")
        self.pl = pl
        deep = params.get("vpt_embeddings_deep") if pl.vpt_deep else None
        self.n_deep = 0 if deep is None else deep.shape[0]
        self.v = pl.vpt_n_ctx
        self.n = pl.ctx.shape[0]
        self.dt, self.dv = pl.ctx.shape[-1], pl.vpt_embeddings.shape[-1]
        self.pd = params[_BLK + "ln_1.weight"].shape[0]
        f16 = pl.ctx.dtype == torch.float16
        for n in PARAM_NAMES[:11]:
            if n in params and params[n].dtype != pl.ctx.dtype:
                raise _lib.MvlptError(f"{n}: dtype {params[n].dtype} differs from ctx ({pl.ctx.dtype})")
        self.desc = UptDesc(self.n, self.v, self.n_deep, self.dt, self.dv, self.pd, int(f16))
        dev = pl.ctx.device
        ws = int(_lib.lib().mvlpt_upt_workspace(ctypes.byref(self.desc)))
        self.ws = torch.empty(ws // 4, device=dev, dtype=torch.float32)
        rows = (1 + self.n_deep) * self.v
        self.ctx_out = torch.empty(self.n, self.dt, device=dev, dtype=torch.float32)
        self.vpt_out = torch.empty(rows, self.dv, device=dev, dtype=torch.float32)
        self.d_ctx_out = torch.zeros(self.n, self.dt, device=dev, dtype=torch.float32)
        self.d_vpt_out = torch.zeros(rows, self.dv, device=dev, dtype=torch.float32)

    def _ptrs(self, tensors: Dict[str, torch.Tensor]):
        arr = (ctypes.c_void_p * len(PARAM_NAMES))()
        for i, n in enumerate(PARAM_NAMES):
            t = tensors.get(n)
            if n == "vpt_embeddings_deep" and self.n_deep == 0:
                t = None
            if t is not None and (not t.is_cuda or not t.is_contiguous()):
                raise _lib.MvlptError(f"{n}: expected a contiguous CUDA tensor")
            arr[i] = None if t is None else t.data_ptr()
        return arr

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def forward(self):
        """-> (ctx' [n,dt], vpt' [1,v,dv], vpt_deep' [n_deep,v,dv] | None), fp32 views of the output buffers."""
        params = {n: p.detach() for n, p in self.pl.named_parameters()}
        P = self._ptrs(params)
        _lib.check(_lib.lib().mvlpt_upt_fwd(ctypes.byref(self.desc), P, self.ws.data_ptr(), self.ws.numel() * 4,
                                            self.ctx_out.data_ptr(), self.vpt_out.data_ptr(), self._stream()),
                   "mvlpt_upt_fwd")
        v = self.v
        deep = self.vpt_out[v:].view(self.n_deep, v, self.dv) if self.n_deep else None
        return self.ctx_out, self.vpt_out[:v].view(1, v, self.dv), deep

    def grad_input_views(self):
        """Buffers the towers write d(ctx'), d(vpt'), d(vpt_deep') into (fp32, unscaled)."""
        v = self.v
        deep = self.d_vpt_out[v:].view(self.n_deep, v, self.dv) if self.n_deep else None
        return self.d_ctx_out, self.d_vpt_out[:v], deep

    def backward(self, grad_views: Dict[str, torch.Tensor]):
        """Back-propagates d_ctx_out / d_vpt_out (filled by the towers) through the projection into `grad_views`
        (fp32 tensors keyed like named_parameters(), e.g. views of CustomCLIP.grad_buffer())."""
        params = {n: p.detach() for n, p in self.pl.named_parameters()}
        P = self._ptrs(params)
        G = self._ptrs(grad_views)
        _lib.check(_lib.lib().mvlpt_upt_bwd(ctypes.byref(self.desc), P, self.ws.data_ptr(), self.ws.numel() * 4,
                                            self.d_ctx_out.data_ptr(), self.d_vpt_out.data_ptr(), G, self._stream()),
                   "mvlpt_upt_bwd")


class VptProjection:
    """vpt_proj = Linear(VPT.PROJECT -> vision width) (trainers/mvlpt.py:170-175), applied to the shallow prompts (:425)
    and to every deep slab (:76-77) before they enter the image tower; forward and backward incl. the weight gradient,
    over mvlpt_vpt_proj_fwd / mvlpt_vpt_proj_bwd (csrc/upt.cu).  Outputs are fp32 (the reference rounds them to CLIP's
    dtype; the tower rounds once, when it writes the rows)."""

    def __init__(self, prompt_learner):
        pl = prompt_learner
        self.pl = pl
        W = pl.vpt_proj.weight
        self.d, self.p = W.shape
        self.v = pl.vpt_n_ctx
        n_deep = pl.vpt_embeddings_deep.shape[0] if pl.vpt_embeddings_deep is not None else 0
        rows = (1 + n_deep) * self.v
        dev = W.device
        self.out = torch.empty(rows, self.d, device=dev, dtype=torch.float32)
        self.d_out = torch.zeros(rows, self.d, device=dev, dtype=torch.float32)
        self._emb = None

    def _params(self):
        W, b = self.pl.vpt_proj.weight.detach(), self.pl.vpt_proj.bias.detach()
        if not W.is_cuda:
            raise _lib.MvlptError("vpt_proj: parameters must live on a CUDA device")
        return W.contiguous(), b.contiguous()

    def forward(self, vpt: torch.Tensor, deep: Optional[torch.Tensor]):
        """(vpt [1,v,p], deep [n_deep,v,p] | None) -> fp32 ([1,v,d], [n_deep,v,d] | None), views of one buffer."""
        from . import ops
        W, b = self._params()
        v, d = self.v, self.d
        vpt = vpt.detach().to(W.dtype).contiguous()
        if vpt.shape[-1] != self.p or vpt.numel() != v * self.p:
            raise _lib.MvlptError(f"vpt_proj: expected prompts of shape [1,{v},{self.p}], got {tuple(vpt.shape)}")
        ops.vpt_proj_fwd(vpt, W, b, self.out[:v])
        out_deep = None
        if deep is not None:
            deep = deep.detach().to(W.dtype).contiguous()
            nd = deep.shape[0]
            if (1 + nd) * v > self.out.shape[0]:
                self.out = torch.empty((1 + nd) * v, d, device=W.device, dtype=torch.float32)
                self.d_out = torch.zeros((1 + nd) * v, d, device=W.device, dtype=torch.float32)
                ops.vpt_proj_fwd(vpt, W, b, self.out[:v])
            ops.vpt_proj_fwd(deep, W, b, self.out[v:(1 + nd) * v])
            out_deep = self.out[v:(1 + nd) * v].view(nd, v, d)
        self._emb = (vpt, deep)
        return self.out[:v].view(1, v, d), out_deep

    def grad_input_views(self, n_deep):
        """Buffers the image tower writes d(projected prompts) into (fp32, unscaled)."""
        v = self.v
        deep = self.d_out[v:(1 + n_deep) * v].view(n_deep, v, self.d) if n_deep else None
        return self.d_out[:v], deep

    def backward(self, grad_views: Dict[str, torch.Tensor], n_deep):
        """d_out (filled by the tower) -> gradients of vpt_embeddings[_deep], vpt_proj.weight, vpt_proj.bias."""
        from . import ops
        W, _ = self._params()
        vpt, deep = self._emb
        v = self.v
        dW, db = grad_views["vpt_proj.weight"], grad_views["vpt_proj.bias"]
        ops.vpt_proj_bwd(self.d_out[:v], vpt, W, grad_views["vpt_embeddings"], dW, db, accumulate=False)
        if n_deep and deep is not None:
            ops.vpt_proj_bwd(self.d_out[v:(1 + n_deep) * v], deep, W, grad_views["vpt_embeddings_deep"], dW, db,
                             accumulate=True)
