"""Tensor-level wrappers over the C ABI: pointer extraction + shape/dtype checks, nothing else.

torch is used here only as the owner of device memory and streams; every function launches kernels from
libmvlpt_sm100.so on the current CUDA stream and raises MvlptError on failure (no fallback).
"""
from __future__ import annotations

import ctypes
from ctypes import byref
from typing import Optional

import torch

from . import _lib
from ._lib import GemmDesc, check

ACT_NONE, ACT_QUICKGELU, ACT_MUL_DQUICKGELU = 0, 1, 2
LN_EPS = 1e-5
LN_REC = _lib.LN_REC  # floats per row record of the LayerNorm-carrying GEMM chain


class Profiler:
    """Per-launch CUDA-event timing of the heavy kernels (bench.py's roofline pass).  While `ops.PROFILER` is set,
    gemm / fmha / layer-norm wrappers bracket their launch with events on the current stream and record the
    algorithmic FLOPs and bytes of that launch."""

    def __init__(self):
        self.records = []  # (kernel, start_event, end_event, flops, bytes)

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, kernel: str, start, flops: float, nbytes: float):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((kernel, start, e, flops, nbytes))

    def summary(self):
        """{kernel: dict(launches, ms, flops, bytes)} — call after a device synchronize."""
        out = {}
        for k, a, b, f, n in self.records:
            r = out.setdefault(k, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            r["launches"] += 1
            r["ms"] += a.elapsed_time(b)
            r["flops"] += f
            r["bytes"] += n
        return out


PROFILER: Optional[Profiler] = None
# When a list, every gemm() call appends its shape key: the n-th entry is the n-th GEMM kernel launch of the process
# (tools/ncu_traffic.py matches an ncu capture of `-k regex:gemm_f16` against it).
GEMM_LOG: Optional[list] = None


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
        raise _lib.MvlptError(f"{name}: expected contiguous CUDA {dtype}, got {t.dtype} cuda={t.is_cuda} "
                              f"contig={t.is_contiguous()}")


def gemm_ln_supported(M: int, width: int) -> bool:
    """Can the linears around a LayerNorm of `width`-wide rows carry it (mvlpt_gemm_ln) for M rows?"""
    return bool(_lib.lib().mvlpt_gemm_ln_supported(int(M), int(width)))


def gemm(A, W, out, *, bias=None, act=ACT_NONE, aux_in=None, aux_out=None, resid=None, alpha=1.0,
         M=None, N=None, K=None, lda=None, ldw=None, ld_out=None, ld_aux=None, ln_prod=None, ln_cons=None):
    """out[M,N] = epi(alpha * A[M,K] @ W[N,K]^T); see include/mvlpt_sm100.h:mvlpt_gemm / mvlpt_gemm_ln.
    ln_prod=(rec_in | None, rec_out, gamma, xt): the residual-stream linear also emits xt = (out - c) * gamma and the row
    records of `out`.  ln_cons=(rec, sg, bp): A is such an xt; the LayerNorm is finished in the epilogue (bias in bp)."""
    _chk(A, torch.float16, "gemm A")
    _chk(W, torch.float16, "gemm W")
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = W.shape[0] if N is None else N
    lda = A.stride(0) if lda is None else lda
    ldw = W.stride(0) if ldw is None else ldw
    ld_out = out.stride(0) if ld_out is None else ld_out
    aux = aux_in if aux_in is not None else aux_out
    ld_aux = (aux.stride(0) if aux is not None else 0) if ld_aux is None else ld_aux
    out_f32 = int(out.dtype == torch.float32)
    if not out_f32 and out.dtype != torch.float16:
        raise _lib.MvlptError("gemm out must be fp16 or fp32")
    d = GemmDesc(M, N, K, lda, ldw, ld_out, ld_aux, act, out_f32, float(alpha))
    t0 = PROFILER.begin() if PROFILER is not None else None
    tag = ",lnp=1" if ln_prod is not None else (",lnc=1" if ln_cons is not None else "")
    if aux_out is not None:
        tag += ",aux=1"  # the QuickGELU epilogue also stores the pre-activation
    key = f"gemm[M={M},N={N},K={K},act={act},f32={out_f32},resid={int(resid is not None)}{tag}]"
    if GEMM_LOG is not None:
        GEMM_LOG.append(key)
    if ln_prod is not None or ln_cons is not None:
        c = _lib.LnCarry()
        if ln_prod is not None:
            rec_in, rec_out, gamma, xt = ln_prod
            _chk(xt, torch.float16, "gemm xt")
            c.rec_in, c.rec_out, c.gamma, c.xt = _pv(rec_in), _pv(rec_out), _pv(gamma), _pv(xt)
            c.width = N
        else:
            rec, sg, bp = ln_cons
            if bias is not None:
                raise _lib.MvlptError("gemm: with ln_cons the bias is folded into bp")
            _chk(sg, torch.float16, "gemm sg")
            _chk(bp, torch.float16, "gemm bp")
            c.rec, c.sg, bias = _pv(rec), _pv(sg), bp
            c.width = K
        c.eps = LN_EPS
        check(_lib.lib().mvlpt_gemm_ln(byref(d), _p(A), _p(W), _p(bias), _p(aux_in), _p(aux_out), _p(resid), _p(out),
                                       byref(c), _stream()), "mvlpt_gemm_ln")
    else:
        check(_lib.lib().mvlpt_gemm(byref(d), _p(A), _p(W), _p(bias), _p(aux_in), _p(aux_out), _p(resid), _p(out),
                                    _stream()), "mvlpt_gemm")
    if t0 is not None:
        nbytes = 2.0 * (M * K + N * K) + (4.0 if out_f32 else 2.0) * M * N + (4.0 * M * N if resid is not None else 0.0) \
            + (2.0 * M * N if aux is not None else 0.0) + (2.0 * M * N if ln_prod is not None else 0.0)
        PROFILER.end("gemm_f16_tn", t0, 2.0 * M * N * K, nbytes)
        PROFILER.records.append((key, *PROFILER.records[-1][1:]))
    return out


def _pv(t):
    return None if t is None else t.data_ptr()


def ln_prep(x, gamma, xt, rec, rows, d):
    """xt = (x - mean) * gamma (fp16) and the row records that start a LayerNorm-carrying GEMM chain (mvlpt_ln_prep)."""
    t0 = PROFILER.begin() if PROFILER is not None else None
    check(_lib.lib().mvlpt_ln_prep(_p(x), _p(gamma), _p(xt), _p(rec), rows, d, _stream()), "mvlpt_ln_prep")
    if t0 is not None:
        PROFILER.end("ln_fwd", t0, 0.0, 6.0 * rows * d)


def fmha_fwd(qkv, out, lse, N, L, d, heads, causal):
    t0 = PROFILER.begin() if PROFILER is not None else None
    check(_lib.lib().mvlpt_fmha_fwd(_p(qkv), _p(out), _p(lse), N, L, d, heads, int(causal), _stream()), "mvlpt_fmha_fwd")
    if t0 is not None:  # algorithmic: full L^2 for causal too
        PROFILER.end("fmha_fwd", t0, 4.0 * N * L * L * d, 2.0 * N * L * 4 * d + 4.0 * N * heads * L)


def fmha_bwd(qkv, o, d_o, lse, dqkv, N, L, d, heads, causal):
    t0 = PROFILER.begin() if PROFILER is not None else None
    check(_lib.lib().mvlpt_fmha_bwd(_p(qkv), _p(o), _p(d_o), _p(lse), _p(dqkv), N, L, d, heads, int(causal), _stream()),
          "mvlpt_fmha_bwd")
    if t0 is not None:
        PROFILER.end("fmha_bwd", t0, 8.0 * N * L * L * d, 2.0 * N * L * 8 * d + 4.0 * N * heads * L)


def ln_fwd(x, gamma, beta, y, rows, d, row_index=None, hilo=False):
    """hilo: y is [rows, 2d] = [hi | lo] (mvlpt_ln_fwd_hilo)."""
    t0 = PROFILER.begin() if PROFILER is not None else None
    check(_lib.lib().mvlpt_ln_fwd_hilo(_p(x), _p(row_index), _p(gamma), _p(beta), _p(y), rows, d, LN_EPS, int(hilo),
                                       _stream()), "mvlpt_ln_fwd")
    if t0 is not None:
        PROFILER.end("ln_fwd", t0, 0.0, 6.0 * rows * d)


def ln_bwd(dy, x, gamma, dx_stream, dx16, rows, d, accumulate, row_index=None):
    t0 = PROFILER.begin() if PROFILER is not None else None
    check(_lib.lib().mvlpt_ln_bwd(_p(dy), _p(x), _p(row_index), _p(gamma), _p(dx_stream), _p(dx16), rows, d, LN_EPS,
                                  int(accumulate), _stream()), "mvlpt_ln_bwd")
    if t0 is not None:
        if dx_stream is not None:  # dy + x + dx write (+ dx read) + dx16 write
            nbytes = (2.0 + 4.0 + 4.0 + (4.0 if accumulate else 0.0) + 2.0) * rows * d
        else:                      # fp16 gradient stream: dy + x + dx16 write (+ dx16 read)
            nbytes = (2.0 + 4.0 + 2.0 + (2.0 if accumulate else 0.0)) * rows * d
        PROFILER.end("ln_bwd", t0, 0.0, nbytes)


def im2col(img, patches, B, H, W, p, Kp):
    check(_lib.lib().mvlpt_im2col(_p(img), int(img.dtype == torch.float32), _p(patches), B, H, W, p, Kp, _stream()),
          "mvlpt_im2col")


def embed_assemble(pe, cls, pos, gamma, beta, prompt, x0, B, G, v, d):
    f16 = int(prompt is not None and prompt.dtype == torch.float16)
    check(_lib.lib().mvlpt_embed_assemble(_p(pe), _p(cls), _p(pos), _p(gamma), _p(beta), _p(prompt), f16, _p(x0), B, G,
                                          v, d, LN_EPS, _stream()), "mvlpt_embed_assemble")


def set_prompt_rows(x, prompt, B, L, v, d, drop_p=0.0, seed=0, slab=0, ln=None):
    """ln=(xt, rec, gamma): also the xt / row records of the new rows (mvlpt_set_prompt_rows_ln)."""
    xt, rec, g_ = ln if ln is not None else (None, None, None)
    check(_lib.lib().mvlpt_set_prompt_rows_ln(_p(x), _p(prompt), int(prompt.dtype == torch.float16), B, L, v, d,
                                              float(drop_p), int(seed), int(slab), _p(xt), _p(rec), _p(g_), _stream()),
          "mvlpt_set_prompt_rows")


def dropout_keep(keep, B, v, d, drop_p, seed, slab):
    """uint8 [B,v,d]: the keep mask set_prompt_rows / prompt_grad use for (seed, slab)."""
    check(_lib.lib().mvlpt_dropout_keep(_p(keep), B, v, d, float(drop_p), int(seed), int(slab), _stream()),
          "mvlpt_dropout_keep")


def prompt_grad(dx, dx16, grad, B, L, v, d, inv_scale, zero_rows, drop_p=0.0, seed=0, slab=0):
    check(_lib.lib().mvlpt_prompt_grad(_p(dx), _p(dx16), _p(grad), B, L, v, d, float(inv_scale), int(zero_rows),
                                       float(drop_p), int(seed), int(slab), _stream()), "mvlpt_prompt_grad")


def text_assemble(emb, ctx, slot, pos, x0, C, Lt, n_ctx, d, csc):
    f16 = int(ctx is not None and ctx.dtype == torch.float16)
    check(_lib.lib().mvlpt_text_assemble(_p(emb), _p(ctx), f16, _p(slot), _p(pos), _p(x0), C, Lt, n_ctx, d, int(csc),
                                         _stream()), "mvlpt_text_assemble")


def ctx_grad(dx0, ctx_pos, grad, C, Lt, n_ctx, d, csc, inv_scale):
    check(_lib.lib().mvlpt_ctx_grad(_p(dx0), int(dx0.dtype == torch.float16), _p(ctx_pos), _p(grad), C, Lt, n_ctx, d,
                                    int(csc), float(inv_scale), _stream()), "mvlpt_ctx_grad")


def l2norm_fwd(x, y16, y32, inv_norm, rows, e, y16x3=None, pattern=0):
    check(_lib.lib().mvlpt_l2norm_fwd_split(_p(x), _p(y16), _p(y32), _p(inv_norm), _p(y16x3), int(pattern), rows, e,
                                            _stream()), "mvlpt_l2norm_fwd")


def l2norm_bwd(dy, y32, inv_norm, dx16, rows, e):
    check(_lib.lib().mvlpt_l2norm_bwd(_p(dy), _p(y32), _p(inv_norm), _p(dx16), rows, e, _stream()), "mvlpt_l2norm_bwd")


def ce_fwd_bwd(logits, ldc, label, soft, task, ranges, loss_rows, pred, dz16, B, C, coef, hit=None):
    check(_lib.lib().mvlpt_ce_fwd_bwd(_p(logits), ldc, _p(label), _p(soft), _p(task), _p(ranges), _p(loss_rows),
                                      _p(pred), _p(hit), _p(dz16), B, C, float(coef), _stream()), "mvlpt_ce_fwd_bwd")


def step_metrics(loss_rows, hit, B, inv_div, out2):
    check(_lib.lib().mvlpt_step_metrics(_p(loss_rows), _p(hit), B, float(inv_div), _p(out2), _stream()),
          "mvlpt_step_metrics")


def dlogits_prepare(dlogits, ld_in, task, ranges, dz16, ldc, B, C, coef):
    check(_lib.lib().mvlpt_dlogits_prepare(_p(dlogits), ld_in, _p(task), _p(ranges), _p(dz16), ldc, B, C, float(coef),
                                           _stream()), "mvlpt_dlogits_prepare")


def task_mask(logits, ldc, task, ranges, B, C):
    check(_lib.lib().mvlpt_task_mask(_p(logits), ldc, _p(task), _p(ranges), B, C, _stream()), "mvlpt_task_mask")


def transpose_f16(inp, out, R, Cc, ld_in, ld_out):
    check(_lib.lib().mvlpt_transpose_f16(_p(inp), _p(out), R, Cc, ld_in, ld_out, _stream()), "mvlpt_transpose_f16")


def sgd(p, buf, g, lr, momentum, wd, first_step):
    check(_lib.lib().mvlpt_sgd(_p(p), _p(buf), _p(g), p.numel(), int(p.dtype == torch.float16), float(lr),
                               float(momentum), float(wd), int(first_step), _stream()), "mvlpt_sgd")


def zero(t):
    check(_lib.lib().mvlpt_zero(_p(t), t.numel() * t.element_size(), _stream()), "mvlpt_zero")


# ---- CoCoOp branch (include/mvlpt_sm100.h: mvlpt_metanet_*, mvlpt_cocoop_*, mvlpt_pair_logits_*) ---------------------
def _pf16(*ts):
    dt = {t.dtype for t in ts}
    if dt == {torch.float16}:
        return 1
    if dt == {torch.float32}:
        return 0
    raise _lib.MvlptError(f"parameters of one module must be all fp16 or all fp32, got {dt}")


def vpt_proj_fwd(emb, W, b, out):
    """out[rows,d] fp32 = emb[rows,p] . W[d,p]^T + b   (vpt_proj, trainers/mvlpt.py:170-175)."""
    rows, p = emb.numel() // emb.shape[-1], emb.shape[-1]
    check(_lib.lib().mvlpt_vpt_proj_fwd(_p(emb), _p(W), _p(b), _pf16(emb, W, b), _p(out), rows, W.shape[0], p, _stream()),
          "mvlpt_vpt_proj_fwd")


def vpt_proj_bwd(d_out, emb, W, d_emb, dW, db, accumulate):
    rows, p = emb.numel() // emb.shape[-1], emb.shape[-1]
    check(_lib.lib().mvlpt_vpt_proj_bwd(_p(d_out), _p(emb), _p(W), _pf16(emb, W), _p(d_emb), _p(dW), _p(db), rows, W.shape[0], p,
                                        int(accumulate), _stream()), "mvlpt_vpt_proj_bwd")


def metanet_fwd(imf, W1, b1, W2, b2, h1, bias):
    B, e = imf.shape
    check(_lib.lib().mvlpt_metanet_fwd(_p(imf), _p(W1), _p(b1), _p(W2), _p(b2), _pf16(W1, b1, W2, b2), _p(h1), _p(bias), B, e,
                                       W1.shape[0], W2.shape[0], _stream()), "mvlpt_metanet_fwd")


def metanet_bwd(d_bias, imf, h1, W1, W2, d_h1_ws, dW1, db1, dW2, db2, d_imf, d_imf_scale):
    B, e = imf.shape
    check(_lib.lib().mvlpt_metanet_bwd(_p(d_bias), _p(imf), _p(h1), _p(W1), _p(W2), _pf16(W1, W2), _p(d_h1_ws), _p(dW1), _p(db1),
                                       _p(dW2), _p(db2), _p(d_imf), float(d_imf_scale), B, e, W1.shape[0], W2.shape[0],
                                       _stream()), "mvlpt_metanet_bwd")


def cocoop_assemble(emb, ctx, bias, slot, pos, x0, B, C, Lk, d):
    check(_lib.lib().mvlpt_cocoop_assemble(_p(emb), _p(ctx), int(ctx.dtype == torch.float16), _p(bias), _p(slot), _p(pos), _p(x0),
                                           B, C, Lk, d, _stream()), "mvlpt_cocoop_assemble")


def pair_logits_fwd(img, txt, scale, logits, ldc, B, C, e):
    check(_lib.lib().mvlpt_pair_logits_fwd(_p(img), _p(txt), float(scale), _p(logits), ldc, B, C, e, _stream()),
          "mvlpt_pair_logits_fwd")


def pair_logits_bwd(dz16, ldc, img, txt, scale, d_txt, d_img, B, C, e):
    check(_lib.lib().mvlpt_pair_logits_bwd(_p(dz16), ldc, _p(img), _p(txt), float(scale), _p(d_txt), _p(d_img), B, C, e,
                                           _stream()), "mvlpt_pair_logits_bwd")


def cocoop_ctx_grad(dx16, ctx_pos, part_ws, d_ctx, d_bias, B, C, Lk, n_ctx, d, inv_scale):
    check(_lib.lib().mvlpt_cocoop_ctx_grad(_p(dx16), _p(ctx_pos), _p(part_ws), _p(d_ctx), _p(d_bias), B, C, Lk, n_ctx, d,
                                           float(inv_scale), _stream()), "mvlpt_cocoop_ctx_grad")


def normalize_u8(src, out, mean, std):
    """uint8 [B,3,H,W] -> ToTensor + Normalize -> out fp16 / fp32 [B,3,H,W] (mvlpt_normalize_u8)."""
    import ctypes as C
    B, _, H, W = src.shape
    m = (C.c_float * 3)(*[float(x) for x in mean])
    sd = (C.c_float * 3)(*[float(x) for x in std])
    check(_lib.lib().mvlpt_normalize_u8(_p(src), _p(out), int(out.dtype == torch.float16), B, H, W, m, sd, _stream()),
          "mvlpt_normalize_u8")
    return out
