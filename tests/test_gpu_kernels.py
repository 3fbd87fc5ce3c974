"""Per-kernel GPU checks through the C ABI against plain fp32 torch formulas (the ops are floating point)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K,act,resid,f32", [
    (128, 256, 64, 0, False, False), (300, 768, 768, 0, False, False), (1000, 2304, 768, 0, False, False),
    (1001, 512, 2048, 1, False, False), (515, 768, 3072, 0, True, True), (515, 3072, 768, 2, False, False),
    (33, 104, 72, 0, False, True), (52480, 768, 768, 0, True, True), (7, 8, 8, 0, False, False),
])
def test_gemm(M, N, K, act, resid, f32):
    from mvlpt_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = (torch.randn(N, device=dev) * 0.1).half()
    aux_in = torch.randn(M, N, device=dev).half() if act == 2 else None
    aux_out = torch.empty(M, N, device=dev, dtype=torch.half) if act == 1 else None
    r = torch.randn(M, N, device=dev) if resid else None
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32 if f32 else torch.half)
    ops.gemm(A, W, out, bias=b, act=act, aux_in=aux_in, aux_out=aux_out, resid=r)
    ref = A.float() @ W.float().t() + b.float()
    if act == 1:
        assert _rel(aux_out.float(), ref) < 2e-3
        ref = ref * torch.sigmoid(1.702 * ref)
    if act == 2:
        t = aux_in.float()
        s = torch.sigmoid(1.702 * t)
        ref = ref * (s * (1 + 1.702 * t * (1 - s)))
    if resid:
        ref = ref + r
    assert not torch.isnan(out).any()
    assert _rel(out.float(), ref) < 2e-3


@pytest.mark.parametrize("M,N,K,inplace", [
    (256, 768, 768, False), (410, 768, 768, False), (515, 512, 2048, True), (1000, 1024, 1024, False),
    (52480, 768, 768, True), (52480, 768, 3072, False), (25000, 512, 512, False), (300, 256, 64, False),
    (41 * 256 + 7, 768, 768, False),  # more row blocks than CTA pairs is not needed for the ragged last block
    (74 * 256 * 2 + 130, 512, 512, True),  # every pair owns several row blocks, the last one ragged
])
def test_gemm_with_fused_layernorm(M, N, K, inplace):
    """mvlpt_gemm_ln: out = A.W^T + b + resid (fp32) and h = LayerNorm(out)*gamma+beta (fp16) from one kernel, against
    torch's two-pass layer_norm of the same fp32 rows; rows with a large mean exercise the sum-of-squares statistics."""
    from mvlpt_b200 import ops
    assert ops.gemm_ln_supported(M, N)
    assert not ops.gemm_ln_supported(255, N) and not ops.gemm_ln_supported(M, 128) and not ops.gemm_ln_supported(M, 1280)
    torch.manual_seed(0)
    dev = "cuda"
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = (torch.randn(N, device=dev) * 0.1).half()
    r = torch.randn(M, N, device=dev) * 2
    r[::7] += 6.0   # rows whose mean is several standard deviations from zero
    gamma = torch.randn(N, device=dev) * 0.1 + 1
    beta = torch.randn(N, device=dev) * 0.1
    ref = A.float() @ W.float().t() + b.float() + r
    out = r.clone() if inplace else torch.full((M, N), float("nan"), device=dev)
    h = torch.full((M, N), float("nan"), device=dev, dtype=torch.half)
    ops.gemm(A, W, out, bias=b, resid=out if inplace else r, ln=(gamma, beta, h))
    assert not torch.isnan(out).any() and not torch.isnan(h).any()
    assert _rel(out, ref) < 2e-3
    # the LayerNorm of the rows the kernel itself produced: exact up to fp16 rounding of h
    want = torch.nn.functional.layer_norm(out, (N,), gamma, beta, 1e-5)
    assert _rel(h.float(), want) < 1.5e-3
    # and identical (to fp16 rounding) to the stand-alone LayerNorm kernel on the same rows
    h2 = torch.empty_like(h)
    ops.ln_fwd(out, gamma, beta, h2, M, N)
    assert (h.float() - h2.float()).abs().max() <= 2e-3 * want.abs().max()
    with pytest.raises(Exception):
        ops.gemm(A[:100], W, out[:100], bias=b, resid=r[:100], ln=(gamma, beta, h[:100]))


@pytest.mark.parametrize("N,L,heads,causal", [(3, 197, 12, 0), (2, 205, 12, 0), (2, 50, 12, 0), (1, 257, 16, 0),
                                              (5, 77, 8, 1), (4, 20, 8, 1), (2, 1, 2, 0), (2, 16, 2, 1), (40, 205, 12, 0),
                                              (3, 130, 2, 1), (2, 64, 1, 0), (1, 272, 1, 0), (2, 256, 2, 0), (3, 128, 3, 1),
                                              (2, 129, 2, 0), (2, 240, 2, 1), (300, 205, 12, 0), (150, 77, 8, 1),
                                              # longer than the single-pass kernels hold: the streaming kernels
                                              # (ViT-L/14@336px: 577 tokens + prompts)
                                              (2, 585, 16, 0), (2, 300, 2, 1), (1, 289, 1, 0), (3, 273, 2, 1), (1, 577, 3, 0),
                                              (1, 280, 2, 0),
                                              # short sequences packed several to a tile (block-diagonal mask), with
                                              # whole groups + a left-over group, causal and not
                                              (100, 13, 8, 1), (37, 25, 8, 1), (9, 13, 2, 0), (10, 13, 2, 1), (7, 30, 3, 1),
                                              (5, 64, 2, 0), (3, 50, 12, 0), (21, 9, 1, 1), (2, 2, 1, 1), (130, 1, 2, 0),
                                              (64, 16, 2, 1), (33, 20, 8, 1)])
def test_fmha_fwd_bwd(N, L, heads, causal):
    from mvlpt_b200 import ops
    torch.manual_seed(1)
    d = heads * 64
    qkv = (torch.randn(N * L, 3 * d, device="cuda") * 0.7).half()
    out = torch.empty(N * L, d, device="cuda", dtype=torch.half)
    lse = torch.empty(N, heads, L, device="cuda", dtype=torch.float32)
    ops.fmha_fwd(qkv, out, lse, N, L, d, heads, causal)
    q, k, v = [t.reshape(N, L, heads, 64).permute(0, 2, 1, 3).float().requires_grad_(True) for t in qkv.split(d, dim=1)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device="cuda").triu(1)
    p = s.softmax(-1)
    o = p @ v
    ref = o.permute(0, 2, 1, 3).reshape(N * L, d)
    assert _rel(lse, torch.logsumexp(s, -1)) < 1e-3          # QK^T + softmax statistics
    assert _rel(out.float(), ref) < 2e-3                      # PV
    do = (torch.randn(N * L, d, device="cuda") * 0.3).half()
    dqkv = torch.empty_like(qkv)
    ops.fmha_bwd(qkv, out, do, lse, dqkv, N, L, d, heads, causal)
    o.backward(do.float().reshape(N, L, heads, 64).permute(0, 2, 1, 3))
    refg = torch.cat([t.grad.permute(0, 2, 1, 3).reshape(N * L, d) for t in (q, k, v)], dim=1)
    assert _rel(dqkv.float(), refg) < 4e-3


@pytest.mark.parametrize("rows,d", [(1000, 768), (77, 512), (33, 1024), (5, 128)])
def test_layernorm_fwd_bwd(rows, d):
    from mvlpt_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(rows, d, device="cuda") * 2 + 0.3
    g = torch.randn(d, device="cuda") * 0.1 + 1
    b = torch.randn(d, device="cuda") * 0.1
    y = torch.empty(rows, d, device="cuda", dtype=torch.half)
    ops.ln_fwd(x, g, b, y, rows, d)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), g, b, 1e-5)
    assert _rel(y.float(), ref) < 2e-3
    dy = (torch.randn(rows, d, device="cuda")).half()
    dx = torch.ones(rows, d, device="cuda")
    dx16 = torch.empty(rows, d, device="cuda", dtype=torch.half)
    ops.ln_bwd(dy, x, g, dx, dx16, rows, d, accumulate=True)
    ref.backward(dy.float())
    assert _rel(dx, xr.grad + 1) < 1e-4
    assert _rel(dx16.float(), xr.grad + 1) < 2e-3
    # fp16 gradient stream (dx_stream = None): the running sum is read from / written to dx16 alone
    run16 = torch.full((rows, d), 0.5, device="cuda", dtype=torch.half)
    ops.ln_bwd(dy, x, g, None, run16, rows, d, accumulate=True)
    assert _rel(run16.float(), xr.grad + 0.5) < 2e-3
    ops.ln_bwd(dy, x, g, None, run16, rows, d, accumulate=False)
    assert _rel(run16.float(), xr.grad) < 2e-3


def test_ce_softlabels_taskmask_and_metrics():
    from mvlpt_b200 import ops
    torch.manual_seed(3)
    B, C, ldc = 37, 50, 56
    z = torch.zeros(B, ldc, device="cuda")
    z[:, :C] = torch.randn(B, C, device="cuda") * 3
    soft = torch.zeros(B, C, device="cuda")
    idx = torch.randint(0, C, (B, 2), device="cuda")
    soft[torch.arange(B), idx[:, 0]] = 1
    soft[torch.arange(B), idx[:, 1]] = 1
    ranges = torch.tensor([[0, 20], [20, 50]], dtype=torch.int32, device="cuda")
    task = torch.randint(0, 2, (B,), device="cuda", dtype=torch.int32)
    zin = z.clone()
    loss_rows = torch.empty(B, device="cuda")
    pred = torch.empty(B, device="cuda", dtype=torch.int32)
    hit = torch.empty(B, device="cuda", dtype=torch.int32)
    dz = torch.empty(B, ldc, device="cuda", dtype=torch.half)
    ops.ce_fwd_bwd(z, ldc, None, soft, task, ranges, loss_rows, pred, dz, B, C, coef=4096.0 / B, hit=hit)
    mask = torch.zeros(B, C, device="cuda")
    for b in range(B):
        lo, hi = ranges[task[b]].tolist()
        mask[b, lo:hi] = 1
    zr = (zin[:, :C] * mask).requires_grad_(True)
    y = soft / soft.sum(-1, keepdim=True)
    loss = torch.nn.functional.cross_entropy(zr, y)
    loss.backward()
    assert _rel(loss_rows.mean(), loss.detach()) < 1e-5
    assert _rel(dz[:, :C].float() / 4096.0, zr.grad * mask) < 2e-3
    assert torch.equal(pred.long(), zr.argmax(-1))
    out2 = torch.empty(2, device="cuda")
    ops.step_metrics(loss_rows, hit, B, 1.0 / B, out2)
    acc = 100.0 * (zr.argmax(-1) == y.argmax(-1)).float().mean()
    assert abs(float(out2[0]) - float(loss)) < 1e-4 and abs(float(out2[1]) - float(acc)) < 1e-3


def test_sgd_kernel_matches_torch():
    from mvlpt_b200 import ops
    torch.manual_seed(4)
    p = torch.randn(1000, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([ref], lr=0.002, momentum=0.9, weight_decay=5e-4)
    buf = torch.zeros_like(p)
    for step in range(3):
        g = torch.randn(1000, device="cuda")
        ref.grad = g.clone()
        opt.step()
        ops.sgd(p, buf, g, 0.002, 0.9, 5e-4, first_step=step == 0)
    assert torch.allclose(p, ref.detach(), atol=1e-6)


def test_errors_are_loud():
    from mvlpt_b200 import ops
    from mvlpt_b200._lib import MvlptError
    A = torch.zeros(8, 12, device="cuda", dtype=torch.half)  # lda = 12: not a multiple of 8
    W = torch.zeros(8, 12, device="cuda", dtype=torch.half)
    out = torch.zeros(8, 8, device="cuda", dtype=torch.half)
    with pytest.raises(MvlptError):
        ops.gemm(A, W, out)
    with pytest.raises(MvlptError):
        ops.gemm(A.cpu(), W, out)
    qkv = torch.zeros(4, 3 * 100, device="cuda", dtype=torch.half)
    with pytest.raises(MvlptError):
        ops.fmha_fwd(qkv, qkv, qkv, 1, 4, 100, 2, 0)  # d != heads*64


@pytest.mark.parametrize("B,res,p,dtype", [(3, 224, 16, torch.float16), (2, 224, 32, torch.float16), (2, 224, 14, torch.float16),
                                           (2, 64, 16, torch.float32)])
def test_im2col_matches_unfold(B, res, p, dtype):
    """patch gather of the stride-p conv (clip/model.py:207): the vectorised fp16 path and the scalar one"""
    from mvlpt_b200 import ops
    torch.manual_seed(4)
    img = torch.randn(B, 3, res, res, device="cuda").to(dtype)
    g = res // p
    K = 3 * p * p
    Kp = (K + 7) // 8 * 8
    patches = torch.full((B * g * g, Kp), 7.0, device="cuda", dtype=torch.half)
    ops.im2col(img, patches, B, res, res, p, Kp)
    ref = torch.nn.functional.unfold(img.float(), kernel_size=p, stride=p).transpose(1, 2).reshape(B * g * g, K)
    assert torch.equal(patches[:, :K].float(), ref.half().float())
    assert (patches[:, K:] == 0).all()


@pytest.mark.parametrize("f16_stream", [False, True])
def test_prompt_grad_reduces_over_batch_and_clears_rows(f16_stream):
    from mvlpt_b200 import ops
    torch.manual_seed(5)
    B, L, v, d = 37, 21, 4, 768
    dx = torch.randn(B * L, d, device="cuda")
    dx16 = dx.half()
    ref = dx16.float().view(B, L, d)[:, 1:1 + v].sum(0) * 0.25 if f16_stream else dx.view(B, L, d)[:, 1:1 + v].sum(0) * 0.25
    grad = torch.empty(v, d, device="cuda")
    ops.prompt_grad(None if f16_stream else dx, dx16, grad, B, L, v, d, 0.25, zero_rows=True)
    assert _rel(grad, ref) < 1e-5
    assert (dx16.view(B, L, d)[:, 1:1 + v] == 0).all() and (dx16.view(B, L, d)[:, 0] != 0).any()
    if not f16_stream:
        assert (dx.view(B, L, d)[:, 1:1 + v] == 0).all()
    g2 = torch.empty(v, d, device="cuda")
    ops.prompt_grad(None if f16_stream else dx, dx16, g2, B, L, v, d, 1.0, zero_rows=False)
    assert (g2 == 0).all()
