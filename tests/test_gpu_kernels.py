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


@pytest.mark.parametrize("M,d,inplace", [(256, 768, False), (410, 768, False), (515, 512, True), (1000, 1024, False),
                                         (52480, 768, True), (25000, 512, False), (300, 256, False),
                                         (74 * 512 + 130, 512, True)])
def test_layernorm_carried_through_the_linears(M, d, inplace):
    """mvlpt_gemm_ln: the chain  ln_prep -> consumer GEMM -> ... -> producer GEMM (+xt, row records) -> consumer GEMM
    against the explicit path (LayerNorm kernel, plain GEMMs) and against torch's two-pass layer_norm in fp32.  Rows with
    a large mean exercise the centring; the second consumer reads what the producer's epilogue wrote."""
    from mvlpt_b200 import ops
    F = torch.nn.functional
    assert ops.gemm_ln_supported(M, d)
    assert not ops.gemm_ln_supported(255, d) and not ops.gemm_ln_supported(M, 128) and not ops.gemm_ln_supported(M, 1280)
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(M, d, device=dev) * 2
    x[::7] += 6.0   # rows whose mean is several standard deviations from zero
    g1, b1 = torch.randn(d, device=dev) * 0.1 + 1, torch.randn(d, device=dev) * 0.1
    g2, b2 = torch.randn(d, device=dev) * 0.1 + 1, torch.randn(d, device=dev) * 0.1
    W1 = (torch.randn(3 * d, d, device=dev) * d ** -0.5).half()    # "QKV"
    bias1 = (torch.randn(3 * d, device=dev) * 0.1).half()
    Wo = (torch.randn(d, 3 * d, device=dev) * (3 * d) ** -0.5).half()  # stands in for attention + out-proj
    bo = (torch.randn(d, device=dev) * 0.1).half()
    W2 = (torch.randn(4 * d, d, device=dev) * d ** -0.5).half()    # "FC1"
    bias2 = (torch.randn(4 * d, device=dev) * 0.1).half()
    sg1, bp1 = (W1.float() @ g1).half(), (bias1.float() + W1.float() @ b1).half()
    sg2, bp2 = (W2.float() @ g2).half(), (bias2.float() + W2.float() @ b2).half()
    xt = torch.full((M, d), float("nan"), device=dev, dtype=torch.half)
    rec0 = torch.full((M, ops.LN_REC), float("nan"), device=dev)
    rec1 = torch.full((M, ops.LN_REC), float("nan"), device=dev)
    # --- consumer after ln_prep
    ops.ln_prep(x, g1, xt, rec0, M, d)
    y1 = torch.full((M, 3 * d), float("nan"), device=dev, dtype=torch.half)
    ops.gemm(xt, W1, y1, ln_cons=(rec0, sg1, bp1))
    want1 = F.layer_norm(x, (d,), g1, b1, 1e-5) @ W1.float().t() + bias1.float()
    assert not torch.isnan(y1).any()
    assert _rel(y1.float(), want1) < 2e-3
    h = torch.empty(M, d, device=dev, dtype=torch.half)
    ops.ln_fwd(x, g1, b1, h, M, d)
    y1e = torch.empty_like(y1)
    ops.gemm(h, W1, y1e, bias=bias1)                     # the explicit path: same accuracy class
    assert _rel(y1e.float(), want1) < 2e-3
    assert _rel(y1.float(), want1) < 1.5 * _rel(y1e.float(), want1) + 2e-4
    # --- producer: x2 = x + y1.Wo^T + bo, plus xt / records for the LayerNorm (g2, b2) that follows
    x2 = x.clone() if inplace else torch.full((M, d), float("nan"), device=dev)
    ops.gemm(y1, Wo, x2, bias=bo, resid=x2 if inplace else x, ln_prod=(rec0, rec1, g2, xt))
    want_x2 = x + y1.float() @ Wo.float().t() + bo.float()
    assert not torch.isnan(x2).any() and not torch.isnan(xt).any()
    assert _rel(x2, want_x2) < 1e-3
    used = 2 * (d // 128)
    assert not torch.isnan(rec1[:, :used]).any() and not torch.isnan(rec1[:, 16]).any()
    mean = rec1[:, 16] + rec1[:, 0:used:2].sum(1) / d
    var = rec1[:, 1:used:2].sum(1) / d - (rec1[:, 0:used:2].sum(1) / d) ** 2
    assert _rel(mean, x2.mean(1)) < 1e-5
    assert _rel(var, x2.var(1, unbiased=False)) < 1e-4
    assert (rec1[:, 16] - x.mean(1)).abs().max() < 1e-4 * x.abs().max()   # centred on the residual row's mean
    # --- consumer of the producer's output
    y2 = torch.full((M, 4 * d), float("nan"), device=dev, dtype=torch.half)
    t2 = torch.full((M, 4 * d), float("nan"), device=dev, dtype=torch.half)
    ops.gemm(xt, W2, y2, act=ops.ACT_QUICKGELU, aux_out=t2, ln_cons=(rec1, sg2, bp2))
    pre = F.layer_norm(x2, (d,), g2, b2, 1e-5) @ W2.float().t() + bias2.float()
    assert _rel(t2.float(), pre) < 2e-3
    assert _rel(y2.float(), pre * torch.sigmoid(1.702 * pre)) < 2e-3
    with pytest.raises(Exception):   # below one 256-row tile the CTA-pair kernel (and with it the carry) is not available
        ops.gemm(xt[:100], W1, y1[:100], ln_cons=(rec0, sg1, bp1))


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
