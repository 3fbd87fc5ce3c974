"""GPU sanity check of the tcgen05 GEMM against torch (run under gpurun)."""
import ctypes, sys, time
import torch
sys.path.insert(0, ".")
from mvlpt_b200 import _lib

L = _lib.lib()
_lib.check(L.mvlpt_check_device(0), "check_device")
dev = torch.device("cuda:0")
torch.manual_seed(0)

def ptr(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None

def run(M, N, K, act=0, bias=True, resid=False, out_f32=False, aux_out=False, alpha=1.0):
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = (torch.randn(N, device=dev) * 0.1).half() if bias else None
    aux_in = (torch.randn(M, N, device=dev)).half() if act == 2 else None
    aux_o = torch.empty(M, N, device=dev, dtype=torch.half) if aux_out else None
    r = torch.randn(M, N, device=dev) if resid else None
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32 if out_f32 else torch.half)
    d = _lib.GemmDesc(M, N, K, K, K, N, N, act, int(out_f32), alpha)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.mvlpt_gemm(ctypes.byref(d), ptr(A), ptr(W), ptr(b), ptr(aux_in), ptr(aux_o), ptr(r), ptr(out), stream)
    _lib.check(rc, "gemm")
    torch.cuda.synchronize()
    ref = alpha * (A.float() @ W.float().t())
    if bias: ref = ref + b.float()
    if act == 1:
        t = ref
        ref = t * torch.sigmoid(1.702 * t)
    if act == 2:
        t = aux_in.float(); s = torch.sigmoid(1.702 * t)
        ref = ref * (s * (1 + 1.702 * t * (1 - s)))
    if resid: ref = ref + r
    err = (out.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
    ok = err < 2e-3 and not torch.isnan(out).any().item()
    extra = ""
    if aux_out:
        e2 = (aux_o.float() - t).abs().max().item() / t.abs().max().item()
        ok = ok and e2 < 2e-3
        extra = f" aux_err={e2:.2e}"
    print(f"{'OK ' if ok else 'BAD'} M={M} N={N} K={K} act={act} bias={bias} resid={resid} f32={out_f32} err={err:.3e}{extra}", flush=True)
    return ok

ok = True
ok &= run(128, 256, 64)
ok &= run(128, 256, 128)
ok &= run(256, 512, 768)
ok &= run(300, 768, 768)
ok &= run(1000, 2304, 768)
ok &= run(77 * 13, 512, 2048, act=1, aux_out=True)
ok &= run(515, 3072, 768, act=1)
ok &= run(515, 768, 3072, resid=True, out_f32=True)
ok &= run(515, 3072, 768, act=2, bias=False)
ok &= run(64, 128, 128, bias=False)
ok &= run(33, 104, 72, alpha=3.5, out_f32=True)
ok &= run(50432, 768, 768, resid=True, out_f32=True)

# timing: the block linears at B=256, L=197 in every epilogue mode the engine uses
def bench(M, N, K, iters=20, act=0, resid=False, f32=False, aux_out=False):
    A = (torch.randn(M, K, device=dev) * 0.5).half(); W = (torch.randn(N, K, device=dev) * 0.05).half()
    b = torch.zeros(N, device=dev).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.half)
    r = torch.randn(M, N, device=dev) if resid else None
    aux_in = torch.randn(M, N, device=dev).half() if act == 2 else None
    aux_o = torch.empty(M, N, device=dev, dtype=torch.half) if aux_out else None
    d = _lib.GemmDesc(M, N, K, K, K, N, N, act, int(f32), 1.0)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.mvlpt_gemm(ctypes.byref(d), ptr(A), ptr(W), ptr(b), ptr(aux_in), ptr(aux_o), ptr(r), ptr(out), stream)
    for _ in range(3):
        _lib.check(call(), "gemm")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2 * M * N * K / ms / 1e9
    print(f"bench M={M} N={N} K={K} act={act} resid={int(resid)} f32={int(f32)} aux_out={int(aux_out)}: {ms:.3f} ms = {tf:.0f} TFLOP/s", flush=True)

import os
print("MVLPT_GEMM_STAGES =", os.environ.get("MVLPT_GEMM_STAGES"))
bench(50432, 2304, 768)
bench(50432, 768, 768)
bench(50432, 768, 768, resid=True, f32=True)
bench(50432, 768, 3072, resid=True, f32=True)
bench(50432, 3072, 768, act=1)
bench(50432, 3072, 768, act=1, aux_out=True)
bench(50432, 3072, 768, act=2)
bench(50432, 768, 3072)
bench(7700, 2048, 512, act=1, aux_out=True)
bench(7700, 512, 2048, resid=True, f32=True)
bench(7700, 512, 512, resid=True, f32=True)
bench(8192, 8192, 8192)
A = torch.randn(8192, 8192, device=dev).half(); W = torch.randn(8192, 8192, device=dev).half()
for _ in range(3): torch.nn.functional.linear(A, W)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): torch.nn.functional.linear(A, W)
e1.record(); torch.cuda.synchronize()
print(f"torch (cuBLAS) 8192^3: {2*8192**3/(e0.elapsed_time(e1)/10)/1e9:.0f} TFLOP/s")
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
