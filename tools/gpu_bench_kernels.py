"""Times the heavy kernels of the hot path in isolation at the bench shapes (run under gpurun).

  python tools/gpu_bench_kernels.py [fmha] [gemm] [ln]

CUDA events around `iters` back-to-back launches after warm-up; operands are rotated over several buffers so that
consecutive launches do not hit each other's data in L2.  Prints one line per case with us/launch and TFLOP/s or GB/s.
"""
import sys

import torch

sys.path.insert(0, ".")
from mvlpt_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
what = set(sys.argv[1:]) or {"fmha", "gemm", "ln"}
if any(x.startswith("g") and x[1:].isdigit() for x in what):
    what.add("gemm")
ONE = "one" in what  # only the first case of each family, few iterations (for ncu captures)
what.discard("one")


def timeit(fn, iters=20, warm=3):
    if ONE:
        iters, warm = 2, 1
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


if "fmha" in what:
    for (N, L, heads, causal) in [(256, 205, 12, 0), (256, 197, 12, 0), (100, 77, 8, 1), (1000, 77, 8, 1), (256, 50, 12, 0),
                                  (1000, 30, 8, 1)][:1 if ONE else None]:
        d = heads * 64
        nb = 3
        qkv = [(torch.randn(N * L, 3 * d, device=dev) * 0.7).half() for _ in range(nb)]
        out = [torch.empty(N * L, d, device=dev, dtype=torch.half) for _ in range(nb)]
        lse = [torch.empty(N, heads, L, device=dev) for _ in range(nb)]
        do = [(torch.randn(N * L, d, device=dev) * 0.3).half() for _ in range(nb)]
        dqkv = [torch.empty_like(qkv[0]) for _ in range(nb)]
        t_f = timeit(lambda i: ops.fmha_fwd(qkv[i % nb], out[i % nb], lse[i % nb], N, L, d, heads, causal))
        t_b = timeit(lambda i: ops.fmha_bwd(qkv[i % nb], out[i % nb], do[i % nb], lse[i % nb], dqkv[i % nb], N, L, d, heads,
                                            causal))
        f_f = 4.0 * N * L * L * d
        print(f"fmha N={N} L={L} heads={heads} causal={causal}: fwd {t_f:8.1f} us ({f_f / t_f / 1e6:6.1f} TF/s)   "
              f"bwd {t_b:8.1f} us ({2 * f_f / t_b / 1e6:6.1f} TF/s)", flush=True)
        del qkv, out, lse, do, dqkv

if "gemm" in what:
    only = [int(x[1:]) for x in what if x.startswith("g") and x[1:].isdigit()]
    cases = [
        # M, N, K, act, f32out, resid, aux_out
        (52480, 2304, 768, 0, 0, 0, 0), (52480, 768, 768, 0, 1, 1, 0), (52480, 3072, 768, 1, 0, 0, 1),
        (52480, 3072, 768, 1, 0, 0, 0), (52480, 768, 3072, 0, 1, 1, 0), (52480, 3072, 768, 2, 0, 0, 0),
        (52480, 768, 3072, 0, 0, 0, 0), (52480, 768, 768, 0, 0, 0, 0), (52480, 768, 2304, 0, 0, 0, 0),
        (7700, 1536, 512, 0, 0, 0, 0), (7700, 512, 512, 0, 1, 1, 0), (7700, 2048, 512, 1, 0, 0, 1),
        (7700, 512, 2048, 0, 1, 1, 0), (7700, 2048, 512, 2, 0, 0, 0), (7700, 512, 2048, 0, 0, 0, 0),
        (7700, 512, 512, 0, 0, 0, 0), (7700, 512, 1536, 0, 0, 0, 0), (77000, 2048, 512, 1, 0, 0, 1),
        # small problems: the text tower at C=100 (M=2500), a 32-image shard (M=6560)      [18..]
        (2500, 512, 2048, 0, 0, 0, 0), (2500, 512, 512, 0, 0, 0, 0), (2500, 512, 1536, 0, 0, 0, 0),
        (2500, 2048, 512, 2, 0, 0, 0), (6560, 768, 3072, 0, 0, 0, 0), (6560, 768, 768, 0, 0, 0, 0),
        (6560, 768, 2304, 0, 0, 0, 0), (6560, 3072, 768, 2, 0, 0, 0), (3750, 512, 2048, 0, 0, 0, 0),
    ]
    for (M, N, K, act, f32, resid, aux) in ([cases[0], cases[2], cases[6]] if ONE else ([cases[i] for i in only] if only else cases)):
        nb = 3
        A = [(torch.randn(M, K, device=dev) * 0.5).half() for _ in range(nb)]
        W = (torch.randn(N, K, device=dev) * 0.05).half()
        b = (torch.randn(N, device=dev) * 0.1).half()
        out = [torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.half) for _ in range(nb)]
        r = [torch.randn(M, N, device=dev) for _ in range(nb)] if resid else None
        ai = [torch.randn(M, N, device=dev).half() for _ in range(nb)] if act == 2 else None
        ao = [torch.empty(M, N, device=dev, dtype=torch.half) for _ in range(nb)] if aux else None

        def run(i):
            j = i % nb
            ops.gemm(A[j], W, out[j], bias=None if act == 2 else b, act=act, aux_in=ai[j] if ai else None,
                     aux_out=ao[j] if ao else None, resid=r[j] if r else None)

        t = timeit(run)
        print(f"gemm M={M} N={N} K={K} act={act} f32={f32} resid={resid} aux={aux}: {t:8.1f} us  "
              f"{2.0 * M * N * K / t / 1e6:7.1f} TF/s", flush=True)
        del A, out, r, ai, ao

if "lnfuse" in what:
    # explicit path (LayerNorm kernel + plain GEMMs) against the LayerNorm carried through the linears, per block half:
    #   producer = residual-stream linear [M,d] (K = d: out-proj, K = 4d: FC2), consumer = QKV (N = 3d) / FC1 (N = 4d)
    for (M, d) in [(52480, 768), (50432, 768), (25000, 512), (2500, 512), (16448, 1024)]:
        nb = 3
        g = torch.ones(d, device=dev)
        bb = torch.zeros(d, device=dev)
        x = [torch.randn(M, d, device=dev) for _ in range(nb)]
        xo = [torch.empty(M, d, device=dev) for _ in range(nb)]
        h = [torch.empty(M, d, device=dev, dtype=torch.half) for _ in range(nb)]
        rec0 = torch.zeros(M, ops.LN_REC, device=dev)
        rec1 = torch.zeros(M, ops.LN_REC, device=dev)
        ops.ln_prep(x[0], g, h[0], rec0, M, d)
        res = {}
        for K in (d, 4 * d):
            A = [(torch.randn(M, K, device=dev) * 0.5).half() for _ in range(nb)]
            W = (torch.randn(d, K, device=dev) * 0.05).half()
            b = (torch.randn(d, device=dev) * 0.1).half()
            res[f"prod K={K} plain"] = timeit(lambda i: ops.gemm(A[i % nb], W, xo[i % nb], bias=b, resid=x[i % nb]))
            res[f"prod K={K} carry"] = timeit(lambda i: ops.gemm(A[i % nb], W, xo[i % nb], bias=b, resid=x[i % nb],
                                                                 ln_prod=(rec0, rec1, g, h[i % nb])))
            del A
        res["ln_kernel"] = timeit(lambda i: ops.ln_fwd(x[i % nb], g, bb, h[i % nb], M, d))
        for N, act in ((3 * d, 0), (4 * d, 1)):
            W = (torch.randn(N, d, device=dev) * 0.05).half()
            b = (torch.randn(N, device=dev) * 0.1).half()
            sg, bp = (W.float() @ g).half(), b
            y = [torch.empty(M, N, device=dev, dtype=torch.half) for _ in range(nb)]
            t = [torch.empty(M, N, device=dev, dtype=torch.half) for _ in range(nb)] if act else None
            res[f"cons N={N} plain"] = timeit(lambda i: ops.gemm(h[i % nb], W, y[i % nb], bias=b, act=act,
                                                                 aux_out=t[i % nb] if t else None))
            res[f"cons N={N} carry"] = timeit(lambda i: ops.gemm(h[i % nb], W, y[i % nb], act=act,
                                                                 aux_out=t[i % nb] if t else None, ln_cons=(rec1, sg, bp)))
            del y, t
        print(f"lnfuse M={M} d={d}: " + "  ".join(f"{k} {v:6.1f}" for k, v in res.items()), flush=True)
        del x, xo, h

if "ln" in what:
    for (M, d) in [(52480, 768), (7700, 512)]:
        nb = 3
        x = [torch.randn(M, d, device=dev) for _ in range(nb)]
        g = torch.ones(d, device=dev)
        bb = torch.zeros(d, device=dev)
        y = [torch.empty(M, d, device=dev, dtype=torch.half) for _ in range(nb)]
        dy = [torch.randn(M, d, device=dev).half() for _ in range(nb)]
        dx = [torch.zeros(M, d, device=dev) for _ in range(nb)]
        dx16 = [torch.empty(M, d, device=dev, dtype=torch.half) for _ in range(nb)]
        t_f = timeit(lambda i: ops.ln_fwd(x[i % nb], g, bb, y[i % nb], M, d))
        t_b = timeit(lambda i: ops.ln_bwd(dy[i % nb], x[i % nb], g, dx[i % nb], dx16[i % nb], M, d, accumulate=True))
        t_h = timeit(lambda i: ops.ln_bwd(dy[i % nb], x[i % nb], g, None, dx16[i % nb], M, d, accumulate=True))
        print(f"ln M={M} d={d}: fwd {t_f:7.1f} us ({M * d * 6 / t_f / 1e3:6.0f} GB/s)  bwd {t_b:7.1f} us "
              f"({M * d * 16 / t_b / 1e3:6.0f} GB/s)  bwd fp16-stream {t_h:7.1f} us ({M * d * 10 / t_h / 1e3:6.0f} GB/s)",
              flush=True)
