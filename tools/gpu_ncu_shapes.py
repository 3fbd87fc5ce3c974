"""One launch of every GEMM shape of a ViT-B/16 image-tower layer at the bench batch (B=256, L=205: M=52480) — forward
with the LayerNorm carry, dgrad-only backward — for an `ncu --set full -k regex:gemm_f16` capture; the shape keys are
written in launch order to gpurun_out/gemm_order_shapes.json (tools/ncu_traffic.py matches the capture to them).
  python tools/gpu_ncu_shapes.py [M] [d]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from mvlpt_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 52480
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda:0")
torch.manual_seed(0)
h = lambda *s: (torch.randn(*s, device=dev) * 0.1).half()
g = torch.ones(d, device=dev)
x, xo = torch.randn(M, d, device=dev), torch.empty(M, d, device=dev)
xt = torch.empty(M, d, device=dev, dtype=torch.half)
rec0, rec1 = torch.zeros(M, ops.LN_REC, device=dev), torch.zeros(M, ops.LN_REC, device=dev)
ops.ln_prep(x, g, xt, rec0, M, d)
w_qkv, w_o, w_fc, w_pr = h(3 * d, d), h(d, d), h(4 * d, d), h(d, 4 * d)
qkv, o, t, gg = h(M, 3 * d), h(M, d), h(M, 4 * d), h(M, 4 * d)
dh, dt, dqkv = h(M, d), h(M, 4 * d), h(M, 3 * d)
sg3, bp3, sg4, bp4 = h(3 * d), h(3 * d), h(4 * d), h(4 * d)
b_d = h(d)
torch.cuda.synchronize()
ops.GEMM_LOG = []
ops.gemm(xt, w_qkv, qkv, ln_cons=(rec0, sg3, bp3))                                        # QKV
ops.gemm(o, w_o, xo, bias=b_d, resid=x, ln_prod=(rec0, rec1, g, xt))                      # out-proj (+ xt, records)
ops.gemm(xt, w_fc, gg, act=ops.ACT_QUICKGELU, aux_out=t, ln_cons=(rec1, sg4, bp4))       # FC1 (training: saves t)
ops.gemm(xt, w_fc, gg, act=ops.ACT_QUICKGELU, ln_cons=(rec1, sg4, bp4))                  # FC1 (inference)
ops.gemm(gg, w_pr, xo, bias=b_d, resid=x, ln_prod=(rec1, rec0, g, xt))                    # FC2 (+ xt, records)
ops.gemm(dh, w_pr.t().contiguous(), dt, act=ops.ACT_MUL_DQUICKGELU, aux_in=t)             # dgrad FC2 (* QuickGELU')
ops.gemm(dt, w_fc.t().contiguous(), dh)                                                   # dgrad FC1
ops.gemm(dh, w_o.t().contiguous(), o)                                                     # dgrad out-proj
ops.gemm(dqkv, w_qkv.t().contiguous(), dh)                                                # dgrad QKV
torch.cuda.synchronize()
json.dump(ops.GEMM_LOG, open(f"gpurun_out/gemm_order_shapes_{M}.json", "w"))
print(len(ops.GEMM_LOG), "gemm launches logged")
