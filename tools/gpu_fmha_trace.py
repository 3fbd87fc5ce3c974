"""In-kernel timeline of the attention backward (debug build: MVLPT_NVCC_EXTRA=-DMVLPT_FMHA_DBG python -m mvlpt_b200.build).
Prints, for CTA 0, the (tag, t_us) sequence of the MMA thread and of the first row thread for the first units."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from mvlpt_b200 import ops, _lib

N, L, heads, causal = 256, 205, 12, 0
d = heads * 64
dev = "cuda:0"
qkv = (torch.randn(N * L, 3 * d, device=dev) * 0.7).half()
out = torch.empty(N * L, d, device=dev, dtype=torch.half)
lse = torch.empty(N, heads, L, device=dev)
do = (torch.randn(N * L, d, device=dev) * 0.3).half()
dqkv = torch.empty_like(qkv)
ops.fmha_fwd(qkv, out, lse, N, L, d, heads, causal)
Lb = _lib.lib()
buf = (ctypes.c_ulonglong * (2 * 2048))()
cnt = (ctypes.c_int * 2)()
for rep in range(2):
    ops.fmha_bwd(qkv, out, do, lse, dqkv, N, L, d, heads, causal)
    torch.cuda.synchronize()
    Lb.mvlpt_dbg_fmha_trace(buf, cnt)
t0 = min(buf[1], buf[2048 + 1])
names = {1: "mma:wait_ld", 2: "mma:ld_ok", 3: "mma:wait_p", 4: "mma:p_ok", 10: "row:unit", 11: "row:D_done", 12: "row:D_bar",
         13: "row:wait_st", 14: "row:st_ok", 15: "row:p_arrive", 16: "row:acc_ok", 17: "row:kv_store", 18: "row:dq_store"}
for s in range(2):
    print(f"--- stream {s}: {cnt[s]} events")
    prev = None
    for i in range(min(cnt[s], 110)):
        tag, t = buf[s * 2048 + 2 * i], buf[s * 2048 + 2 * i + 1]
        us = (t - t0) / 1e3
        print(f"{str(names.get(tag, tag)):14s} {us:9.2f} us  (+{0 if prev is None else us - prev:6.2f})")
        prev = us
