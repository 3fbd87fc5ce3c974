"""Is the (small-kernel) text tower host-bound?  Times text forward+backward alone: host issue time vs GPU time, and the
same captured in a CUDA graph.  (run under gpurun)"""
import sys, time
from types import SimpleNamespace as NS
import torch
sys.path.insert(0, ".")
import bench as B
from mvlpt_b200 import synth, _lib
from mvlpt_b200.trainers.mvlpt import MVLPT
from mvlpt_b200.trainers.runtime import DataParallelGroup

a = NS(mode="coop", batch=256, classes=int(sys.argv[1]) if len(sys.argv) > 1 else 100, ctx_len=77)
dev = torch.device("cuda:0")
sd, toks, name_lens, dm = B.make_problem(a)
tr = MVLPT(B.make_cfg(a), dm=dm, clip_state_dict=sd, device=dev, tokenized_prompts=toks, name_lens=name_lens,
           dp=DataParallelGroup())
m = tr.model
pl = m.prompt_learner
img = synth.synth_images(256, 224, seed=1).half().to(dev)
lab = torch.randint(0, a.classes, (256,)).to(dev)
m.loss_and_grads(img, lab, None)
torch.cuda.synchronize()
tt = m.text_encoder.tower(dev)
C, Lk = pl.n_cls, pl.kernel_len
ctx = pl.ctx.detach().contiguous()
head = m.head(dev)
bf = head.buffers(256, C)
d_ctx = m.grad_views()["ctx"]


def text_fb():
    tt.forward(pl._emb, ctx, pl._slot, pl._eot_rows, pl.coop_n_ctx, pl.csc, train=True)
    tt.backward(bf["dtfeat16"], C, Lk, pl._eot_rows, pl._ctx_pos, pl.coop_n_ctx, pl.csc, d_ctx, 1.0)


for _ in range(3):
    text_fb()
torch.cuda.synchronize()
n = 20
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    text_fb()
e1.record()
host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
launches = (_lib.launch_count() - l0) / n
print(f"text fwd+bwd C={C} rows={Lk}: {launches:.0f} launches; host issue {host * 1e3:.2f} ms, GPU (back to back) "
      f"{e0.elapsed_time(e1) / n:.2f} ms")
# the same in a CUDA graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    text_fb()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    text_fb()
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"  as a CUDA graph: GPU {e0.elapsed_time(e1) / n:.2f} ms per replay")
