"""How long does the HOST take to issue one train step's kernels, compared with how long the GPU takes to run them?
(run under gpurun)   python tools/gpu_cpu_overhead.py [coop|vpt|upt]"""
import sys, time
from types import SimpleNamespace as NS
import torch
sys.path.insert(0, ".")
import bench as B
from mvlpt_b200 import synth, _lib
from mvlpt_b200.trainers.mvlpt import MVLPT
from mvlpt_b200.trainers.runtime import DataParallelGroup

mode = sys.argv[1] if len(sys.argv) > 1 else "coop"
a = NS(mode=mode, batch=256, classes=100, ctx_len=77)
dev = torch.device("cuda:0")
sd, toks, name_lens, dm = B.make_problem(a)
tr = MVLPT(B.make_cfg(a), dm=dm, clip_state_dict=sd, device=dev, tokenized_prompts=toks, name_lens=name_lens,
           dp=DataParallelGroup())
img = synth.synth_images(256, 224, seed=1).half().to(dev)
lab = torch.randint(0, 100, (256,)).to(dev)
m = tr.model
for _ in range(3):
    m.loss_and_grads(img, lab, None)
torch.cuda.synchronize()
n = 10
l0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    m.loss_and_grads(img, lab, None)
e1.record()
t_issue = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
launches = (_lib.launch_count() - l0) / n
print(f"{mode}: host issue {t_issue * 1e3:.2f} ms/step ({launches:.0f} launches, {t_issue * 1e6 / launches:.1f} us each); "
      f"GPU {e0.elapsed_time(e1) / n:.2f} ms/step")
# one isolated step starting from an empty queue (what a synchronising trainer sees)
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.loss_and_grads(img, lab, None)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"{mode}: isolated step (queue empty at start, sync at end): {min(ts) * 1e3:.2f} ms")
