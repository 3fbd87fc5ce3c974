"""Data-parallel consistency check on >= 2 GPUs (torchrun): one train step of MVLPT-UPT with the class-sharded text tower
must give the same loss and the same all-reduced prompt gradients as the replicated text tower, and both must match a
single-rank step over the whole global batch.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_dp_check.py
"""
import os
import sys
from types import SimpleNamespace as NS

import torch

sys.path.insert(0, ".")
from mvlpt_b200 import synth  # noqa: E402
from mvlpt_b200.trainers.mvlpt import MVLPT  # noqa: E402
from mvlpt_b200.trainers.runtime import DataParallelGroup, default_cfg  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dp = DataParallelGroup.from_env("nccl")
ARCH = "ViT-B/16"
C, n, v, Bl = 37, 4, 2, 6  # 37 classes: ragged shards


def build(dp_group, csc, cocoop=False):
    cfg = default_cfg()
    T = cfg.TRAINER.MVLPT
    T.PREC = "fp16"
    T.PROJECT_METHOD = "identity"
    T.COOP.N_CTX, T.COOP.CLASS_TOKEN_POSITION, T.COOP.CSC = n, "end", csc
    if cocoop:  # instance-conditioned branch: no class sharding, B*C text sequences per rank
        T.COOP.N_CTX, T.COCOOP.N_CTX, T.COCOOP.PREC = 0, n, "fp16"
    T.VPT.N_CTX, T.VPT.DEEP = v, True
    cfg.DATASET.COOP = True
    cfg.MODEL.BACKBONE.NAME = ARCH
    sd = synth.synth_clip_state_dict(ARCH, seed=0)
    toks, name_lens = synth.synth_token_ids(C, n, context_length=24, seed=3)
    names = [f"class{c}" for c in range(C)]
    dm = NS(dataset=NS(classnames=names), lab2cname=dict(enumerate(names)), num_classes=C, num_source_domains=1)
    tr = MVLPT(cfg, dm=dm, clip_state_dict=sd, device=dev, tokenized_prompts=toks, name_lens=name_lens, dp=dp_group)
    torch.manual_seed(5)
    for _, p in tr.model._trainables():
        p.data.copy_((torch.randn(p.shape) * 0.02).to(p.dtype))
    return tr


def step(tr, image, label, world):
    tr.model.loss_and_grads(image, label, None, global_batch=image.shape[0] * world)
    flat = tr.model.grad_buffer().clone()
    tr.dp.all_reduce_sum(flat)
    torch.cuda.synchronize()
    return flat


res = synth.ARCHS[ARCH]["image_resolution"]
g = torch.Generator().manual_seed(11)
images = (torch.randn(dp.world * Bl, 3, res, res, generator=g) * 0.5).half().to(dev)
labels = torch.randint(0, C, (dp.world * Bl,), generator=g).to(dev)
mine = slice(dp.rank * Bl, (dp.rank + 1) * Bl)
ok = True
for csc in (False, True):
    tr = build(dp, csc)
    tr.model.shard_text = True
    g_sh = step(tr, images[mine], labels[mine], dp.world)
    tr.model.shard_text = False
    g_rep = step(tr, images[mine], labels[mine], dp.world)
    # single-rank reference over the whole global batch (no collectives): a fresh group object with world = 1
    one = DataParallelGroup.__new__(DataParallelGroup)
    one.enabled, one.world, one.rank, one.group, one.dist = False, 1, 0, None, None
    tr1 = build(one, csc)
    g_one = step(tr1, images, labels, 1)
    den = g_one.abs().max().item()
    e1 = (g_sh - g_rep).abs().max().item() / den
    e2 = (g_sh - g_one).abs().max().item() / den
    good = e1 < 5e-3 and e2 < 5e-3
    ok &= good
    if dp.rank == 0:
        print(f"csc={csc}: |sharded - replicated| = {e1:.2e}, |sharded - single rank| = {e2:.2e} (relative to max |g|) "
              f"{'OK' if good else 'BAD'}", flush=True)
# CoCoOp: per-rank images, all-reduced gradients (context, meta-net weights, visual prompts) vs one rank over the global batch
tr = build(dp, False, cocoop=True)
g_dp = step(tr, images[mine], labels[mine], dp.world)
one = DataParallelGroup.__new__(DataParallelGroup)
one.enabled, one.world, one.rank, one.group, one.dist = False, 1, 0, None, None
tr1 = build(one, False, cocoop=True)
g_one = step(tr1, images, labels, 1)
views = tr1.model.grad_views()
off, worst = 0, 0.0
for k, t in views.items():  # per-tensor scale: the meta-net gradients are orders of magnitude apart
    a, b = g_dp[off:off + t.numel()], g_one[off:off + t.numel()]
    worst = max(worst, ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item())
    off += t.numel()
good = worst < 2e-2
ok &= good
if dp.rank == 0:
    print(f"cocoop: max over tensors |dp - single rank| / max|g| = {worst:.2e} {'OK' if good else 'BAD'}", flush=True)
dp.barrier()
if dp.rank == 0:
    print("ALL OK" if ok else "FAILED", flush=True)
torch.distributed.destroy_process_group()
sys.exit(0 if ok else 1)
