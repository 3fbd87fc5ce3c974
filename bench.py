#!/usr/bin/env python
"""Benchmark of the MVLPT prompt-tuning hot path (BASELINE.json: "prompt-tuning images/sec ViT-B/16").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--scaling weak|strong]
  torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL; the driver launches it this way for N > 1)

A step = one full prompt-tuning train step (forward, cross-entropy, dgrad-only backward, prompt-gradient all-reduce,
SGD) of the reference's MVLPT trainer on one synthetic batch.  `--config` selects one of BASELINE.json's configurations
(default 2 = configs[1], the one the metric is quoted on at one GPU):

  1  CoOp ViT-B/32 n_ctx=4, Caltech-101 shape (C=100), batch 1                       (configs[0]; the CPU anchor)
  2  MVLPT-CoOp ViT-B/16 n_ctx=16, 224x224, batch 256, C=100, L_t=77, 'end'          (configs[1])
  3  MVLPT-VPT-deep ViT-B/16 vctx=8 x 12 layers, 11-task label space C=2193 (the reference's class lists, tokenised by
     its BPE: tests/golden/b16_vpt_deep_11task.pt), one task id per sample, per-task logit mask, L_t=77    (configs[2])
  4  MVLPT-UPT ViT-B/16 n_ctx=16 + vctx=8, transformer projection, ImageNet-1k names (tests/golden/
     b16_coop_c1000_cut.pt), 'middle', the scripts' context cut L_t=30 (`--ctx-len 77` for the uncut variant) (configs[3])
  5  MVLPT-CoOp ViT-L/14 n_ctx=16, C=1000, L_t=77, batch 512 over 8 GPUs = 64 per GPU                       (configs[4])

`--scaling weak` (default) keeps `--batch` images per GPU; `--scaling strong` splits `--batch` over the ranks.

Prints ONE JSON line on rank 0 (see the contract in the task statement): `value` = images/s with the batch already
resident in HBM (CUDA events, max over ranks); `e2e` = the same through MVLPT.forward_backward with pinned HOST batches
(H2D copy + D2H loss read inside the timed region); `roofline` = the dominant tcgen05 GEMM shape's achieved TFLOP/s from
per-launch CUDA events in a second pass over the same steps, with its DRAM traffic from the committed ncu capture;
`cpu_baseline` = the reference's own modules (oracle/_ref, built by oracle/build_ref.py) on a bounded sample on this
host's cores; `config.eager_fp16_b200_images_per_s` = the same reference modules in fp16 PyTorch-eager on this GPU.

`--impl reference` times the reference alone (rank 0 only), same metric/config, and says so in the line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace as NS

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

from mvlpt_b200 import synth  # noqa: E402

MODES = {
    # name: (coop_n_ctx, vpt_n_ctx, deep, project_method)
    "coop": (16, 0, False, "identity"),
    "vpt": (0, 8, True, "identity"),
    "upt": (16, 8, True, "transformer"),
    "cocoop": (0, 0, False, "identity"),  # COCOOP.N_CTX = 4 (instance-conditioned context)
}
COCOOP_N_CTX = {"cocoop": 4}

CONFIGS = {
    1: dict(mode="coop", arch="ViT-B/32", batch=1, classes=100, ctx_len=77, n_ctx=4, position="end", labels="synthetic",
            ref="BASELINE.json configs[0]"),
    2: dict(mode="coop", arch="ViT-B/16", batch=256, classes=100, ctx_len=77, position="end", labels="synthetic",
            ref="BASELINE.json configs[1]"),
    3: dict(mode="vpt", arch="ViT-B/16", batch=256, classes=2193, ctx_len=77, position="end", labels="11task",
            ref="BASELINE.json configs[2]"),
    4: dict(mode="upt", arch="ViT-B/16", batch=256, classes=1000, ctx_len=30, position="middle", labels="imagenet",
            ref="BASELINE.json configs[3]"),
    5: dict(mode="coop", arch="ViT-L/14", batch=64, classes=1000, ctx_len=77, position="end", labels="imagenet",
            ref="BASELINE.json configs[4] (512 images over 8 GPUs)"),
}
LABEL_FIXTURE = {"imagenet": "b16_coop_c1000_cut", "11task": "b16_vpt_deep_11task"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--mode", default=None, choices=sorted(MODES))
    ap.add_argument("--arch", default=None, choices=[k for k in synth.ARCHS if k != "tiny"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (weak) / in total (strong)")
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--ctx-len", type=int, default=None)
    ap.add_argument("--position", default=None, choices=["end", "middle", "front"])
    ap.add_argument("--labels", default=None, choices=["synthetic", "imagenet", "11task"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-settle", action="store_true", help="skip the ~1.2 s of extra warm-up steps (profiler runs)")
    ap.add_argument("--watchdog-s", type=float, default=1500.0,
                    help="GPU arm: abort the process (exit 124) if no result line after this many seconds; 0 = off")
    ap.add_argument("--gemm-log", default="", help="write the ordered shape keys of every GEMM launch of the process to "
                                                   "this JSON file (tools/ncu_traffic.py matches an ncu capture to it)")
    ap.add_argument("--eval", action="store_true",
                    help="time the inference path (MVLPT.test's inner loop: parse_batch_test -> model_inference -> argmax; "
                         "SURVEY.md 8f-2) instead of the training step; not a BASELINE metric")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    explicit = {k: getattr(a, k) is not None for k in ("mode", "arch", "batch", "classes", "ctx_len", "position", "labels")}
    for k in ("mode", "arch", "batch", "classes", "ctx_len", "position", "labels"):
        if getattr(a, k) is None:
            setattr(a, k, c[k])
    if explicit["classes"] and not explicit["labels"] and a.classes != c["classes"]:
        a.labels = "synthetic"
    a.n_ctx_override = c.get("n_ctx") if not explicit["mode"] else None
    a.config_ref = c["ref"] if not any(explicit.values()) else f"{c['ref']} with overrides " + \
        ",".join(k for k, v in explicit.items() if v)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    a.global_batch = a.batch * world if a.scaling == "weak" else a.batch
    if a.global_batch % world:
        raise SystemExit(f"--scaling strong: batch {a.batch} does not divide over {world} ranks")
    a.local_batch = a.global_batch // world
    return a


def mode_of(a):
    n, v, deep, method = MODES[a.mode]
    if a.n_ctx_override and n:
        n = a.n_ctx_override
    return n, v, deep, method


def workload_name(a) -> str:
    n, v, deep, method = mode_of(a)
    tag = {"coop": f"MVLPT-CoOp n_ctx={n}", "vpt": f"MVLPT-VPT-deep vctx={v}", "upt": f"MVLPT-UPT n_ctx={n}+vctx={v}",
           "cocoop": f"MVLPT-CoCoOp n_ctx={COCOOP_N_CTX.get(a.mode, 0)}"}[a.mode]
    ev = " INFERENCE (MVLPT.test inner loop, SURVEY.md 8f-2)" if getattr(a, "eval", False) else ""
    res = synth.ARCHS[a.arch]["image_resolution"]
    lab = {"synthetic": "random class-name tokens", "imagenet": "ImageNet-1k names (reference BPE)",
           "11task": "11-task label space (reference BPE), task id per sample, per-task logit mask"}[a.labels]
    return (f"{tag}{ev} {a.arch} {res}x{res} batch={a.local_batch}/GPU C={a.classes} L_t={a.ctx_len} "
            f"'{a.position}' fp16, {lab} ({a.config_ref})")


def make_cfg(a):
    from mvlpt_b200.trainers.runtime import default_cfg
    n, v, deep, method = mode_of(a)
    cfg = default_cfg()
    T = cfg.TRAINER.MVLPT
    T.PREC = "fp16"
    T.PROJECT_METHOD = method
    T.COOP.N_CTX, T.COOP.CLASS_TOKEN_POSITION = n, a.position
    T.VPT.N_CTX, T.VPT.DEEP = v, deep
    T.COCOOP.N_CTX = COCOOP_N_CTX.get(a.mode, 0)
    cfg.DATASET.COOP = True
    cfg.DATASET.MULTITASK = a.labels == "11task"
    cfg.DATASET.MULTITASK_LABEL_PERTASK = a.labels == "11task"
    cfg.MODEL.BACKBONE.NAME = a.arch
    res = synth.ARCHS[a.arch]["image_resolution"]
    cfg.INPUT.SIZE = (res, res)
    cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD = CLIP_MEAN, CLIP_STD
    return cfg


def make_problem(a):
    """Synthetic CLIP weights, class-name token ids, data-manager stub, per-task class counts (11-task only)."""
    n, v, deep, method = mode_of(a)
    n_text = n or COCOOP_N_CTX.get(a.mode, 0)
    sd = synth.synth_clip_state_dict(a.arch, seed=0)
    task_sizes = None
    if a.labels == "synthetic":
        toks, name_lens = synth.synth_token_ids(a.classes, n_text, context_length=a.ctx_len, seed=3)
    else:
        fx = torch.load(REPO / "tests" / "golden" / f"{LABEL_FIXTURE[a.labels]}.pt", map_location="cpu", weights_only=False)
        toks, name_lens = fx["tokenized_prompts"].long(), list(fx["name_lens"])
        fx_n = fx["case"].get("coop_n_ctx", 0)
        if fx_n != n_text:
            raise SystemExit(f"--labels {a.labels} was tokenised with {fx_n} context placeholders, this mode uses {n_text}")
        if toks.shape[0] != a.classes:
            raise SystemExit(f"--labels {a.labels} has {toks.shape[0]} classes, not {a.classes}")
        used = int((toks != 0).sum(1).max())
        if a.ctx_len < used:
            raise SystemExit(f"--ctx-len {a.ctx_len} is shorter than the longest prompt ({used})")
        if toks.shape[1] >= a.ctx_len:
            toks = toks[:, :a.ctx_len].contiguous()
        else:
            toks = torch.cat([toks, torch.zeros(toks.shape[0], a.ctx_len - toks.shape[1], dtype=torch.long)], 1)
        task_sizes = fx["case"].get("tasks") if a.labels == "11task" else None
    names = [f"class{c}" for c in range(a.classes)]
    dm = NS(dataset=NS(classnames=names), lab2cname={i: nm for i, nm in enumerate(names)}, num_classes=a.classes,
            num_source_domains=1)
    if task_sizes:
        tn = [f"task{i}" for i in range(len(task_sizes))]
        dm._num_classes, dm._task_names = a.classes, tn
        dm._labelmap = {t: list(range(s)) for t, s in zip(tn, task_sizes)}
    return NS(sd=sd, toks=toks, name_lens=name_lens, dm=dm, task_sizes=task_sizes)


CLIP_MEAN = [0.48145466, 0.4578275, 0.40821073]  # configs/trainers/MVLPT/vit_b16.yaml:10-11
CLIP_STD = [0.26862954, 0.26130258, 0.27577711]


def make_batch(a, prob, B, seed, as_u8=False):
    """One synthetic batch in the reference's CoOp-data format {'img','label','domain'} (trainers/mvlpt.py:953-968); with
    the 11-task label space 'domain' is the task id and the label is uniform inside that task's class range."""
    res = synth.ARCHS[a.arch]["image_resolution"]
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    u8 = torch.randint(0, 256, (B, 3, res, res), generator=g, dtype=torch.uint8)  # synthetic 8-bit RGB crops
    if as_u8:
        img = u8
    else:  # ToTensor + Normalize with CLIP's statistics, as torchvision computes them
        mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
        img = (u8.float().div(255) - mean) / std
    g = torch.Generator().manual_seed(7 + seed)
    if prob.task_sizes:
        sizes = torch.tensor(prob.task_sizes)
        starts = torch.cumsum(sizes, 0) - sizes
        task = torch.randint(0, len(sizes), (B,), generator=g)
        lab = starts[task] + (torch.rand(B, generator=g) * sizes[task]).long().clamp_max(sizes[task] - 1)
    else:
        task = torch.zeros(B, dtype=torch.long)
        lab = torch.randint(0, a.classes, (B,), generator=g)
    return img, lab, task


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def report(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------- reference arms
def reference_available() -> bool:
    from oracle import build_ref
    return build_ref.available()


def reference_step_fn(a, prob, B, device="cpu", fp16=False, seed=1):
    """One train step of the UNMODIFIED reference (oracle/_ref: clip/model.py + trainers/mvlpt.py byte-compiled by
    oracle/build_ref.py): its CustomCLIP forward, F.cross_entropy, autograd backward and torch.optim.SGD with Dassl's
    defaults (lr 0.002, momentum 0.9, weight decay 5e-4; trainers/mvlpt.py:910-951), on a B-image sample of the workload
    with the SAME synthetic weights, class-token ids and prompt parameters as the CUDA arm."""
    from oracle import build_ref
    cm, tm = build_ref.import_reference()
    n, v, deep, method = mode_of(a)
    cc = COCOOP_N_CTX.get(a.mode, 0)
    arch = synth.ARCHS[a.arch]
    clip_model = cm.CLIP(**arch)
    clip_model.load_state_dict(prob.sd)
    clip_model.eval()
    if fp16:
        cm.convert_weights(clip_model)  # what clip.build_model does (clip/model.py:430), the default PREC="fp16" path
    res = arch["image_resolution"]
    prec = "fp16" if fp16 else "fp32"
    cfg = NS(TRAINER=NS(MVLPT=NS(PREC=prec, PROJECT_METHOD=method, PROJECT_DIM=128,
                                 VPT=NS(N_CTX=v, CTX_INIT="", DROPOUT=0.0, PROJECT=-1, DEEP=deep),
                                 COOP=NS(N_CTX=n, CTX_INIT="", CSC=False, CLASS_TOKEN_POSITION=a.position),
                                 COCOOP=NS(N_CTX=cc, CTX_INIT="", PREC=prec)),
                        CUT_CONTEXTLEN=a.ctx_len < 77, ACT_CKPT=1),
             INPUT=NS(SIZE=(res, res)), DATASET=NS(MULTITASK_LABEL_PERTASK=bool(prob.task_sizes)))
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints its prompt template
        model = tm.CustomCLIP(cfg, [f"c{i}" for i in range(a.classes)], clip_model, dm=prob.dm if prob.task_sizes else None)
    pl = model.prompt_learner
    # identical inputs: the workload's token ids replace the ones the constructor derived from the placeholder names
    n_text = n or cc
    with torch.no_grad():
        emb = clip_model.token_embedding(prob.toks).type(clip_model.dtype)
    pl.token_prefix, pl.token_suffix = emb[:, :1, :].clone(), emb[:, 1 + n_text:, :].clone()
    pl.tokenized_prompts = model.tokenized_prompts = prob.toks
    pl.name_lens = list(prob.name_lens)
    pp = synth.synth_prompt_params(a.arch, n, v, deep, project_dim=128 if method == "transformer" else 0, seed=0,
                                   cocoop_n_ctx=cc)
    missing, unexpected = pl.load_state_dict(pp, strict=False)
    assert not unexpected, unexpected
    for name, p in model.named_parameters():
        p.requires_grad_("prompt_learner" in name)
    model.to(device)
    optim = torch.optim.SGD([p for p in pl.parameters() if p.requires_grad], lr=0.002, momentum=0.9, weight_decay=5e-4)
    img, lab, task = make_batch(a, prob, B, seed)
    img = (img.half() if fp16 else img).to(device)
    lab = lab.to(device)
    task = task if prob.task_sizes else None
    F = torch.nn.functional

    def step():
        out = model(img, task=task)
        loss = F.cross_entropy(out, lab)
        optim.zero_grad()
        loss.backward()
        optim.step()
        return float(loss.item())

    return step


def port_step_fn(a, prob, B):
    """Fallback when oracle/_ref was not built: the oracle restatement (oracle/mvlpt_oracle.py) + its SGD."""
    from oracle import mvlpt_oracle as O
    n, v, deep, method = mode_of(a)
    cc = COCOOP_N_CTX.get(a.mode, 0)
    pp = synth.synth_prompt_params(a.arch, n, v, deep, project_dim=128 if method == "transformer" else 0, seed=0,
                                   cocoop_n_ctx=cc)
    image, label, task = make_batch(a, prob, B, 1)
    emb = prob.sd["token_embedding.weight"][prob.toks]
    kw = dict(embedding=emb, eot_index=prob.toks.argmax(-1), name_lens=prob.name_lens, n_ctx=n, v=v, position=a.position,
              upt=method == "transformer", cocoop_n_ctx=cc)
    if prob.task_sizes:
        ends = torch.cumsum(torch.tensor(prob.task_sizes), 0)
        kw.update(task=task, task_ranges=torch.stack([ends - torch.tensor(prob.task_sizes), ends], 1))
    params = [p.clone() for p in pp.values()]
    keys = list(pp)
    bufs = [None] * len(params)

    def step():
        nonlocal bufs
        _, loss, grads = O.train_step(image, label, prob.sd, dict(zip(keys, params)), **kw)
        gl = [grads.get(k, torch.zeros_like(p)) for k, p in zip(keys, params)]
        bufs = O.sgd_step(params, gl, bufs, lr=0.002)
        return float(loss)

    return step


def run_cpu(a, prob, steps, warmup, budget_s):
    """Times `steps` CPU steps of the reference on a sample batch sized so warmup+steps fit `budget_s`."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    use_ref = reference_available()
    make = (lambda B: reference_step_fn(a, prob, B)) if use_ref else (lambda B: port_step_fn(a, prob, B))
    B = min(4, a.local_batch)
    fn = make(B)
    t0 = time.perf_counter()
    fn()
    t_probe = time.perf_counter() - t0  # includes first-touch cost: an upper bound
    total = steps + warmup
    t = t_probe
    while B * 2 <= min(a.local_batch, 64) and t * 2 * total < budget_s:  # pessimistic: cost linear in B
        B, t = B * 2, t * 2
    fn = make(B)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    what = ("the reference's own CustomCLIP + F.cross_entropy + torch.optim.SGD (oracle/_ref, unmodified modules)" if use_ref
            else "oracle/mvlpt_oracle.py train_step + SGD (port; oracle/_ref not built)")
    sample = (f"{what}, fp32, {B} images x {a.classes} classes (L_t={a.ctx_len}) per step, same weights/tokens/prompts as "
              f"the CUDA arm, {steps} steps after {warmup} warm-up, torch CPU threads={cores}")
    return dict(ips=B / sec, sec=sec, cores=cores, sample=sample, B=B, kind="reference" if use_ref else "port")


def run_eager_gpu(a, prob, steps=8, warmup=3):
    """SURVEY.md §8d's secondary, honest baseline: the same unmodified reference modules in fp16 PyTorch-eager on cuda:0,
    full local batch, device-resident inputs, CUDA-event timed.  None when oracle/_ref is absent or it does not fit."""
    if not (reference_available() and torch.cuda.is_available()):
        return None
    try:
        fn = reference_step_fn(a, prob, a.local_batch, device="cuda", fp16=True)
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"images_per_s": a.local_batch / (ms * 1e-3), "ms_per_step": ms, "batch": a.local_batch, "steps": steps,
               "what": "reference clip/model.py + trainers/mvlpt.py (oracle/_ref), convert_weights fp16, PyTorch-eager "
                       f"{torch.__version__} on cuda:0, autograd backward + torch.optim.SGD, loss.item() per step"}
    except torch.cuda.OutOfMemoryError as e:
        out = {"images_per_s": None, "error": f"out of memory at batch {a.local_batch}: {str(e)[:120]}"}
    finally:
        fn = None
        torch.cuda.empty_cache()
    return out


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = make_problem(a)
    r = run_cpu(a, prob, a.steps, max(1, min(a.warmup, 2)), budget_s=150.0)
    eager = None if a.no_eager_baseline else run_eager_gpu(a, prob)
    line = {
        "impl": "reference", "metric": "prompt-tuning images/sec", "value": r["ips"], "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample_images_per_step": r["B"],
                   "same_config": r["B"] == a.local_batch,
                   "sample_note": f"CPU steps run {r['B']} of the {a.local_batch} images of a step (class count, sequence "
                                  f"lengths, weights and prompts identical); images/s is the normalised figure",
                   "eager_fp16_b200_images_per_s": None if not eager else eager.get("images_per_s"),
                   "eager_fp16_b200": eager},
        "cpu_baseline": {"value": r["ips"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["ips"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- GPU arm
def ncu_traffic_table():
    """{shape key: DRAM bytes per launch} parsed from the committed `ncu --set full` capture summary
    (profiles/ncu_gemm_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of the same bench command)."""
    p = REPO / "profiles" / "ncu_gemm_traffic.json"
    if not p.exists():
        return {}, None
    d = json.loads(p.read_text())
    return d.get("shapes", {}), d.get("source")


def arm_watchdog(seconds):
    """Bound a hung run: ranks stuck in a mismatched collective spin on their GPUs until the CALLER's limit expires (that is
    how the round's last 8-GPU session was lost).  A daemon timer ends this process instead; a finished run never sees it."""
    if seconds <= 0:
        return

    def fire():
        sys.stderr.write(f"bench.py: watchdog: no result line after {seconds:g} s, aborting (exit 124)\n")
        sys.stderr.flush()
        os._exit(124)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def settle_steps(dp, step_s, dev, target_s=1.2, cap=512):
    """How many steps make up `target_s` seconds, THE SAME NUMBER ON EVERY RANK: the slowest rank's measured seconds per
    step is taken (max all-reduce) and every rank derives the count from that one value."""
    t = torch.tensor([float(step_s)], device=dev, dtype=torch.float64)
    dp.all_reduce_max(t)
    return max(0, min(cap, int(math.ceil(target_s / max(float(t), 1e-4)))))


def ours_arm(a):
    from mvlpt_b200 import _lib, ops
    from mvlpt_b200.trainers.mvlpt import MVLPT
    from mvlpt_b200.trainers.runtime import DataParallelGroup
    from mvlpt_b200.accounting import flops_step, text_flops

    arm_watchdog(a.watchdog_s)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dp = DataParallelGroup.from_env("nccl")
    world, rank = dp.world, dp.rank
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}")
    _lib.check(_lib.lib().mvlpt_check_device(local_rank), "mvlpt_check_device")

    if a.gemm_log:
        ops.GEMM_LOG = []
    prob = make_problem(a)
    cfg = make_cfg(a)
    trainer = MVLPT(cfg, dm=prob.dm, clip_state_dict=prob.sd, device=dev, tokenized_prompts=prob.toks,
                    name_lens=prob.name_lens, dp=dp)
    trainer.num_batches = 1 << 30  # never hits the per-epoch LR update inside the timed loop
    n, v, deep, method = mode_of(a)
    cc = COCOOP_N_CTX.get(a.mode, 0)
    pp = synth.synth_prompt_params(a.arch, n, v, deep, project_dim=128 if method == "transformer" else 0, cocoop_n_ctx=cc)
    trainer.model.prompt_learner.load_state_dict(pp, strict=False)

    B = a.local_batch
    nbuf = 3  # distinct batches rotated so no step re-reads the previous step's inputs
    # host batches: 8-bit RGB crops at the model's size in pinned memory (what a loader holds after decode + crop + resize);
    # ToTensor + Normalize run on the device (mvlpt_normalize_u8), so one byte per value crosses PCIe.  The device-resident
    # batches of the `value` loop are those same images, already normalised.
    host_batches, dev_batches = [], []
    for i in range(nbuf):
        img, lab, task = make_batch(a, prob, B, seed=100 + rank * nbuf + i, as_u8=True)
        host_batches.append({"img": img.pin_memory(), "label": lab.pin_memory(), "domain": task})
        dev_batches.append({"img": trainer._normalize_u8(img.to(dev)), "label": lab.to(dev), "domain": task})
    torch.cuda.synchronize()

    if a.eval:
        trainer.set_model_mode("eval")
        trainer.model.hold_text_features(True)  # as MVLPT.test does: the text features are constant during an evaluation

    def step(batch):
        if not a.eval:
            return trainer.forward_backward(batch)
        with torch.no_grad():
            inp, label, task = trainer.parse_batch_test(batch)
            pred = trainer.model_inference(inp, task=task).argmax(dim=1)
        return (pred == label).sum()

    def timed(batches, steps, lookahead=False):
        """`lookahead`: the loop of MVLPT.run_epoch — the host->device copy of batch i+1 is issued (stage_batch) before
        step i is enqueued; every step's copy still happens inside the timed region."""
        dp.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        nxt = trainer.stage_batch(batches[0]) if lookahead else None
        for i in range(steps):
            if lookahead:
                cur = nxt
                nxt = trainer.stage_batch(batches[(i + 1) % len(batches)]) if i + 1 < steps else None
                out = step(cur)
                continue
            out = step(batches[i % len(batches)])
        if a.eval:
            out.item()  # the evaluator's read of the batch result
        if hasattr(out, "resolve"):
            out.resolve()  # the host reads the last step's loss / accuracy (every step's were copied back asynchronously)
        e1.record()
        torch.cuda.synchronize()
        dp.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dp.all_reduce_max(ms)
        return float(ms) / steps, (_lib.launch_count() - l0)

    for i in range(a.warmup):
        step(dev_batches[i % nbuf])
    torch.cuda.synchronize()
    # thermal settle: a B200 under its 1 kW cap sheds ~5 % of its clocks over the first second of sustained load; the timed
    # regions below (device-resident, end to end, instrumented) run back to back, so the first one would otherwise see a
    # cooler chip than the others.  Extra warm-up steps (never fewer than --warmup) worth ~1.2 s of steps.  Their NUMBER
    # is agreed across ranks (settle_steps): every training step issues a gradient all-reduce, so ranks that each ran
    # "until 1.2 s have passed" on their own clocks would issue different numbers of collectives and deadlock.
    settle = 0
    if not a.no_settle:
        probe = 4
        t0 = time.perf_counter()
        for i in range(probe):
            step(dev_batches[i % nbuf])
        torch.cuda.synchronize()
        settle = max(probe, settle_steps(dp, (time.perf_counter() - t0) / probe, dev))
        for i in range(settle - probe):
            step(dev_batches[i % nbuf])
            if i % 8 == 7:
                torch.cuda.synchronize()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches = timed(dev_batches, a.steps)
    sampler.stop()
    clocks = sampler.report()
    value = world * B / (ms_dev * 1e-3)

    e2e = None
    if not a.no_e2e:
        for i in range(2):
            step(host_batches[i % nbuf])
        ms_e2e, _ = timed(host_batches, a.steps, lookahead=True)
        e2e = {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(host_batches[0]["img"].numel() * host_batches[0]["img"].element_size()
                                         + host_batches[0]["label"].numel() * 8),
               "input": "uint8 [B,3,H,W] crops in pinned host memory; ToTensor + Normalize on the device",
               "d2h_bytes_per_step": 8,
               "loop": "MVLPT.run_epoch's: stage_batch(i+1) (pinned host -> device on the copy stream), then step i; the "
                       "loss / accuracy of every step are copied to pinned host memory asynchronously and the host reads "
                       "the last one before the region ends"}

    # ---- FLOP accounting: the reference's algorithmic count, and what this step actually executes ----
    arch = synth.ARCHS[a.arch]
    passes = B if cc else 1
    Lk = int(trainer.model.prompt_learner.kernel_len)
    C_local = a.classes  # per rank the text tower covers classes/world under class sharding; FLOPs are per GPU below
    txt_sharded = world > 1 and trainer.model.shard_text and (n or cc)
    C_exec = -(-a.classes // world) if txt_sharded else a.classes
    flops_alg = flops_step(arch, B, a.classes, a.ctx_len, v, n or cc, text_passes=passes)
    text_cached = (n == 0 and cc == 0)  # no text-side parameter trains: features computed once, then held
    flops_exec = flops_step(arch, B, 0 if text_cached else C_exec, Lk, v, n or cc, text_passes=passes, head_classes=a.classes)
    skipped = {
        "text_forward_cached_tflop": text_flops(arch, a.classes, a.ctx_len, False) / 1e12 if text_cached else 0.0,
        "text_causal_cut_tflop": 0.0 if text_cached else
        (text_flops(arch, C_exec, a.ctx_len, bool(n or cc)) - text_flops(arch, C_exec, Lk, bool(n or cc))) * passes / 1e12,
        "text_class_sharding_tflop": 0.0 if not txt_sharded else
        (text_flops(arch, a.classes, a.ctx_len, True) - text_flops(arch, C_exec, a.ctx_len, True)) / 1e12,
        "note": "result-identical work the reference performs every step and this path does not (SURVEY.md App. D): text "
                "features of a label space no parameter of which trains are computed once; rows behind the last EOT of a "
                "causal tower cannot reach an EOT row; under data parallelism each rank encodes its shard of the classes",
    }
    if a.eval:
        from mvlpt_b200.accounting import flops_inference
        flops_alg = flops_exec = flops_inference(arch, B, a.classes, v)
        skipped = None
    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"

    roofline = None
    kernels = None
    if not a.no_roofline:
        ops.PROFILER = ops.Profiler()
        ms_prof, _ = timed(dev_batches, a.steps)
        summ = ops.PROFILER.summary()
        ops.PROFILER = None
        kernels = {k: {"launches_per_step": r["launches"] / a.steps, "ms_per_step": r["ms"] / a.steps,
                       "tflops": (r["flops"] / (r["ms"] * 1e-3) / 1e12) if r["ms"] and r["flops"] else None,
                       "gbs": (r["bytes"] / (r["ms"] * 1e-3) / 1e9) if r["ms"] else None} for k, r in summ.items()}
        g_all = summ.get("gemm_f16_tn")
        shapes = {k: r for k, r in summ.items() if k.startswith("gemm[")}
        if shapes:
            # the dominant kernel = the GEMM shape with the largest share of the step
            key, g = max(shapes.items(), key=lambda kv: kv[1]["ms"])
            traffic_tab, traffic_src = ncu_traffic_table()
            tr = traffic_tab.get(key)
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": f"gemm_f16_tn_2sm_kernel {key} (tcgen05 + TMA linear; the shape with "
                                                     "the largest share of the step)",
                        "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                        "traffic": None if tr is None else tr["dram_bytes"],
                        "traffic_source": traffic_src if tr is not None else "no ncu capture of this shape committed",
                        "algorithmic_bytes": g["bytes"] / g["launches"],
                        "peak_source": peak_src, "launches_per_step": g["launches"] / a.steps,
                        "avg_launch_us": g["ms"] * 1e3 / g["launches"],
                        "flops_per_launch": g["flops"] / g["launches"],
                        "share_of_step": g["ms"] / a.steps / ms_prof,
                        "all_gemm_launches": None if not g_all else {
                            "achieved": g_all["flops"] / (g_all["ms"] * 1e-3) / 1e12,
                            "frac": g_all["flops"] / (g_all["ms"] * 1e-3) / 1e12 / tf_peak,
                            "launches_per_step": g_all["launches"] / a.steps,
                            "share_of_step": g_all["ms"] / a.steps / ms_prof},
                        "timing": "CUDA events around every launch, second pass over the same steps",
                        "ms_per_step_instrumented": ms_prof}

    cpu = eager = None
    if rank == 0 and world == 1 and not a.eval:
        if not a.no_eager_baseline:
            del trainer
            torch.cuda.empty_cache()
            eager = run_eager_gpu(a, prob)
        if not a.no_cpu_baseline:
            r = run_cpu(a, prob, steps=2, warmup=1, budget_s=30.0)
            cpu = {"value": r["ips"], "unit": "images/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    if a.gemm_log and rank == 0:
        Path(a.gemm_log).write_text(json.dumps(ops.GEMM_LOG))
    if rank == 0:
        line = {
            "metric": "inference images/sec (MVLPT.test inner loop)" if a.eval else "prompt-tuning images/sec",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": f"{nbuf} distinct input batches rotated; per-step activation working set >> 126 MB L2",
                       "thermal_settle_steps": settle,
                       "step_tflop_algorithmic": flops_alg / 1e12,
                       "step_tflop_executed": flops_exec / 1e12,
                       "skipped": skipped,
                       "text_rows": {"L_t": a.ctx_len, "rows_computed": Lk},
                       "step_tflops_achieved_per_gpu": flops_exec / (ms_dev * 1e-3) / 1e12,
                       "step_frac_of_peak": flops_exec / (ms_dev * 1e-3) / 1e12 / tf_peak,
                       "eager_fp16_b200_images_per_s": None if not eager else eager.get("images_per_s"),
                       "eager_fp16_b200": eager},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    dp.barrier()
    if dp.enabled:
        torch.distributed.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours_arm(a)


if __name__ == "__main__":
    main()
