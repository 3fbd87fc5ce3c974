#!/usr/bin/env python
"""Benchmark of the MVLPT prompt-tuning hot path (BASELINE.json: "prompt-tuning images/sec ViT-B/16").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode coop|vpt|upt]
  torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL; the driver launches it this way for N > 1)

A step = one full prompt-tuning train step (forward, cross-entropy, dgrad-only backward, prompt-gradient all-reduce,
SGD) of the reference's MVLPT trainer on one synthetic batch.  Default workload = BASELINE.json configs[1]:
MVLPT-CoOp, ViT-B/16, n_ctx=16, 224x224, 256 images per GPU (weak scaling), 100 classes, L_t=77, fp16.

Prints ONE JSON line on rank 0 (see the contract in the task statement): `value` = images/s with the batch already
resident in HBM (CUDA events, max over ranks); `e2e` = the same through MVLPT.forward_backward with pinned HOST batches
(H2D copy + D2H loss read inside the timed region); `roofline` = the tcgen05 GEMM kernel's achieved TFLOP/s from
per-launch CUDA events in a second pass over the same steps; `cpu_baseline` = the CPU oracle (oracle/mvlpt_oracle.py,
a restatement of the reference's PyTorch path) on a bounded sample on this host's cores.

`--impl reference` times that CPU path alone (rank 0 only), same metric/config, and says so in the line.
"""
from __future__ import annotations

import argparse
import json
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

ARCH = "ViT-B/16"
MODES = {
    # name: (coop_n_ctx, vpt_n_ctx, deep, project_method, position)
    "coop": (16, 0, False, "identity", "end"),
    "vpt": (0, 8, True, "identity", "end"),
    "upt": (16, 8, True, "transformer", "end"),
    "cocoop": (0, 0, False, "identity", "end"),  # COCOOP.N_CTX = 4 (instance-conditioned context)
}
COCOOP_N_CTX = {"cocoop": 4}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="coop", choices=sorted(MODES))
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--ctx-len", type=int, default=77)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eval", action="store_true",
                    help="time the inference path (MVLPT.test's inner loop: parse_batch_test -> model_inference -> argmax; "
                         "SURVEY.md 8f-2) instead of the training step; not a BASELINE metric")
    return ap.parse_args()


def workload_name(a) -> str:
    n, v, deep, method, pos = MODES[a.mode]
    tag = {"coop": f"MVLPT-CoOp n_ctx={n}", "vpt": f"MVLPT-VPT-deep vctx={v}", "upt": f"MVLPT-UPT n_ctx={n}+vctx={v}",
           "cocoop": f"MVLPT-CoCoOp n_ctx={COCOOP_N_CTX.get(a.mode, 0)}"}
    ref = {"coop": "BASELINE.json configs[1] shape", "vpt": "BASELINE.json configs[2] shape",
           "upt": "BASELINE.json configs[3] shape", "cocoop": "SURVEY.md 8f-3, not a BASELINE config"}[a.mode]
    ev = " INFERENCE (MVLPT.test inner loop, SURVEY.md 8f-2)" if getattr(a, "eval", False) else ""
    return f"{tag[a.mode]}{ev} {ARCH} 224x224 batch={a.batch}/GPU C={a.classes} L_t={a.ctx_len} fp16 ({ref})"


def make_cfg(a):
    from mvlpt_b200.trainers.runtime import default_cfg
    n, v, deep, method, pos = MODES[a.mode]
    cfg = default_cfg()
    T = cfg.TRAINER.MVLPT
    T.PREC = "fp16"
    T.PROJECT_METHOD = method
    T.COOP.N_CTX, T.COOP.CLASS_TOKEN_POSITION = n, pos
    T.VPT.N_CTX, T.VPT.DEEP = v, deep
    T.COCOOP.N_CTX = COCOOP_N_CTX.get(a.mode, 0)
    cfg.DATASET.COOP = True
    cfg.MODEL.BACKBONE.NAME = ARCH
    return cfg


def make_problem(a):
    """Synthetic CLIP weights, class-name token ids, data-manager stub."""
    n, v, deep, method, pos = MODES[a.mode]
    sd = synth.synth_clip_state_dict(ARCH, seed=0)
    toks, name_lens = synth.synth_token_ids(a.classes, n or COCOOP_N_CTX.get(a.mode, 0), context_length=a.ctx_len, seed=3)
    names = [f"class{c}" for c in range(a.classes)]
    dm = NS(dataset=NS(classnames=names), lab2cname={i: nm for i, nm in enumerate(names)}, num_classes=a.classes,
            num_source_domains=1)
    return sd, toks, name_lens, dm


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


# --------------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(a, sd, toks, name_lens, B):
    """One reference train step on the CPU oracle for a B-image sample of the workload (fp32, all host threads)."""
    from oracle import mvlpt_oracle as O
    n, v, deep, method, pos = MODES[a.mode]
    cc = COCOOP_N_CTX.get(a.mode, 0)
    pp = synth.synth_prompt_params(ARCH, n, v, deep, project_dim=128 if method == "transformer" else 0, seed=0,
                                   cocoop_n_ctx=cc)
    res = synth.ARCHS[ARCH]["image_resolution"]
    image = synth.synth_images(B, res, seed=1)
    g = torch.Generator().manual_seed(2)
    label = torch.randint(0, a.classes, (B,), generator=g)
    emb = sd["token_embedding.weight"][toks]
    kw = dict(embedding=emb, eot_index=toks.argmax(-1), name_lens=name_lens, n_ctx=n, v=v, position=pos,
              upt=method == "transformer", cocoop_n_ctx=cc)
    params = [p.clone() for p in pp.values()]
    keys = list(pp)
    bufs = [None] * len(params)

    def step():
        nonlocal bufs
        cur = dict(zip(keys, params))
        _, loss, grads = O.train_step(image, label, sd, cur, **kw)
        gl = [grads.get(k, torch.zeros_like(p)) for k, p in zip(keys, params)]
        bufs = O.sgd_step(params, gl, bufs, lr=0.002)
        return float(loss)

    return step


def run_cpu(a, sd, toks, name_lens, steps, warmup, budget_s):
    """Times `steps` CPU steps on a sample batch sized so warmup+steps fit `budget_s`; returns images/s + description."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = 4
    fn = cpu_step_fn(a, sd, toks, name_lens, B)
    t0 = time.perf_counter()
    fn()
    t_probe = time.perf_counter() - t0  # includes first-touch cost: an upper bound
    # per-step cost model: text tower is per step, image tower per image -> scale only the image part up
    total = steps + warmup
    t = t_probe
    while B * 2 <= min(a.batch, 64) and t * 2 * total < budget_s:  # pessimistic: cost linear in B
        B, t = B * 2, t * 2
    fn = cpu_step_fn(a, sd, toks, name_lens, B)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    sample = (f"oracle/mvlpt_oracle.py train_step+SGD, fp32, {B} images x {a.classes} classes (L_t={a.ctx_len}) per step, "
              f"{steps} steps after {warmup} warm-up, torch CPU threads={cores}")
    return B / sec, sec, cores, sample, B


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd, toks, name_lens, dm = make_problem(a)
    ips, sec, cores, sample, B = run_cpu(a, sd, toks, name_lens, a.steps, max(1, min(a.warmup, 2)), budget_s=150.0)
    line = {
        "impl": "reference", "metric": "prompt-tuning images/sec", "value": ips, "unit": "images/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample_images_per_step": B},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- GPU arm
def ours_arm(a):
    from mvlpt_b200 import _lib, ops
    from mvlpt_b200.trainers.mvlpt import MVLPT
    from mvlpt_b200.trainers.runtime import DataParallelGroup
    from mvlpt_b200.accounting import flops_step

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dp = DataParallelGroup.from_env("nccl")
    world, rank = dp.world, dp.rank
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}")
    _lib.check(_lib.lib().mvlpt_check_device(local_rank), "mvlpt_check_device")

    sd, toks, name_lens, dm = make_problem(a)
    cfg = make_cfg(a)
    trainer = MVLPT(cfg, dm=dm, clip_state_dict=sd, device=dev, tokenized_prompts=toks, name_lens=name_lens, dp=dp)
    trainer.num_batches = 1 << 30  # never hits the per-epoch LR update inside the timed loop
    pp = synth.synth_prompt_params(ARCH, *MODES[a.mode][:3], project_dim=128 if MODES[a.mode][3] == "transformer" else 0,
                                   cocoop_n_ctx=COCOOP_N_CTX.get(a.mode, 0))
    trainer.model.prompt_learner.load_state_dict(pp, strict=False)

    n, v, deep, method, pos = MODES[a.mode]
    res = synth.ARCHS[ARCH]["image_resolution"]
    B = a.batch
    nbuf = 3  # distinct batches rotated so no step re-reads the previous step's inputs
    host_batches = []
    for i in range(nbuf):
        img = synth.synth_images(B, res, seed=100 + rank * nbuf + i).half().pin_memory()
        g = torch.Generator().manual_seed(7 + rank * nbuf + i)
        lab = torch.randint(0, a.classes, (B,), generator=g).pin_memory()
        host_batches.append({"img": img, "label": lab, "domain": torch.zeros(B, dtype=torch.long)})
    dev_batches = [{k: (t.to(dev) if k != "domain" else t) for k, t in hb.items()} for hb in host_batches]
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
        e1.record()
        torch.cuda.synchronize()
        dp.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dp.all_reduce_max(ms)
        return float(ms) / steps, (_lib.launch_count() - l0)

    for i in range(a.warmup):
        step(dev_batches[i % nbuf])
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
               "h2d_bytes_per_step": int(host_batches[0]["img"].numel() * 2 + host_batches[0]["label"].numel() * 8),
               "d2h_bytes_per_step": 8,
               "loop": "MVLPT.run_epoch's: stage_batch(i+1) (pinned host -> device on the copy stream), then step i"}

    cc = COCOOP_N_CTX.get(a.mode, 0)
    passes = B if cc else 1
    flops = flops_step(synth.ARCHS[ARCH], B, a.classes, a.ctx_len, v, n or cc, text_passes=passes)
    # rows of the causal text tower behind the last EOT are not computed (they cannot reach any output): the FLOPs
    # actually executed are reported next to the reference's algorithmic count and are the ones "achieved" uses
    Lk = int(trainer.model.prompt_learner.kernel_len)
    flops_exec = flops_step(synth.ARCHS[ARCH], B, a.classes, Lk, v, n or cc, text_passes=passes)
    if a.eval:
        from mvlpt_b200.accounting import flops_inference
        flops = flops_exec = flops_inference(synth.ARCHS[ARCH], B, a.classes, v)
    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"

    # DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the four image-tower linears at B=256, L=205 from the
    # committed `ncu --set full` capture profiles/r01_ncu_full_v4_summary.txt, next to their algorithmic bytes
    # (A + W + output [+ residual in] [+ saved pre-activation]); MB per launch.  Traffic stays below the algorithmic
    # figure (part of each output is still in the 126 MB L2 when the kernel ends): no wasted re-reads.
    NCU_TRAFFIC_MB = {"qkv 52480x2304x768": {"traffic": 307.3, "algorithmic": 326.0},
                      "out_proj 52480x768x768 (+fp32 residual)": {"traffic": 122.1, "algorithmic": 404.2},
                      "fc1 52480x3072x768 (QuickGELU, 2 outputs)": {"traffic": 692.8, "algorithmic": 730.2},
                      "fc2 52480x768x3072 (+fp32 residual)": {"traffic": 395.6, "algorithmic": 649.5}}
    roofline = None
    kernels = None
    if not a.no_roofline:
        ops.PROFILER = ops.Profiler()
        ms_prof, _ = timed(dev_batches, a.steps)
        summ = ops.PROFILER.summary()
        ops.PROFILER = None
        g = summ.get("gemm_f16_tn")
        kernels = {k: {"launches_per_step": r["launches"] / a.steps, "ms_per_step": r["ms"] / a.steps,
                       "tflops": (r["flops"] / (r["ms"] * 1e-3) / 1e12) if r["ms"] and r["flops"] else None,
                       "gbs": (r["bytes"] / (r["ms"] * 1e-3) / 1e9) if r["ms"] else None} for k, r in summ.items()}
        if g:
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "gemm_f16_tn_2sm_kernel / gemm_f16_tn_kernel (tcgen05+TMA linear, all launches of the step)",
                        "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                        "peak_source": peak_src, "launches_per_step": g["launches"] / a.steps,
                        "avg_launch_us": g["ms"] * 1e3 / g["launches"],
                        "flops_per_launch": g["flops"] / g["launches"],
                        "traffic_ncu_mb_per_launch": NCU_TRAFFIC_MB,
                        "traffic_note": "`traffic` is null because this entry averages every GEMM launch of the step; the "
                                        "per-shape DRAM traffic of the dominant launches is in traffic_ncu_mb_per_launch",
                        "share_of_step": g["ms"] / a.steps / ms_prof,
                        "timing": "CUDA events around every launch, second pass over the same steps",
                        "ms_per_step_instrumented": ms_prof}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and not a.eval:
        ips, sec, cores, sample, Bc = run_cpu(a, sd, toks, name_lens, steps=2, warmup=1, budget_s=30.0)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "inference images/sec (MVLPT.test inner loop)" if a.eval else "prompt-tuning images/sec",
            "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": f"{nbuf} distinct input batches rotated; per-step activation working set >> 126 MB L2",
                       "step_tflop_algorithmic": flops / 1e12,
                       "step_tflop_executed": flops_exec / 1e12,
                       "text_causal_cut": {"L_t": a.ctx_len, "rows_computed": Lk,
                                           "note": "rows behind the last EOT cannot influence the EOT rows of a causal "
                                                   "tower; features and gradients are bit-identical to all L_t rows "
                                                   "under the same attention tiling"},
                       "step_tflops_achieved_per_gpu": flops_exec / (ms_dev * 1e-3) / 1e12,
                       "step_frac_of_peak": flops_exec / (ms_dev * 1e-3) / 1e12 / tf_peak},
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
