#!/usr/bin/env python
"""Throughput of the on-GPU input pipeline next to the reference's CPU stack (torchvision over Pillow).

  python tools/gpu_preprocess_bench.py [--batch 256] [--steps 20] [--h 375 --w 500]

Prints one JSON line: images/s end to end (host uint8 images -> normalised fp16 batch on the device, H2D inside), the
device-only time of the three kernels, their algorithmic HBM bytes (source box read once + uint8 intermediate written and
read once + output written once) against the measured copy bandwidth, and the CPU stack timed on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--h", type=int, default=375)
    ap.add_argument("--w", type=int, default=500)
    ap.add_argument("--cpu-images", type=int, default=64)
    a = ap.parse_args()
    from mvlpt_b200.input_pipeline import GpuTransform, CLIP_MEAN, CLIP_STD
    g = np.random.default_rng(0)
    imgs = [g.integers(0, 256, (a.h, a.w, 3), dtype=np.uint8) for _ in range(a.batch)]
    tf = GpuTransform((224, 224), CLIP_MEAN, CLIP_STD, "train", out_dtype=torch.float16)
    torch.manual_seed(0)
    for _ in range(a.warmup):
        tf(imgs)
    torch.cuda.synchronize()
    # end to end: host packing + H2D + kernels
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = tf(imgs)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / a.steps
    # device only: replay the kernels on the staged batch
    import ctypes
    from mvlpt_b200 import _lib
    L = _lib.lib()
    descs = (type(tf.last_descs[0]) * a.batch)(*tf.last_descs)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n0 = _lib.launch_count()
    ev[0].record()
    for _ in range(a.steps):
        _lib.check(L.mvlpt_preprocess(tf._dev.data_ptr(), descs, tf._dev.data_ptr(), a.batch, tf.mean, tf.std, out.data_ptr(), 1,
                                      224, 224, tf._ws.data_ptr(), tf._ws.numel(), stream))
    ev[1].record()
    torch.cuda.synchronize()
    dev = ev[0].elapsed_time(ev[1]) / a.steps / 1e3
    launches = (_lib.launch_count() - n0) // a.steps
    alg = sum(d.bh * d.bw * 3 + 2 * d.bh * 224 * 3 + 3 * 224 * 224 * 2 for d in tf.last_descs)
    peak = None
    try:
        peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs")
    except Exception:
        pass
    # the reference's CPU stack on a bounded sample, one process (a DataLoader multiplies this by its workers)
    import torchvision.transforms as T
    from PIL import Image
    stack = T.Compose([T.RandomResizedCrop((224, 224), scale=(0.08, 1.0), interpolation=T.InterpolationMode.BICUBIC),
                       T.RandomHorizontalFlip(), T.ToTensor(), T.Normalize(CLIP_MEAN, CLIP_STD)])
    pil = [Image.fromarray(x) for x in imgs[:a.cpu_images]]
    torch.set_num_threads(1)
    t0 = time.perf_counter()
    for p in pil:
        stack(p)
    cpu = (time.perf_counter() - t0) / len(pil)
    print(json.dumps({
        "metric": "input-pipeline images/sec", "workload": f"{a.batch} x {a.h}x{a.w} uint8 -> RandomResizedCrop(224, bicubic)+flip+normalize fp16",
        "e2e_images_per_s": a.batch / e2e, "e2e_ms": e2e * 1e3, "h2d_bytes": int(tf.h2d_bytes),
        "device_images_per_s": a.batch / dev, "device_ms": dev * 1e3, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": alg / dev / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": (alg / dev / 1e9 / peak) if peak else None, "algorithmic_bytes": alg},
        "cpu_baseline": {"images_per_s_per_core": 1 / cpu, "cores": 1, "kind": "reference (torchvision over Pillow)",
                         "sample": f"{len(pil)} images"}}))


if __name__ == "__main__":
    main()
