"""CPU oracle for the input pipeline in front of the hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(mvlpt_b200/) never does.

What it restates.  The reference turns a decoded PIL image into the `[3,S,S]` tensor the image tower eats with torchvision
transforms over Pillow:
  * ELEVATER path — trainers/vision_benchmark/evaluation/feature.py:540-553: `Resize(SIZE, BICUBIC)` [+ `CenterCrop`] →
    `ToTensor` → `Normalize(PIXEL_MEAN, PIXEL_STD)`;
  * CoOp-data path — `dassl.data.transforms.build_transform` (trainers/mvlpt.py:17; Dassl is an unpinned git dependency,
    requirements.txt:21, absent here; its published behaviour for configs/trainers/MVLPT/vit_b16.yaml:8-13 is
    `RandomResizedCrop(SIZE, scale=(0.08,1), BICUBIC)` → `RandomHorizontalFlip` → `ToTensor` → `Normalize` for training and
    `Resize(max(SIZE), BICUBIC)` → `CenterCrop(SIZE)` → `ToTensor` → `Normalize` for testing).
The arithmetic lives in two third-party dependencies that are not under /root/reference: **Pillow** (pinned 8.3.1,
requirements.txt:7; 12.2.0 in this image) — `Image.resize` = `ImagingResample` in src/libImaging/Resample.c: separable
two-pass convolution, horizontal first, with a uint8 intermediate, double-precision bicubic (a = -0.5) coefficients whose
support grows with the down-scaling factor, normalised, converted to 22-bit fixed point, accumulated in int32 with a rounding
constant and clipped to 8 bits — and **torchvision** (pinned 0.11.0, env_mvlpt.yml:61; 0.26.0 here) for the crop-box
draws, ToTensor (`/255` in fp32) and Normalize (`(x - mean) / std` in fp32).

Pinning: tests/test_preprocess.py checks this restatement BIT-EXACTLY against Pillow + torchvision themselves (both present
in this image and on the GPU box) over random image sizes, crops, up- and down-scaling, and against the committed fixture
tests/golden/preprocess.pt (oracle/gen_golden_preprocess.py)."""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c: fixed-point fraction bits of the 8-bit path
BICUBIC_SUPPORT = 2.0


def bicubic_filter(x: float) -> float:
    """Resample.c `bicubic_filter`, a = -0.5 (Keys)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the whole axis (box = [0, in_size]).
    -> bounds int32 [out, 2] (first tap, tap count), kk int32 [out, ksize] (22-bit fixed point)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for t in w:
            ww += t
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_last_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """img uint8 [..., in_size, C] -> uint8 [..., out_size, C] (one pass of ImagingResample{Horizontal,Vertical}_8bpc)."""
    in_size = img.shape[-2]
    bounds, kk = precompute_coeffs(in_size, out_size)
    out = np.empty(img.shape[:-2] + (out_size, img.shape[-1]), np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (src[..., xmin:xmin + n, :] * kk[xx, :n].astype(np.int64)[:, None]).sum(axis=-2) + (1 << (PRECISION_BITS - 1))
        out[..., xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)  # clip8: arithmetic shift, then clamp
    return out


def resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`PIL.Image.resize((out_w, out_h), BICUBIC)` of a uint8 [H, W, C] image: horizontal pass, then vertical."""
    tmp = _resample_last_axis(img, out_w)                                  # [H, out_w, C]
    return _resample_last_axis(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2)


def to_tensor_normalize(img: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """torchvision ToTensor (`uint8 -> float32, / 255`) then Normalize (`(x - mean) / std`, fp32) -> float32 [C, H, W]."""
    x = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def resized_crop(img: np.ndarray, box: Tuple[int, int, int, int], size: Tuple[int, int], flip: bool) -> np.ndarray:
    """torchvision `F.resized_crop` (crop FIRST: the filter never sees pixels outside the box) + `F.hflip`.
    box = (top, left, height, width); size = (S_h, S_w) -> uint8 [S_h, S_w, C]."""
    i, j, h, w = box
    out = resize_bicubic(img[i:i + h, j:j + w], size[0], size[1])
    return out[:, ::-1] if flip else out


def resize_center_crop(img: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """`Resize(max(size))` (shorter edge -> max(size), the other `int(s * long / short)`) then `CenterCrop(size)`."""
    H, W = img.shape[:2]
    s = max(size)
    if W <= H:
        rw, rh = s, int(s * H / W)
    else:
        rh, rw = s, int(s * W / H)
    r = resize_bicubic(img, rh, rw)
    top = int(round((rh - size[0]) / 2.0))
    left = int(round((rw - size[1]) / 2.0))
    return r[top:top + size[0], left:left + size[1]]
