"""On-GPU input pipeline in front of the image tower: decoded uint8 RGB images -> the normalised
`[B,3,S,S]` batch, bit-identical to the reference's CPU transforms.

What it mirrors
  * `construct_dataloader`'s `transform_clip` (trainers/vision_benchmark/evaluation/feature.py:540-553):
    `Resize(SIZE, BICUBIC)` (a stretch) or, with `DATASET.CENTER_CROP`, `Resize(SIZE[0])` + `CenterCrop(SIZE)`; then
    `ToTensor`, `Normalize(PIXEL_MEAN, PIXEL_STD)`                                      -> `elevater_transform(cfg)`
  * Dassl's `build_transform(cfg, is_train)` for the choices the MVLPT configs use (configs/trainers/MVLPT/vit_b16.yaml:
    8-13: `random_resized_crop`, `random_flip`, `normalize`; test = `Resize(max(SIZE))` + `CenterCrop(SIZE)`)
                                                                                       -> `build_transform(cfg, is_train)`
The random draws (crop box, flip) are made on the host from torch's global CPU generator in exactly the order and with
exactly the calls torchvision's `RandomResizedCrop.get_params` / `RandomHorizontalFlip.forward` make, so a seeded
single-worker reference pipeline and this one see the same crops; the pixels go through `mvlpt_preprocess`
(csrc/preprocess.cu).  There is no CPU path: without CUDA / the library this raises."""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class ImageDesc(ctypes.Structure):
    """mvlpt_image_desc (include/mvlpt_sm100.h)."""
    _fields_ = [("src_off", ctypes.c_uint64), ("H", ctypes.c_int), ("W", ctypes.c_int), ("by", ctypes.c_int),
                ("bx", ctypes.c_int), ("bh", ctypes.c_int), ("bw", ctypes.c_int), ("rh", ctypes.c_int),
                ("rw", ctypes.c_int), ("oy", ctypes.c_int), ("ox", ctypes.c_int), ("flip", ctypes.c_int)]


def random_resized_crop_params(height: int, width: int, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision `RandomResizedCrop.get_params`: same draws from torch's global generator, same arithmetic.
    -> (top, left, h, w)."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)  # fallback: central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


class GpuTransform:
    """One of the reference's transform stacks as a batched GPU call.

    mode: "train" (random-resized-crop [+ flip]), "test" (resize shorter edge + centre crop), "stretch" (resize to SIZE).
    __call__(images) takes a list of uint8 [H,W,3] arrays (numpy or CPU torch; what `np.asarray(PIL.Image)` gives) and
    returns a CUDA tensor [B,3,S_h,S_w] (`out_dtype`), produced on the current stream."""

    def __init__(self, size: Sequence[int] = (224, 224), mean: Sequence[float] = CLIP_MEAN,
                 std: Sequence[float] = CLIP_STD, mode: str = "train", scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                 flip_p: Optional[float] = None, normalize: bool = True, resize_edge: Optional[int] = None,
                 out_dtype=torch.float16, device="cuda"):
        if mode not in ("train", "test", "stretch"):
            raise ValueError(f"unknown mode {mode!r}")
        if out_dtype not in (torch.float16, torch.float32):
            raise ValueError("out_dtype must be torch.float16 or torch.float32")
        self.size = (int(size[0]), int(size[1]))
        if flip_p is None:  # RandomHorizontalFlip belongs to the training stacks only
            flip_p = 0.5 if mode == "train" else 0.0
        self.mode, self.scale, self.ratio, self.flip_p = mode, tuple(scale), tuple(ratio), float(flip_p)
        self.mean = (ctypes.c_float * 3)(*(mean if normalize else (0.0, 0.0, 0.0)))
        self.std = (ctypes.c_float * 3)(*(std if normalize else (1.0, 1.0, 1.0)))
        self.resize_edge = int(resize_edge) if resize_edge is not None else max(self.size)
        self.out_dtype = out_dtype
        self.device = torch.device(device)
        self._copied: Optional[torch.cuda.Event] = None
        self._stage: Optional[torch.Tensor] = None   # pinned host staging: descriptors + pixels
        self._dev: Optional[torch.Tensor] = None
        self._ws: Optional[torch.Tensor] = None
        self.last_descs: List[ImageDesc] = []

    # ---- geometry (host) ---------------------------------------------------------------------------------------------
    def geometry(self, H: int, W: int) -> Tuple[Tuple[int, int, int, int], Tuple[int, int], Tuple[int, int], int]:
        """-> (box (top,left,h,w), resized (rh,rw), window offset (oy,ox), flip) for one image; draws for mode 'train'."""
        sh, sw = self.size
        if self.mode == "train":
            box = random_resized_crop_params(H, W, self.scale, self.ratio)
            flip = int(bool(torch.rand(1) < self.flip_p)) if self.flip_p > 0 else 0
            return box, (sh, sw), (0, 0), flip
        if self.mode == "stretch":  # Resize(SIZE) [+ RandomHorizontalFlip when the caller asked for it]
            flip = int(bool(torch.rand(1) < self.flip_p)) if self.flip_p > 0 else 0
            return (0, 0, H, W), (sh, sw), (0, 0), flip
        s = self.resize_edge  # Resize(int): shorter edge -> s, the other int(s * long / short); then CenterCrop
        if W <= H:
            rw, rh = s, int(s * H / W)
        else:
            rh, rw = s, int(s * W / H)
        if rh < sh or rw < sw:
            raise _lib.MvlptError("CenterCrop larger than the resized image (torchvision would pad); not supported")
        return (0, 0, H, W), (rh, rw), (int(round((rh - sh) / 2.0)), int(round((rw - sw) / 2.0))), 0

    # ---- the batched call ------------------------------------------------------------------------------------------------
    def __call__(self, images: Sequence) -> torch.Tensor:
        if not torch.cuda.is_available():
            raise _lib.MvlptError("GpuTransform needs a CUDA device (there is no CPU path)")
        B = len(images)
        if B == 0:
            return torch.empty(0, 3, *self.size, device=self.device, dtype=self.out_dtype)
        arrs = []
        for im in images:
            a = im.numpy() if isinstance(im, torch.Tensor) else np.asarray(im)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise _lib.MvlptError(f"expected uint8 [H,W,3] images, got {a.dtype} {a.shape}")
            arrs.append(np.ascontiguousarray(a))
        desc_bytes = ctypes.sizeof(ImageDesc) * B
        head = (desc_bytes + 255) & ~255
        descs = (ImageDesc * B)()
        off = head
        boxes = []
        for b, a in enumerate(arrs):
            H, W = a.shape[:2]
            (by, bx, bh, bw), (rh, rw), (oy, ox), flip = self.geometry(H, W)
            # only the crop box crosses PCIe: it is staged as an image of its own (the filter never looks outside it)
            boxes.append(a[by:by + bh, bx:bx + bw])
            descs[b] = ImageDesc(off, bh, bw, 0, 0, bh, bw, rh, rw, oy, ox, flip)
            off += (bh * bw * 3 + 15) & ~15
        total = off
        if self._stage is None or self._stage.numel() < total:
            self._stage = torch.empty(int(total * 1.25), dtype=torch.uint8).pin_memory()
            self._dev = torch.empty(self._stage.numel(), dtype=torch.uint8, device=self.device)
        if self._copied is not None:
            self._copied.synchronize()  # the previous batch has left the pinned staging buffer
        st = self._stage.numpy()
        st[:desc_bytes] = np.frombuffer(descs, dtype=np.uint8)
        for b, a in enumerate(boxes):
            o = descs[b].src_off
            st[o:o + a.size].reshape(a.shape)[...] = a
        self._dev[:total].copy_(self._stage[:total], non_blocking=True)
        self._copied = torch.cuda.Event()
        self._copied.record(torch.cuda.current_stream(self.device))
        L = _lib.lib()
        ws = int(L.mvlpt_preprocess_workspace(descs, B, self.size[0], self.size[1]))
        if ws == 0:
            raise _lib.MvlptError("mvlpt_preprocess_workspace: " + (L.mvlpt_last_error() or b"failed").decode())
        if self._ws is None or self._ws.numel() < ws:
            self._ws = torch.empty(int(ws * 1.25), dtype=torch.uint8, device=self.device)
        out = torch.empty(B, 3, self.size[0], self.size[1], device=self.device, dtype=self.out_dtype)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(L.mvlpt_preprocess(self._dev.data_ptr(), descs, self._dev.data_ptr(), B, self.mean, self.std,
                                      out.data_ptr(), int(self.out_dtype == torch.float16), self.size[0], self.size[1],
                                      self._ws.data_ptr(), self._ws.numel(), stream), "mvlpt_preprocess")
        self.last_descs = list(descs)
        self.h2d_bytes = total
        return out


def _check_input_cfg(cfg):
    interp = getattr(cfg.INPUT, "INTERPOLATION", "bicubic")
    if interp != "bicubic":
        raise NotImplementedError(f"INPUT.INTERPOLATION={interp!r}: only 'bicubic' (every MVLPT config) is implemented")
    return tuple(cfg.INPUT.SIZE), tuple(cfg.INPUT.PIXEL_MEAN), tuple(cfg.INPUT.PIXEL_STD)


def build_transform(cfg, is_train: bool = True, out_dtype=torch.float16, device="cuda") -> GpuTransform:
    """Dassl `build_transform(cfg, is_train)` for the transform choices of the MVLPT configs."""
    size, mean, std = _check_input_cfg(cfg)
    choices = list(getattr(cfg.INPUT, "TRANSFORMS", ["random_resized_crop", "random_flip", "normalize"]))
    known = {"random_resized_crop", "random_flip", "normalize"}
    if set(choices) - known:
        raise NotImplementedError(f"INPUT.TRANSFORMS {sorted(set(choices) - known)} are not implemented")
    norm = "normalize" in choices
    if not is_train:
        return GpuTransform(size, mean, std, "test", normalize=norm, out_dtype=out_dtype, device=device)
    flip_p = 0.5 if "random_flip" in choices else 0.0
    if "random_resized_crop" not in choices:  # Dassl then resizes to SIZE first
        return GpuTransform(size, mean, std, "stretch", flip_p=flip_p, normalize=norm, out_dtype=out_dtype, device=device)
    scale = tuple(getattr(cfg.INPUT, "RRCROP_SCALE", (0.08, 1.0)))
    return GpuTransform(size, mean, std, "train", scale=scale, flip_p=flip_p, normalize=norm, out_dtype=out_dtype,
                        device=device)


def elevater_transform(cfg, out_dtype=torch.float16, device="cuda") -> GpuTransform:
    """`transform_clip` of trainers/vision_benchmark/evaluation/feature.py:540-553."""
    size, mean, std = _check_input_cfg(cfg)
    mode = "test" if getattr(cfg.DATASET, "CENTER_CROP", False) else "stretch"
    return GpuTransform(size, mean, std, mode, resize_edge=size[0], out_dtype=out_dtype, device=device)


def multilabel_to_vec(indices, n_classes: int) -> np.ndarray:
    """trainers/vision_benchmark/evaluation/feature.py:359-363: class indices -> multi-hot vector."""
    vec = np.zeros(n_classes, dtype=np.int64)
    for x in indices:
        vec[x] = 1
    return vec


class ElevaterBatches:
    """The loader the reference builds over `MultiTaskTorchDataset` — `get_dataloader` (feature.py:99-107, 851-857):
    batch 64, unshuffled, last batch kept — with the image transform moved off the CPU workers.

    `items`: any sequence of `(image, class_indices, idx_str, task_id)`, image = PIL image or uint8 [H,W,3] array, the
    tuple `MultiTaskTorchDataset.__getitem__` (feature.py:733-748) returns before its transform.  Yields the 4-tuples
    `MVLPT.parse_batch_train/test` index at 0, 1, 3 (trainers/mvlpt.py:957): `(image [B,3,S,S] on the device,
    target [B, num_classes] multi-hot, [idx_str], task [B] int64 on the CPU)`."""

    def __init__(self, items: Sequence, transform, num_classes: int, batch_size: int = 64):
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.items, self.transform, self.num_classes, self.batch_size = items, transform, int(num_classes), int(batch_size)

    def __len__(self) -> int:
        return (len(self.items) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for b0 in range(0, len(self.items), self.batch_size):
            chunk = [self.items[i] for i in range(b0, min(b0 + self.batch_size, len(self.items)))]
            arrays = []
            for im, *_ in chunk:
                if hasattr(im, "convert"):  # PIL image
                    im = np.asarray(im.convert("RGB"))
                arrays.append(im)
            target = torch.from_numpy(np.stack([multilabel_to_vec(c[1], self.num_classes) for c in chunk]))
            yield (self.transform(arrays), target, [c[2] for c in chunk],
                   torch.tensor([int(c[3]) for c in chunk], dtype=torch.int64))
