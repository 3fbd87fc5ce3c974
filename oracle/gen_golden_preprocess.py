"""Golden vectors for the input pipeline: tests/golden/preprocess.pt.

Produced by the transform stacks the reference itself builds — torchvision over Pillow, with the arguments of
trainers/vision_benchmark/evaluation/feature.py:540-553 and of Dassl's build_transform for
configs/trainers/MVLPT/vit_b16.yaml:8-13 — on small synthetic images.  Run here (torchvision + Pillow are in the image):
    python oracle/gen_golden_preprocess.py
"""
from pathlib import Path

import numpy as np
import torch
import torchvision.transforms as T
import torchvision.transforms.functional as F
from PIL import Image

MEAN = [0.48145466, 0.4578275, 0.40821073]
STD = [0.26862954, 0.26130258, 0.27577711]
S = (48, 48)


def synth_image(h, w, seed):
    """Smooth structure + noise, uint8 [h,w,3]: exercises clipping (over/undershoot of the bicubic lobes) and rounding."""
    g = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 127 * np.sin(x / (3.0 + c) + y / 7.0) * np.cos(y / (5.0 + c)) for c in range(3)], -1)
    img = base + g.normal(0, 25, (h, w, 3))
    img[h // 3: h // 3 + 3] = 255  # hard edges
    img[:, w // 2: w // 2 + 2] = 0
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    out = {"size": S, "mean": MEAN, "std": STD, "images": [], "cases": []}
    shapes = [(37, 53), (120, 90), (48, 48), (20, 31), (200, 333), (64, 17)]
    imgs = [synth_image(h, w, i) for i, (h, w) in enumerate(shapes)]
    out["images"] = [torch.from_numpy(a) for a in imgs]
    bic = T.InterpolationMode.BICUBIC
    norm = T.Compose([T.ToTensor(), T.Normalize(MEAN, STD)])
    stretch = T.Compose([T.Resize(S, interpolation=bic), norm])                      # feature.py:548-553
    center = T.Compose([T.Resize(S[0], interpolation=bic), T.CenterCrop(S), norm])    # feature.py:541-546 / Dassl test
    for i, a in enumerate(imgs):
        pil = Image.fromarray(a)
        out["cases"].append(dict(image=i, mode="stretch", tensor=stretch(pil)))
        if min(a.shape[:2]) >= 1:
            out["cases"].append(dict(image=i, mode="test", tensor=center(pil)))
    # Dassl train stack, seeded: RandomResizedCrop(S, scale=(0.08, 1), bicubic) -> RandomHorizontalFlip -> ToTensor -> Normalize
    torch.manual_seed(4321)
    train = T.Compose([T.RandomResizedCrop(S, scale=(0.08, 1.0), interpolation=bic), T.RandomHorizontalFlip(), norm])
    out["train_seed"] = 4321
    out["train"] = [train(Image.fromarray(a)) for a in imgs]
    p = Path(__file__).resolve().parent.parent / "tests" / "golden" / "preprocess.pt"
    torch.save(out, p)
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
