"""Input pipeline, CPU side: the oracle's restatement of Pillow's bicubic resize + torchvision's
ToTensor/Normalize/crop-box draws is pinned BIT-EXACTLY to Pillow + torchvision themselves (both ship in this image) and
to the committed fixture tests/golden/preprocess.pt (oracle/gen_golden_preprocess.py); the host-side geometry of
mvlpt_b200.input_pipeline makes the same random draws as torchvision."""
import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as P
from tests.conftest import load_golden


@pytest.fixture(scope="module")
def golden():
    return load_golden("preprocess")


def _oracle_case(img, mode, size, mean, std):
    if mode == "stretch":
        u8 = P.resize_bicubic(img, size[0], size[1])
    else:
        u8 = P.resize_center_crop(img, size)
    return P.to_tensor_normalize(u8, mean, std)


def test_oracle_matches_golden_bit_exact(golden):
    imgs = [t.numpy() for t in golden["images"]]
    for case in golden["cases"]:
        out = _oracle_case(imgs[case["image"]], case["mode"], golden["size"], golden["mean"], golden["std"])
        assert np.array_equal(out, case["tensor"].numpy()), (case["image"], case["mode"])


def test_geometry_draws_and_oracle_reproduce_the_seeded_train_stack(golden):
    """RandomResizedCrop + RandomHorizontalFlip: same torch seed -> same boxes/flips -> bit-identical tensors."""
    from mvlpt_b200.input_pipeline import GpuTransform
    tf = GpuTransform(golden["size"], golden["mean"], golden["std"], "train")
    torch.manual_seed(golden["train_seed"])
    for img, ref in zip(golden["images"], golden["train"]):
        a = img.numpy()
        box, (rh, rw), (oy, ox), flip = tf.geometry(a.shape[0], a.shape[1])
        assert (rh, rw) == tuple(golden["size"]) and (oy, ox) == (0, 0)
        out = P.to_tensor_normalize(P.resized_crop(a, box, golden["size"], bool(flip)), golden["mean"], golden["std"])
        assert np.array_equal(out, ref.numpy())


def test_crop_params_equal_torchvision_draw_for_draw():
    import torchvision.transforms as T
    from mvlpt_b200.input_pipeline import random_resized_crop_params
    for seed, (h, w) in enumerate([(375, 500), (224, 224), (31, 900), (1200, 40), (8, 8)]):
        torch.manual_seed(seed)
        ref = [T.RandomResizedCrop.get_params(torch.empty(3, h, w), [0.08, 1.0], [3 / 4, 4 / 3]) for _ in range(20)]
        tail_ref = torch.rand(1)
        torch.manual_seed(seed)
        ours = [random_resized_crop_params(h, w) for _ in range(20)]
        assert ours == [tuple(r) for r in ref]
        assert torch.equal(torch.rand(1), tail_ref)  # consumed exactly the same number of draws


@pytest.mark.parametrize("shape,out", [((37, 53), (24, 24)), ((300, 200), (224, 224)), ((20, 30), (64, 48)),
                                       ((224, 224), (224, 224)), ((375, 500), (224, 224)), ((64, 64), (224, 224)),
                                       ((7, 9), (32, 32)), ((700, 31), (40, 40)), ((1, 1), (16, 16))])
def test_oracle_resize_equals_pillow_live(shape, out):
    from PIL import Image
    g = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = g.integers(0, 256, (*shape, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((out[1], out[0]), Image.BICUBIC))
    assert np.array_equal(P.resize_bicubic(img, out[0], out[1]), ref)


def test_coefficients_are_a_partition_of_unity():
    """Every row of fixed-point taps sums to 2^22 within the rounding of its entries; bounds stay inside the axis."""
    for in_size, out_size in [(500, 224), (224, 224), (100, 224), (3, 7), (1000, 3)]:
        bounds, kk = P.precompute_coeffs(in_size, out_size)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= in_size).all()
        assert np.abs(kk.sum(1) - (1 << P.PRECISION_BITS)).max() <= kk.shape[1]


def test_test_mode_geometry_equals_torchvision_resize_and_center_crop():
    """Resize(int) output size and CenterCrop offsets for odd aspect ratios, against torchvision on a PIL image."""
    import torchvision.transforms as T
    from PIL import Image
    from mvlpt_b200.input_pipeline import GpuTransform
    tf = GpuTransform((224, 224), mode="test")
    for (h, w) in [(375, 500), (500, 375), (224, 224), (225, 223), (1000, 333), (240, 1001), (224, 7000)]:
        box, (rh, rw), (oy, ox), flip = tf.geometry(h, w)
        pil = Image.new("RGB", (w, h))
        r = T.Resize(224, interpolation=T.InterpolationMode.BICUBIC)(pil)
        assert (r.size[1], r.size[0]) == (rh, rw), (h, w)
        # CenterCrop offsets: crop a coordinate ramp
        ramp = np.zeros((rh, rw, 3), np.uint8)
        ramp[..., 0] = (np.arange(rh) % 256)[:, None]
        ramp[..., 1] = (np.arange(rw) % 256)[None, :]
        c = np.asarray(T.CenterCrop((224, 224))(Image.fromarray(ramp)))
        assert c[0, 0, 0] == oy % 256 and c[0, 0, 1] == ox % 256, (h, w)
        assert box == (0, 0, h, w) and flip == 0


def test_cfg_mapping_follows_the_reference_transform_stacks():
    """build_transform (Dassl choices of configs/trainers/MVLPT/vit_b16.yaml) and elevater_transform (feature.py:540-553)."""
    from types import SimpleNamespace as NS
    from mvlpt_b200.input_pipeline import build_transform, elevater_transform, CLIP_MEAN, CLIP_STD
    inp = NS(SIZE=(224, 224), INTERPOLATION="bicubic", PIXEL_MEAN=list(CLIP_MEAN), PIXEL_STD=list(CLIP_STD),
             TRANSFORMS=["random_resized_crop", "random_flip", "normalize"])
    cfg = NS(INPUT=inp, DATASET=NS(CENTER_CROP=False))
    t = build_transform(cfg, True)
    assert (t.mode, t.flip_p, t.scale, t.size) == ("train", 0.5, (0.08, 1.0), (224, 224))
    assert abs(t.mean[0] - CLIP_MEAN[0]) < 1e-7 and abs(t.std[2] - CLIP_STD[2]) < 1e-7
    assert build_transform(cfg, False).mode == "test"
    assert elevater_transform(cfg).mode == "stretch"
    cfg.DATASET.CENTER_CROP = True
    e = elevater_transform(cfg)
    assert e.mode == "test" and e.resize_edge == 224
    inp.TRANSFORMS = ["normalize"]
    assert build_transform(cfg, True).mode == "stretch"
    inp.TRANSFORMS = ["random_flip"]
    nt = build_transform(cfg, True)
    assert nt.mode == "stretch" and nt.flip_p == 0.5 and nt.mean[0] == 0.0 and nt.std[0] == 1.0  # ToTensor only
    torch.manual_seed(0)
    flips = [nt.geometry(50, 60)[3] for _ in range(200)]
    assert 60 < sum(flips) < 140
    assert build_transform(cfg, False).flip_p == 0.0
    inp.TRANSFORMS = ["colorjitter"]
    with pytest.raises(NotImplementedError):
        build_transform(cfg, True)
    inp.TRANSFORMS, inp.INTERPOLATION = ["normalize"], "bilinear"
    with pytest.raises(NotImplementedError):
        build_transform(cfg, True)


def test_elevater_batches_follow_the_reference_loader():
    """get_dataloader (feature.py:99-107): unshuffled, batch 64, last batch kept; items collate to the 4-tuple the trainer
    indexes at 0, 1, 3; multi-hot targets (feature.py:359-363, 743-744).  The transform is stubbed (no GPU here)."""
    from PIL import Image
    from mvlpt_b200.input_pipeline import ElevaterBatches, multilabel_to_vec
    assert multilabel_to_vec([1, 3], 5).tolist() == [0, 1, 0, 1, 0]
    seen = []

    def fake_transform(arrays):
        seen.append([a.shape for a in arrays])
        assert all(a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3 for a in arrays)
        return torch.zeros(len(arrays), 3, 4, 4)

    items = []
    for i in range(150):
        img = Image.new("L" if i % 7 == 0 else "RGB", (10 + i % 5, 8 + i % 3)) if i % 2 else np.zeros((8, 9, 3), np.uint8)
        items.append((img, [i % 11] if i % 5 else [i % 11, (i + 3) % 11], f"id{i}", i % 4))
    loader = ElevaterBatches(items, fake_transform, num_classes=11)
    batches = list(loader)
    assert len(loader) == 3 and [b[0].shape[0] for b in batches] == [64, 64, 22]
    image, target, idx, task = batches[2]
    assert target.shape == (22, 11) and target.dtype == torch.int64 and idx[0] == "id128" and task.tolist()[:4] == [0, 1, 2, 3]
    assert target[2].tolist() == multilabel_to_vec([130 % 11, 133 % 11], 11).tolist()   # item 130: two labels
    assert seen[0][1] == (9, 11, 3)                                                    # PIL (w=11, h=9) -> [H, W, 3]
