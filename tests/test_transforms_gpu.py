"""Input transforms on the device: zs3_augment_batch through the C ABI (zs3_b200.dataloaders.gpu_transforms)
against (1) the golden vectors produced by the reference's transform classes, (2) the pinned oracle on seeded random
cases incl. the edge cases (no resize, pure padding, 1-pixel margins, tiny pictures, deep down-scaling, every blur
radius regime), (3) full-size 513x513 batches through the stored digests of the reference outputs.  Bit-exact."""
import hashlib
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zs3_transforms_oracle as TO  # noqa: E402
from test_transforms_cpu import GOLD, MEAN, STD, synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _tf(base, crop):
    from zs3_b200.dataloaders.gpu_transforms import GpuTransforms
    return GpuTransforms(base_size=base, crop_size=crop, mean=MEAN, std=STD)


def test_train_transform_equals_reference_golden():
    cases = GOLD["train_cases"]
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault((int(c[2]), int(c[3])), []).append(i)
    for (base, crop), idx in groups.items():
        t = _tf(base, crop)
        samples, params = [], []
        for i in idx:
            random.seed(int(cases[i][4]))
            h, w = int(cases[i][0]), int(cases[i][1])
            params.append(t.draw_train(w, h))
            samples.append((GOLD[f"train{i}_img"], GOLD[f"train{i}_lab"]))
        out = t.run(samples, params, crop, crop)
        x, y = out["image"].cpu().numpy(), out["label"].cpu().numpy()
        for j, i in enumerate(idx):
            assert np.array_equal(x[j], GOLD[f"train{i}_x"]), (i, params[j])
            assert np.array_equal(y[j], GOLD[f"train{i}_y"]), (i, params[j])


def test_public_api_with_seeded_random_equals_reference_golden():
    # transform_tr draws from the global `random` exactly as the reference's Compose does for ONE sample
    for i, (h, w, base, crop, seed) in enumerate(GOLD["train_cases"]):
        random.seed(int(seed))
        out = _tf(int(base), int(crop)).transform_tr([{"image": GOLD[f"train{i}_img"], "label": GOLD[f"train{i}_lab"]}])
        assert np.array_equal(out["image"][0].cpu().numpy(), GOLD[f"train{i}_x"])
        assert np.array_equal(out["label"][0].cpu().numpy(), GOLD[f"train{i}_y"])


def test_val_transform_equals_reference_golden():
    for i, (h, w, crop) in enumerate(GOLD["val_cases"]):
        out = _tf(513, int(crop)).transform_val([{"image": GOLD[f"val{i}_img"], "label": GOLD[f"val{i}_lab"]}])
        assert np.array_equal(out["image"][0].cpu().numpy(), GOLD[f"val{i}_x"])
        assert np.array_equal(out["label"][0].cpu().numpy(), GOLD[f"val{i}_y"])
    with pytest.raises(ValueError):
        _tf(513, 48).transform_val([{"image": GOLD["val0_img"], "label": GOLD["val0_lab"]},
                                    {"image": GOLD["val1_img"], "label": GOLD["val1_lab"]}])


def test_full_size_batch_equals_reference_digests():
    rs = np.random.RandomState(11)
    t = _tf(513, 513)
    samples, params = [], []
    for h, w, base, crop, seed in GOLD["big_cases"]:
        samples.append(synth(rs, int(h), int(w)))
        random.seed(int(seed))
        params.append(t.draw_train(int(w), int(h)))
    out = t.run(samples, params, 513, 513)
    x, y = out["image"].cpu().numpy(), out["label"].cpu().numpy()
    for j, digest in enumerate(GOLD["big_digests"]):
        assert hashlib.sha256(x[j].tobytes()).hexdigest() + hashlib.sha256(y[j].tobytes()).hexdigest() == str(digest), j


def _oracle_params(p):
    return dict(flip=bool(p["flip"]), ow=p["rw"], oh=p["rh"], x1=p["x1"], y1=p["y1"], radius=p["blur_radius"])


def test_random_and_edge_cases_equal_the_oracle():
    rs = np.random.RandomState(5)
    crop = 37
    t = _tf(37, crop)
    samples, params = [], []
    # hand-made edge cases: identity size, one axis resized, picture smaller than the crop in both axes (pure padding
    # on two sides), crop flush with the right / bottom edge, 1x1 picture, 8x down-scaling, blur radii 1e-4 .. 0.999, 2.5
    edge = [
        (crop, crop, dict(flip=0, rw=crop, rh=crop, x1=0, y1=0, blur_radius=-1.0)),
        (crop, 50, dict(flip=1, rw=50, rh=crop, x1=13, y1=0, blur_radius=0.5)),
        (20, 25, dict(flip=0, rw=25, rh=20, x1=0, y1=0, blur_radius=0.9)),
        (40, 60, dict(flip=1, rw=90, rh=60, x1=90 - crop, y1=60 - crop, blur_radius=-1.0)),
        (1, 1, dict(flip=1, rw=19, rh=19, x1=0, y1=0, blur_radius=0.3)),
        (296, 290, dict(flip=0, rw=37, rh=38, x1=0, y1=1, blur_radius=-1.0)),
        (64, 64, dict(flip=0, rw=41, rh=45, x1=2, y1=3, blur_radius=1e-4)),
        (64, 64, dict(flip=1, rw=41, rh=45, x1=2, y1=3, blur_radius=0.999)),
        (64, 64, dict(flip=0, rw=80, rh=77, x1=20, y1=3, blur_radius=2.5)),
        (64, 64, dict(flip=0, rw=80, rh=77, x1=20, y1=3, blur_radius=0.0)),
    ]
    for h, w, p in edge:
        samples.append(synth(rs, h, w))
        params.append(p)
    for k in range(22):
        h, w = int(rs.randint(8, 120)), int(rs.randint(8, 120))
        samples.append(synth(rs, h, w))
        random.seed(1000 + k)
        params.append(t.draw_train(w, h))
    out = t.run(samples, params, crop, crop)
    x, y = out["image"].cpu().numpy(), out["label"].cpu().numpy()
    for j, ((img, lab), p) in enumerate(zip(samples, params)):
        ox, oy, _ = TO.train_transform(img, lab, _oracle_params(p), crop, mean=MEAN, std=STD)
        assert np.array_equal(x[j], ox), (j, p)
        assert np.array_equal(y[j], oy), (j, p)


def test_pictures_only_and_argument_errors():
    from zs3_b200 import _lib as L
    rs = np.random.RandomState(3)
    t = _tf(33, 33)
    img, lab = synth(rs, 40, 50)
    p = dict(flip=0, rw=60, rh=48, x1=5, y1=6, blur_radius=0.4)
    a = t.run([(img, lab)], [p], 33, 33)
    b = t.run([(img, lab)], [p], 33, 33, want_label=False)
    assert b["label"] is None and torch.equal(a["image"], b["image"])
    with pytest.raises(L.Zs3NativeError):      # more than 8x down-scaling
        t.run([synth(rs, 400, 400)], [dict(flip=0, rw=40, rh=40, x1=0, y1=0, blur_radius=-1.0)], 33, 33)
    with pytest.raises(L.Zs3NativeError):
        t.run([(img, lab)], [dict(flip=0, rw=0, rh=48, x1=0, y1=0, blur_radius=-1.0)], 33, 33)
