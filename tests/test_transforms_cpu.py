"""Input transforms, CPU side: the oracle against the golden vectors produced by the reference's own
transform classes (tests/golden/make_golden_transforms.py), the host logic of zs3_b200.dataloaders.gpu_transforms
(random draws in the reference's order), and the kernel SOURCE (csrc/aug.cu compiled for the host with
-DZS3_HOST_EMULATION) against the same vectors.  The GPU parity tests proper are tests/test_transforms_gpu.py."""
import ctypes as C
import hashlib
import os
import random
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zs3_transforms_oracle as TO  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "transforms.npz"))
MEAN, STD = tuple(GOLD["mean"]), tuple(GOLD["std"])


def synth(rs, h, w):
    """the picture generator of make_golden_transforms.py (the full-size golden cases store digests only)"""
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(xx / 9.0)[..., None] * np.cos(yy / 6.0)[..., None] * np.array([1.0, 0.7, -0.8])
    img = (base + rs.randint(-35, 35, size=(h, w, 3))).clip(0, 255).astype(np.uint8)
    lab = rs.randint(0, 21, size=(h // 6 + 1, w // 6 + 1)).astype(np.uint8).repeat(6, 0).repeat(6, 1)[:h, :w].copy()
    lab[rs.rand(h, w) < 0.03] = 255
    return img, lab


def test_oracle_matches_reference_train_transform():
    for i, (h, w, base, crop, seed) in enumerate(GOLD["train_cases"]):
        img, lab = GOLD[f"train{i}_img"], GOLD[f"train{i}_lab"]
        random.seed(int(seed))
        p = TO.draw_train_params(random, int(w), int(h), int(base), int(crop))
        x, y, _ = TO.train_transform(img, lab, p, int(crop), mean=MEAN, std=STD)
        assert np.array_equal(x, GOLD[f"train{i}_x"]) and x.dtype == np.float32, (i, p)
        assert np.array_equal(y, GOLD[f"train{i}_y"]) and y.dtype == np.float32, (i, p)


def test_oracle_matches_reference_val_transform_and_normalize():
    for i, (h, w, crop) in enumerate(GOLD["val_cases"]):
        x, y, _ = TO.val_transform(GOLD[f"val{i}_img"], GOLD[f"val{i}_lab"], int(crop), mean=MEAN, std=STD)
        assert np.array_equal(x, GOLD[f"val{i}_x"]) and np.array_equal(y, GOLD[f"val{i}_y"]), i
    assert np.array_equal(TO.normalize_lut(MEAN, STD), GOLD["normalize_ramp"])


def test_oracle_matches_reference_at_full_size():
    rs = np.random.RandomState(11)
    for (h, w, base, crop, seed), digest in zip(GOLD["big_cases"], GOLD["big_digests"]):
        img, lab = synth(rs, int(h), int(w))
        random.seed(int(seed))
        p = TO.draw_train_params(random, int(w), int(h), int(base), int(crop))
        x, y, _ = TO.train_transform(img, lab, p, int(crop), mean=MEAN, std=STD)
        assert hashlib.sha256(x.tobytes()).hexdigest() + hashlib.sha256(y.tobytes()).hexdigest() == str(digest)


def test_host_draws_follow_the_reference_order():
    from zs3_b200.dataloaders.gpu_transforms import GpuTransforms, normalize_table
    assert np.array_equal(normalize_table(MEAN, STD), GOLD["normalize_ramp"])
    for i, (h, w, base, crop, seed) in enumerate(GOLD["train_cases"]):
        t = GpuTransforms(base_size=int(base), crop_size=int(crop), mean=MEAN, std=STD, device="cpu")
        random.seed(int(seed))
        p = t.draw_train(int(w), int(h))
        q = dict(flip=bool(p["flip"]), ow=p["rw"], oh=p["rh"], x1=p["x1"], y1=p["y1"], radius=p["blur_radius"])
        x, y, _ = TO.train_transform(GOLD[f"train{i}_img"], GOLD[f"train{i}_lab"], q, int(crop), mean=MEAN, std=STD)
        assert np.array_equal(x, GOLD[f"train{i}_x"]) and np.array_equal(y, GOLD[f"train{i}_y"]), i
    for i, (h, w, crop) in enumerate(GOLD["val_cases"]):
        t = GpuTransforms(crop_size=int(crop), device="cpu")
        p = t.fix_scale(int(w), int(h))
        assert (p["rh"], p["rw"]) == GOLD[f"val{i}_y"].shape


# ------------------------------------------------------------------------------------------- kernel source on the host
@pytest.fixture(scope="module")
def emul():
    out_dir = os.path.join(HERE, "emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libaug_emul.so")
    src = os.path.join(ROOT, "zs3_b200", "csrc", "aug.cu")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-DZS3_HOST_EMULATION", "-I",
           os.path.join(HERE, "emul"), "-x", "c++", src, "-o", so, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from zs3_b200 import _lib as L
    lib = C.CDLL(so)
    lib.zs3_emul_augment_batch.restype = C.c_int
    lib.zs3_emul_augment_batch.argtypes = [C.POINTER(L.AugmentArgs), C.c_void_p]
    lib.zs3_emul_augment_workspace_size.restype = C.c_ulonglong
    lib.zs3_emul_augment_workspace_size.argtypes = [C.c_int] * 4
    return lib


def run_emulated(lib, arrays, params, out_w, out_h, fill=255):
    from zs3_b200 import _lib as L
    n = len(arrays)
    keep = [(np.ascontiguousarray(a), np.ascontiguousarray(b)) for a, b in arrays]
    items = (L.AugItem * n)()
    for i, ((img, lab), p) in enumerate(zip(keep, params)):
        items[i] = L.AugItem(img.ctypes.data, lab.ctypes.data, lab.shape[1], lab.shape[0], int(p["flip"]), p["ow"], p["oh"],
                             p["x1"], p["y1"], p["radius"])
    max_h = max(b.shape[0] for _, b in keep)
    ws = np.zeros(lib.zs3_emul_augment_workspace_size(n, max_h, out_w, out_h) + 16, dtype=np.uint8)
    off = (-ws.ctypes.data) % 16
    lut = TO.normalize_lut(MEAN, STD)
    x = np.full((n, 3, out_h, out_w), np.nan, dtype=np.float32)
    y = np.full((n, out_h, out_w), np.nan, dtype=np.float32)
    a = L.AugmentArgs(C.addressof(items), items, n, max_h, out_w, out_h, fill, lut.ctypes.data, x.ctypes.data,
                      y.ctypes.data, ws.ctypes.data + off, ws.size - off)
    assert lib.zs3_emul_augment_batch(C.byref(a), None) == 0
    return x, y


def test_kernel_source_matches_reference_on_the_host(emul):
    # three train cases as ONE batch (equal crop): different pictures sizes, flips, padding; then singles incl. blur
    cases = GOLD["train_cases"]
    groups = {}
    for i, c in enumerate(cases):
        groups.setdefault(int(c[3]), []).append(i)
    done_blur = done_plain = 0
    for crop, idx in groups.items():
        if crop > 65:
            continue                      # 256 host threads per block: keep the emulated grids small
        arrays, params = [], []
        for i in idx:
            h, w, base, _, seed = (int(v) for v in cases[i])
            random.seed(seed)
            params.append(TO.draw_train_params(random, w, h, base, crop))
            arrays.append((GOLD[f"train{i}_img"], GOLD[f"train{i}_lab"]))
        x, y = run_emulated(emul, arrays, params, crop, crop)
        for j, i in enumerate(idx):
            assert np.array_equal(x[j], GOLD[f"train{i}_x"]), (i, params[j])
            assert np.array_equal(y[j], GOLD[f"train{i}_y"]), (i, params[j])
            done_blur += params[j]["radius"] > 0
            done_plain += params[j]["radius"] <= 0
    assert done_blur >= 1 and done_plain >= 1
    # validation transform (FixScale): output size = resized size
    i = 2
    h, w, crop = (int(v) for v in GOLD["val_cases"][i])
    ow, oh = TO.fix_scale_sizes(w, h, crop)
    x, y = run_emulated(emul, [(GOLD[f"val{i}_img"], GOLD[f"val{i}_lab"])],
                        [dict(flip=False, ow=ow, oh=oh, x1=0, y1=0, radius=-1.0)], ow, oh)
    assert np.array_equal(x[0], GOLD[f"val{i}_x"]) and np.array_equal(y[0], GOLD[f"val{i}_y"])


def test_oracle_matches_pillow_directly_on_random_cases():
    """beyond the committed golden vectors: the numpy restatement against the installed Pillow itself (the reference's
    dependency) on seeded random sizes -- resize BILINEAR / NEAREST incl. > 2x down-scaling, GaussianBlur radii in [0, 3)"""
    PIL = pytest.importorskip("PIL")
    from PIL import Image, ImageFilter
    rs = np.random.RandomState(17)
    for _ in range(25):
        h, w = int(rs.randint(3, 80)), int(rs.randint(3, 80))
        oh, ow = int(rs.randint(2, 100)), int(rs.randint(2, 100))
        img, lab = synth(rs, h, w)
        assert np.array_equal(np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)), TO.resize_bilinear(img, ow, oh))
        assert np.array_equal(np.array(Image.fromarray(lab).resize((ow, oh), Image.NEAREST)), TO.resize_nearest(lab, ow, oh))
        r = float(rs.rand() * (3.0 if rs.rand() < 0.3 else 1.0))
        assert np.array_equal(np.array(Image.fromarray(img).filter(ImageFilter.GaussianBlur(radius=r))), TO.gaussian_blur(img, r))
    assert PIL.__version__
