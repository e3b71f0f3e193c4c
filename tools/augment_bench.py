"""Times zs3_augment_batch on a VOC-like batch (16 pictures around 500x375 -> 513x513 crops) and the same work through
Pillow on the host (the calls custom_transforms.py makes).  Usage: python tools/augment_bench.py [n] [repeats]"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pictures(n, seed=3):
    rs = np.random.RandomState(seed)
    sizes = [(375, 500), (500, 375), (333, 500), (500, 333), (281, 500), (375, 500), (364, 500), (500, 400)]
    out = []
    for i in range(n):
        h, w = sizes[i % len(sizes)]
        yy, xx = np.mgrid[0:h, 0:w]
        base = 127 + 100 * np.sin(xx / 19.0)[..., None] * np.cos(yy / 13.0)[..., None] * np.array([1.0, 0.7, -0.8])
        img = (base + rs.randint(-35, 35, size=(h, w, 3))).clip(0, 255).astype(np.uint8)
        lab = rs.randint(0, 21, size=(h // 16 + 1, w // 16 + 1)).astype(np.uint8).repeat(16, 0).repeat(16, 1)[:h, :w].copy()
        out.append((img, lab))
    return out


def pillow_batch(samples, params, crop, mean, std):
    """the Pillow calls of RandomHorizontalFlip / RandomScaleCrop / RandomGaussianBlur / Normalize / ToTensor with the
    given draws (custom_transforms.py:47-104, :8-44)"""
    from PIL import Image, ImageFilter, ImageOps
    xs, ys = [], []
    for (img, lab), p in zip(samples, params):
        im, mk = Image.fromarray(img), Image.fromarray(lab)
        if p["flip"]:
            im, mk = im.transpose(Image.FLIP_LEFT_RIGHT), mk.transpose(Image.FLIP_LEFT_RIGHT)
        im, mk = im.resize((p["rw"], p["rh"]), Image.BILINEAR), mk.resize((p["rw"], p["rh"]), Image.NEAREST)
        padw, padh = max(crop - p["rw"], 0), max(crop - p["rh"], 0)
        if padw or padh:
            im = ImageOps.expand(im, border=(0, 0, padw, padh), fill=0)
            mk = ImageOps.expand(mk, border=(0, 0, padw, padh), fill=255)
        box = (p["x1"], p["y1"], p["x1"] + crop, p["y1"] + crop)
        im, mk = im.crop(box), mk.crop(box)
        if p["blur_radius"] >= 0:
            im = im.filter(ImageFilter.GaussianBlur(radius=p["blur_radius"]))
        x = np.array(im).astype(np.float32)
        x /= 255.0
        x -= mean
        x /= std
        xs.append(torch.from_numpy(x.transpose((2, 0, 1))).float())
        ys.append(torch.from_numpy(np.array(mk).astype(np.float32)).float())
    return torch.stack(xs), torch.stack(ys)


def main():
    from zs3_b200.dataloaders.gpu_transforms import PASCAL_MEAN, PASCAL_STD, GpuTransforms
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    t = GpuTransforms(513, 513)
    samples = pictures(n)
    random.seed(1)
    params = [t.draw_train(lab.shape[1], lab.shape[0]) for _, lab in samples]
    out = t.run(samples, params, 513, 513)
    px, py = pillow_batch(samples, params, 513, PASCAL_MEAN, PASCAL_STD)
    assert torch.equal(out["image"].cpu(), px) and torch.equal(out["label"].cpu(), py), "GPU batch != Pillow batch"
    # e2e (host bytes -> device tensors, H2D inside) and device-only time of the five launches
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = t.run(samples, params, 513, 513)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for _ in range(3):
        pillow_batch(samples, params, 513, PASCAL_MEAN, PASCAL_STD)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    src_bytes = sum(a.size + b.size for a, b in samples)
    out_bytes = n * 513 * 513 * 4 * 4
    res = {"n": n, "blurred": int(sum(p["blur_radius"] >= 0 for p in params)), "e2e_ms_per_batch": e2e_ms,
           "e2e_images_per_sec": n / e2e_ms * 1e3, "h2d_bytes": int(t.h2d_bytes), "source_bytes": int(src_bytes),
           "output_bytes": int(out_bytes), "pillow_ms_per_batch_1_thread": cpu_ms,
           "pillow_images_per_sec_1_thread": n / cpu_ms * 1e3, "bit_exact_vs_pillow": True}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
