"""Batch-level input transforms on the device: the host side of csrc/aug.cu.

Mirrors zs3/dataloaders/custom_transforms.py as composed by the datasets (zs3/dataloaders/datasets/pascal.py:120-144,
context.py:168-193, sbd.py:137-158): `transform_tr` = RandomHorizontalFlip -> RandomScaleCrop(base_size, crop_size,
fill=255) -> RandomGaussianBlur -> Normalize(mean, std) -> ToTensor, `transform_val` = FixScale(crop_size) ->
Normalize -> ToTensor.  The reference runs them per sample on PIL images inside the DataLoader workers; here the
decoded pictures of a whole batch are uploaded as bytes, the random numbers are drawn on the host from Python's
`random` module IN THE REFERENCE'S ORDER (so the same `random.seed` gives the same batch, bit for bit), and the five
kernels of `zs3_augment_batch` produce the float tensors the trainer feeds to the model.
"""
import ctypes as C
import random as _random

import numpy as np
import torch

from .. import _lib as L

PASCAL_MEAN = (0.485, 0.456, 0.406)
PASCAL_STD = (0.229, 0.224, 0.225)


def normalize_table(mean, std):
    """Normalize.__call__ (custom_transforms.py:20-27) evaluated for the 256 byte values of each channel with numpy's
    own arithmetic (float32 /255, then -mean and /std through float64): [3, 256] float32."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32)[:, None].repeat(3, axis=1)
    v /= 255.0
    v -= mean
    v /= std
    return np.ascontiguousarray(v.T)


def _as_bytes(sample):
    """{"image": PIL image | [h, w, 3] uint8, "label": PIL image | [h, w] uint8} -> two contiguous uint8 arrays."""
    img = np.ascontiguousarray(np.asarray(sample["image"], dtype=np.uint8))
    lab = np.ascontiguousarray(np.asarray(sample["label"], dtype=np.uint8))
    if img.ndim != 3 or img.shape[2] != 3 or lab.shape != img.shape[:2]:
        raise ValueError(f"expected an RGB picture [h, w, 3] and its label map [h, w], got {img.shape} / {lab.shape}")
    return img, lab


class GpuTransforms:
    """transform_tr / transform_val for a list of samples; returns {"image": [n,3,H,W] float32, "label": [n,H,W]
    float32} on the device (the DataLoader's collated batch)."""

    def __init__(self, base_size=513, crop_size=513, fill=255, mean=PASCAL_MEAN, std=PASCAL_STD, device="cuda"):
        self.base_size, self.crop_size, self.fill = base_size, crop_size, fill
        self.device = torch.device(device)
        self.lut = torch.from_numpy(normalize_table(mean, std)).to(self.device)
        self._ws = None

    # ---- the random draws, in the order the reference's Compose makes them (one sample after the other)
    def draw_train(self, w, h, rng=_random):
        flip = rng.random() < 0.5                                                   # RandomHorizontalFlip :51
        short_size = rng.randint(int(self.base_size * 0.5), int(self.base_size * 2.0))   # RandomScaleCrop :80
        if h > w:                                                                   # :82-88
            ow = short_size
            oh = int(1.0 * h * ow / w)
        else:
            oh = short_size
            ow = int(1.0 * w * oh / h)
        pw = max(ow, self.crop_size)                                                # :92-96 pad right / bottom
        ph = max(oh, self.crop_size)
        x1 = rng.randint(0, pw - self.crop_size)                                    # :98-99
        y1 = rng.randint(0, ph - self.crop_size)
        radius = -1.0
        if rng.random() < 0.5:                                                      # RandomGaussianBlur :62-63
            radius = rng.random()
        return dict(flip=int(flip), rw=ow, rh=oh, x1=x1, y1=y1, blur_radius=radius)

    def fix_scale(self, w, h):
        if w > h:                                                                   # FixScale :115-120
            oh = self.crop_size
            ow = int(1.0 * w * oh / h)
        else:
            ow = self.crop_size
            oh = int(1.0 * h * ow / w)
        return dict(flip=0, rw=ow, rh=oh, x1=0, y1=0, blur_radius=-1.0)

    # ---- device side
    def stage(self, arrays, params):
        """pack the pictures, label maps and the item table into one pinned buffer and start its (single) H2D copy;
        returns the staged batch for launch()"""
        n = len(arrays)
        sizes = [a[0].size + a[1].size for a in arrays]
        offs = np.concatenate([[0], np.cumsum([(s + 15) // 16 * 16 for s in sizes])]).astype(np.int64)
        staging = torch.empty(int(offs[-1]) + n * C.sizeof(L.AugItem) + 16, dtype=torch.uint8).pin_memory()
        host = staging.numpy()
        items = (L.AugItem * n)()
        dev = torch.empty_like(staging, device=self.device)
        base = dev.data_ptr()
        for i, ((img, lab), p) in enumerate(zip(arrays, params)):
            o = int(offs[i])
            host[o:o + img.size] = img.reshape(-1)
            host[o + img.size:o + img.size + lab.size] = lab.reshape(-1)
            h, w = lab.shape
            items[i] = L.AugItem(base + o, base + o + img.size, w, h, p["flip"], p["rw"], p["rh"], p["x1"], p["y1"],
                                 p["blur_radius"])
        item_off = int(offs[-1])
        host[item_off:item_off + C.sizeof(items)] = np.frombuffer(items, dtype=np.uint8)
        dev.copy_(staging, non_blocking=True)
        self.h2d_bytes = staging.numel()
        return dict(dev=dev, items=items, item_ptr=base + item_off, n=n, max_h=max(a[1].shape[0] for a in arrays),
                    staging=staging)

    def launch(self, staged, out_w, out_h, want_label=True):
        """the five launches of zs3_augment_batch on the current stream"""
        lib = L.lib()
        n = staged["n"]
        need = lib.zs3_augment_workspace_size(n, staged["max_h"], out_w, out_h)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        image = torch.empty(n, 3, out_h, out_w, dtype=torch.float32, device=self.device)
        label = torch.empty(n, out_h, out_w, dtype=torch.float32, device=self.device) if want_label else None
        a = L.AugmentArgs(staged["item_ptr"], staged["items"], n, staged["max_h"], out_w, out_h, self.fill,
                          self.lut.data_ptr(), image.data_ptr(), label.data_ptr() if want_label else None,
                          self._ws.data_ptr(), self._ws.numel())
        L.check(lib.zs3_augment_batch(C.byref(a), L.stream_ptr()), "zs3_augment_batch")
        staged["dev"].record_stream(torch.cuda.current_stream())
        return {"image": image, "label": label}

    def run(self, arrays, params, out_w, out_h, want_label=True):
        """arrays: list of (image uint8 [h,w,3], label uint8 [h,w]); params: list of dicts (draw_train / fix_scale)."""
        return self.launch(self.stage(arrays, params), out_w, out_h, want_label)

    def transform_tr(self, samples, rng=_random):
        arrays = [_as_bytes(s) for s in samples]
        params = [self.draw_train(lab.shape[1], lab.shape[0], rng) for _, lab in arrays]
        return self.run(arrays, params, self.crop_size, self.crop_size)

    def transform_val(self, samples):
        """FixScale keeps the aspect ratio, so the samples of one call must resize to the same size (the reference
        validates with pictures of one size per batch or batch size 1)."""
        arrays = [_as_bytes(s) for s in samples]
        params = [self.fix_scale(lab.shape[1], lab.shape[0]) for _, lab in arrays]
        sizes = {(p["rw"], p["rh"]) for p in params}
        if len(sizes) != 1:
            raise ValueError(f"transform_val: the samples resize to different sizes {sorted(sizes)}; batch them by size")
        (ow, oh), = sizes
        return self.run(arrays, params, ow, oh)
