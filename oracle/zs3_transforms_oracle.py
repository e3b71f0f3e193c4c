"""CPU oracle of the training / validation input transforms -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy integer / IEEE arithmetic, what zs3/dataloaders/custom_transforms.py computes through its
third-party dependency Pillow (absent from the reference tree; no version pin in the reference's setup.py:1-4 or
Dockerfile -- pinned HERE against the Pillow 12.2.0 of this image by tests/golden/make_golden_transforms.py, which runs
the reference's own classes):

  * RandomHorizontalFlip (custom_transforms.py:47-56)  -> Image.transpose(FLIP_LEFT_RIGHT): mirrored columns;
  * RandomScaleCrop (:69-104)  -> Image.resize BILINEAR (image) / NEAREST (label), ImageOps.expand on the right/bottom,
    Image.crop;
  * FixScale (:107-124)        -> the same two resizes;
  * RandomGaussianBlur (:58-66) -> ImageFilter.GaussianBlur(radius): three box-blur passes per direction on 8-bit data;
  * Normalize (:8-27) and ToTensor (:30-44): /255 in float32, -mean and /std evaluated in float64 and rounded to float32
    (numpy in-place semantics with a tuple operand), HWC -> CHW.

Pillow algorithms restated (published source, src/libImaging):
  Resample.c   precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / ...Vertical_8bpc:
               antialiased triangle filter, support = max(scale, 1), coefficients rounded to 22 fractional bits,
               horizontal pass then vertical pass, each rounding to 8 bits ((sum + 2^21) >> 22, saturated);
  Geometry.c   ImagingScaleAffine (nearest): source index = int(xo) with xo ACCUMULATED in double
               (xo = a*0.5; xo += a per output pixel), same for rows;
  BoxBlur.c    _gaussian_blur_radius (float32 arithmetic with double sqrt/floor), ImagingLineBoxBlur: window sum with
               replicated edges, ww = uint32(2^24 / (2r+1)) (float32 division), fw = (2^24 - (2*int(r)+1)*ww) / 2,
               out = (acc*ww + (left+right)*fw + 2^23) >> 24 in uint32.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


# ------------------------------------------------------------------------------------------------ resize, bilinear
def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the whole axis.
    -> (xmin [out], count [out], kk [out, ksize] int32)"""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size       # (double)(in1 - in0) / outSize, in0/in1 floats
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    cnt = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.empty(n, dtype=np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            t = -t if t < 0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx] = lo
        cnt[xx] = n
    return xmin, cnt, kk


def _resample_axis0(img, out_size):
    """One 8-bit pass along axis 0 of img [n, ...] uint8."""
    xmin, cnt, kk = bilinear_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt[xx]):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear(img, ow, oh):
    """Image.resize((ow, oh), BILINEAR) of an 8-bit image [h, w, c] (Resample.c ImagingResample: horizontal pass, then
    vertical pass; a pass is skipped when that axis keeps its size)."""
    h, w = img.shape[:2]
    out = img
    if ow != w:
        out = np.swapaxes(_resample_axis0(np.swapaxes(out, 0, 1), ow), 0, 1)
    if oh != h:
        out = _resample_axis0(out, oh)
    return np.ascontiguousarray(out)


# ------------------------------------------------------------------------------------------------- resize, nearest
def nearest_index(in_size, out_size):
    """Geometry.c ImagingScaleAffine: accumulated double coordinate, COORD(v) = v < 0 ? -1 : (int)v; -1 = not copied."""
    a = float(in_size) / out_size
    idx = np.empty(out_size, dtype=np.int32)
    xo = 0.0 + a * 0.5
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        idx[x] = xin if 0 <= xin < in_size else -1
        xo += a
    return idx


def resize_nearest(img, ow, oh):
    """Image.resize((ow, oh), NEAREST) of [h, w] (or [h, w, c]); pixels without a source stay 0 (fill=1 memset)."""
    h, w = img.shape[:2]
    if (ow, oh) == (w, h):
        return img.copy()
    xi = nearest_index(w, ow)
    yi = nearest_index(h, oh)
    out = np.zeros((oh, ow) + img.shape[2:], dtype=img.dtype)
    vy = np.nonzero(yi >= 0)[0]
    vx = np.nonzero(xi >= 0)[0]
    out[np.ix_(vy, vx)] = img[np.ix_(yi[vy], xi[vx])]
    return out


# ------------------------------------------------------------------------------------------------------------ blur
def gaussian_box_radius(radius, passes=3):
    """BoxBlur.c _gaussian_blur_radius: float32 variables, the sqrt / floor operands promoted to double."""
    f = np.float32
    radius = f(radius)
    sigma2 = f(f(radius * radius) / f(passes))
    L = f(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f(math.floor((float(L) - 1.0) / 2.0))
    a = f(f(f(2) * l + f(1)) * f(f(l * f(l + f(1))) - f(f(3) * sigma2)))
    a = f(a / f(f(6) * f(sigma2 - f(f(l + f(1)) * f(l + f(1))))))
    return f(l + a)


def box_blur_weights(float_radius):
    float_radius = np.float32(float_radius)
    r = int(float_radius)
    ww = int(np.float32(1 << 24) / np.float32(float_radius * np.float32(2) + np.float32(1)))   # float division, truncated
    fw = (((1 << 24) - (r * 2 + 1) * ww) & 0xFFFFFFFF) // 2
    return r, ww, fw


def box_blur_axis1(img, float_radius):
    """One ImagingHorizontalBoxBlur pass over axis 1 of [h, w, c] uint8."""
    r, ww, fw = box_blur_weights(float_radius)
    h, w = img.shape[:2]
    src = img.astype(np.uint64)
    x = np.arange(w)
    acc = np.zeros(img.shape, dtype=np.uint64)
    for j in range(-r, r + 1):
        acc += src[:, np.clip(x + j, 0, w - 1)]
    far = src[:, np.clip(x - r - 1, 0, w - 1)] + src[:, np.clip(x + r + 1, 0, w - 1)]
    bulk = (acc * np.uint64(ww) + far * np.uint64(fw)) & np.uint64(0xFFFFFFFF)
    return (((bulk + np.uint64(1 << 23)) & np.uint64(0xFFFFFFFF)) >> np.uint64(24)).astype(np.uint8)


def gaussian_blur(img, radius, passes=3):
    """ImageFilter.GaussianBlur(radius).filter(img) for an 8-bit [h, w, c] image (BoxBlur.c ImagingGaussianBlur ->
    ImagingBoxBlur: `passes` horizontal passes, then `passes` vertical ones through a transpose)."""
    if radius == 0:
        return img.copy()
    br = gaussian_box_radius(radius, passes)
    out = img
    if br != 0:
        for _ in range(passes):
            out = box_blur_axis1(out, br)
        out = np.swapaxes(out, 0, 1)
        for _ in range(passes):
            out = box_blur_axis1(out, br)
        out = np.swapaxes(out, 0, 1)
    return np.ascontiguousarray(out)


# ------------------------------------------------------------------------------------------------------- normalise
def normalize_lut(mean, std):
    """Normalize.__call__ (custom_transforms.py:20-27) for each of the 256 byte values: [3, 256] float32."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32)[:, None].repeat(3, axis=1)     # [256, 3], like an HWC image
    v /= 255.0
    v -= mean
    v /= std
    return np.ascontiguousarray(v.T)


def normalize_to_tensor(img, label, mean, std):
    """Normalize + ToTensor: image [h, w, 3] uint8 -> [3, h, w] float32, label [h, w] uint8 -> float32."""
    x = np.array(img).astype(np.float32)
    x /= 255.0
    x -= mean
    x /= std
    return np.ascontiguousarray(x.transpose(2, 0, 1)), np.array(label).astype(np.float32)


# ------------------------------------------------------------------------------------------------------ pipelines
def scale_crop_sizes(w, h, short_size):
    """RandomScaleCrop :83-88 -- the resized (ow, oh)."""
    if h > w:
        ow = short_size
        oh = int(1.0 * h * ow / w)
    else:
        oh = short_size
        ow = int(1.0 * w * oh / h)
    return ow, oh


def fix_scale_sizes(w, h, crop_size):
    """FixScale :115-120."""
    if w > h:
        oh = crop_size
        ow = int(1.0 * w * oh / h)
    else:
        ow = crop_size
        oh = int(1.0 * h * ow / w)
    return ow, oh


def draw_train_params(rng, w, h, base_size, crop_size):
    """The random draws of transform_tr (datasets/pascal.py:120-134) in the reference's order, from a `random`-module
    compatible generator: flip (:51), short_size (:80), x1, y1 (:98-99), blur (:62) and its radius."""
    flip = rng.random() < 0.5
    short_size = rng.randint(int(base_size * 0.5), int(base_size * 2.0))
    ow, oh = scale_crop_sizes(w, h, short_size)
    pw = max(ow, crop_size) if short_size < crop_size else ow
    ph = max(oh, crop_size) if short_size < crop_size else oh
    x1 = rng.randint(0, pw - crop_size)
    y1 = rng.randint(0, ph - crop_size)
    radius = -1.0
    if rng.random() < 0.5:
        radius = rng.random()
    return dict(flip=flip, ow=ow, oh=oh, x1=x1, y1=y1, radius=radius)


def train_transform(img, label, p, crop_size, fill=255, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """transform_tr with the draws `p` (draw_train_params): image [h, w, 3] uint8, label [h, w] uint8 ->
    ([3, crop, crop] float32, [crop, crop] float32, the 8-bit crop before Normalize)."""
    if p["flip"]:
        img, label = img[:, ::-1], label[:, ::-1]
    ow, oh = p["ow"], p["oh"]
    h, w = label.shape
    im = resize_bilinear(np.ascontiguousarray(img), ow, oh) if (ow, oh) != (w, h) else np.ascontiguousarray(img)
    lb = resize_nearest(np.ascontiguousarray(label), ow, oh)
    pw, ph = max(ow, crop_size), max(oh, crop_size)
    pim = np.zeros((ph, pw, 3), dtype=np.uint8)
    plb = np.full((ph, pw), fill, dtype=np.uint8)
    pim[:oh, :ow] = im
    plb[:oh, :ow] = lb
    x1, y1 = p["x1"], p["y1"]
    cim = pim[y1:y1 + crop_size, x1:x1 + crop_size]
    clb = plb[y1:y1 + crop_size, x1:x1 + crop_size]
    if p["radius"] >= 0:
        cim = gaussian_blur(cim, p["radius"])
    x, y = normalize_to_tensor(cim, clb, mean, std)
    return x, y, np.ascontiguousarray(cim)


def val_transform(img, label, crop_size, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """transform_val (datasets/pascal.py:136-144): FixScale + Normalize + ToTensor."""
    h, w = label.shape
    ow, oh = fix_scale_sizes(w, h, crop_size)
    im = resize_bilinear(img, ow, oh) if (ow, oh) != (w, h) else img
    lb = resize_nearest(label, ow, oh)
    x, y = normalize_to_tensor(im, lb, mean, std)
    return x, y, im
