"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's host-side image preprocessing.

The reference turns OpenFace's 112x112 RGB face crops into network inputs with PIL / torchvision:

  gray stack : Image.open(bmp).convert('L') -> Resize(48, LANCZOS) -> np.stack -> float / 255
               (api/sampler/snippet_sampler.py:156-185, api/utils/data_utils.py:71-120)
  RGB frame  : Resize(256) [PIL bilinear] -> CenterCrop(224) -> ToTensor -> x * 255 -> Normalize(mean, std=1)
               (api/utils/model_utils.py:26-40, api/sampler/image_sampler.py:118-119)

The arithmetic lives in a third-party dependency that is not vendored in /root/reference: Pillow
(unpinned in api/readme.md; 12.2 in this image).  Its published algorithm is restated here from
libImaging/Convert.c (rgb2l: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16) and
libImaging/Resample.c (precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc /
ImagingResampleVertical_8bpc: separable filter, double-precision normalised taps converted to
22-bit fixed point, horizontal pass then vertical pass, each pass rounded and clipped to uint8).
Parity is PINNED: tests/test_preproc_oracle.py checks this restatement bit for bit against Pillow
and torchvision themselves (both are in the image), and tests/golden/preproc_*.npz holds vectors
produced by the real PIL/torchvision pipeline (oracle/make_golden_preproc.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
RESNET_MEAN = (131.0912, 103.8827, 91.4953)


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


FILTERS = {"bilinear": (_bilinear, 1.0), "lanczos": (_lanczos, 3.0)}


def precompute_coeffs(in_size, out_size, filter_name):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    Returns (ksize, bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize])."""
    fn, fsupport = FILTERS[filter_name]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
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
        w = [fn((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            # C casts truncate toward zero
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis_last(img, bounds, kk):
    """One 8bpc pass along the last axis: out[..., xx] = clip8((2^21 + sum_k img[..., xmin+k] * kk[xx,k]) >> 22)."""
    out_size = bounds.shape[0]
    out = np.empty(img.shape[:-1] + (out_size,), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (src[..., xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(axis=-1) + (1 << (PRECISION_BITS - 1))
        # the C code accumulates in int32; the taps sum to ~2^22 and pixels are < 2^8, so no wrap can occur
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_u8(img, out_h, out_w, filter_name):
    """PIL Image.resize of uint8 data laid out (..., H, W): horizontal pass, then vertical pass."""
    h, w = img.shape[-2], img.shape[-1]
    if w != out_w:
        _, b, k = precompute_coeffs(w, out_w, filter_name)
        img = _resample_axis_last(img, b, k)
    if h != out_h:
        _, b, k = precompute_coeffs(h, out_h, filter_name)
        img = np.swapaxes(_resample_axis_last(np.swapaxes(img, -1, -2), b, k), -1, -2)
    return np.ascontiguousarray(img)


def rgb_to_l(crops):
    """Convert.c rgb2l on (..., H, W, 3) uint8."""
    c = crops.astype(np.uint32)
    return ((c[..., 0] * 19595 + c[..., 1] * 38470 + c[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def crops_to_gray(crops, phase_size=48):
    """(n, H, W, 3) uint8 face crops -> (n, phase_size, phase_size) float32 in [0, 1]
    (convert('L') -> LANCZOS resize -> float / 255)."""
    g = resize_u8(rgb_to_l(np.asarray(crops)), phase_size, phase_size, "lanczos")
    return g.astype(np.float32) / np.float32(255)


def crops_to_rgb(crops, resize=256, crop=224, mean=RESNET_MEAN):
    """(n, H, H, 3) uint8 square face crops -> (n, 3, crop, crop) float32 = ((u8 / 255) * 255 - mean),
    the exact fp32 operation order of ToTensor -> x * 255.0 -> Normalize(mean, [1,1,1])."""
    x = np.moveaxis(np.asarray(crops), -1, -3)                    # (n, 3, H, W)
    x = resize_u8(x, resize, resize, "bilinear")
    off = int(round((resize - crop) / 2.0))                       # torchvision center_crop
    x = x[..., off:off + crop, off:off + crop]
    f = x.astype(np.float32) / np.float32(255)
    f = f * np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 3, 1, 1)
    return ((f - m) / np.float32(1.0)).astype(np.float32)
