"""Pillow's resampling tap tables, built on the host (data independent, uploaded once).

Pillow resizes 8-bit images with a separable filter whose taps are computed in double precision,
normalised per output index and converted to 22-bit fixed point (libImaging/Resample.c:
precompute_coeffs, normalize_coeffs_8bpc).  The CUDA resampler (csrc/preproc.cu) consumes exactly
these integers, which is what makes it bit-exact with the reference's PIL preprocessing
(api/utils/data_utils.py:71-84 LANCZOS, api/utils/model_utils.py:31 bilinear).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _tap(name, x):
    if name == 'bilinear':
        x = -x if x < 0.0 else x
        return 1.0 - x if x < 1.0 else 0.0
    if name == 'lanczos':
        if not (-3.0 <= x < 3.0):
            return 0.0

        def sinc(v):
            if v == 0.0:
                return 1.0
            v *= math.pi
            return math.sin(v) / v
        return sinc(x) * sinc(x / 3)
    raise ValueError('unknown filter %r' % name)


_SUPPORT = {'bilinear': 1.0, 'lanczos': 3.0}


def resample_table(in_size, out_size, filter_name):
    """(taps, bounds int32[out,2] = (first input index, tap count), kk int32[out,taps])."""
    scale = float(in_size) / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = _SUPPORT[filter_name] * filterscale
    taps = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, taps), np.int32)
    one = float(1 << PRECISION_BITS)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        ws = [_tap(filter_name, (i + lo - center + 0.5) * inv) for i in range(hi - lo)]
        total = 0.0
        for w in ws:
            total += w
        for i, w in enumerate(ws):
            if total != 0.0:
                w = w / total
            kk[o, i] = int(w * one - 0.5) if w < 0 else int(w * one + 0.5)      # C int casts truncate toward zero
        bounds[o] = (lo, hi - lo)
    return taps, bounds, kk
