"""Host-side (float64) tables for the B200 pyramid kernels.

The reference rebuilds its raised-cosine / angular masks with ``np.interp`` on every call
(api/steerable/SCFpyr_PyTorch.py:94-107,139-158,193-196) and then runs full complex FFTs of
the mirror-extended image.  Here the same masks are built ONCE per (size, height, nbands),
with the same ``np.interp`` calls, and folded into the form the CUDA kernels consume:

* the image is mirror-extended to S = 2H before the FFT
  (api/utils/phase_utils.py:116-129), so its DFT is  F[k,l] = w_k w_l C[|k|,|l|]  with
  w_k = exp(i pi k / S) and C the (real) 2-D DCT-II of the un-extended image;
* every oriented band is  ifft2(ifftshift(F . M))  with a real mask M times (-i)^(nb-1)
  (SCFpyr_PyTorch.py:161-171) and only its top-left quadrant is kept
  (api/phase_difference_extractor.py:84-85).  Folding the +-k, +-l terms gives, per output
  channel ch in {re, im},

      out_ch[y,x] = sum_k cos(a_k(y)) sum_l C[k,l] MA_ch[k,l] TA_ch(b_l(x))
                  + sum_k sin(a_k(y)) sum_l C[k,l] MB_ch[k,l] TB_ch(b_l(x)),
      a_k(y) = pi k / S + 2 pi k y / s,   b_l(x) = pi l / S + 2 pi l x / s,   TA/TB in {cos, sin}

  i.e. four small real matrix products per band; fftshift/ifftshift, the centre crops
  (SCFpyr_PyTorch.py:179-190), the 1/s^2 of the inverse FFT and the (-i)^(nb-1) twist all
  disappear into the real tables MA/MB.

Everything here is data independent; `PyramidTables` is uploaded once by the C ABI
(`mimamo_pyr_plan_create`, include/mimamo_b200.h).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np


def _pad4(n: int) -> int:
    """Leading dimensions are padded to multiples of 8 (the kernel's 4x8 / 8x4 register tiles need no edge handling)."""
    return (n + 7) // 8 * 8


# ---- the reference's mask recipe (host, float64) -------------------------------------------


def _polar_grid(m: int, n: int):
    # api/steerable/math_utils.py:52-60
    ax_m = np.linspace(-(m // 2) / (m / 2), (m // 2) / (m / 2) - (1 - m % 2) * 2 / m, num=m)
    ax_n = np.linspace(-(n // 2) / (n / 2), (n // 2) / (n / 2) - (1 - n % 2) * 2 / n, num=n)
    xv, yv = np.meshgrid(ax_n, ax_m)
    rad = np.hypot(xv, yv)
    rad[m // 2][n // 2] = rad[m // 2][n // 2 - 1]
    return np.log2(rad), np.arctan2(yv, xv)


def _rcos():
    # api/steerable/math_utils.py:62-69 with (width, position) = (1, -0.5)
    x = np.pi * np.arange(-257, 2) / 512
    y = np.cos(x) ** 2
    y[0], y[258] = y[1], y[257]
    return -0.5 + 2 / np.pi * (x + np.pi / 4), np.sqrt(y)


def _interp(grid, ys, xs):
    return np.interp(grid.ravel(), xs, ys).reshape(grid.shape)


def _crop(dim: int):
    # api/steerable/SCFpyr_PyTorch.py:182-183
    lo = int(np.ceil((dim + 0.5) / 2) - np.ceil((np.ceil((dim - 0.5) / 2) + 0.5) / 2))
    return lo, lo + int(np.ceil((dim - 0.5) / 2))


def max_height(size: int) -> int:
    # api/steerable/SCFpyr_PyTorch.py:90
    return int(np.floor(np.log2(size)) - 2)


def oriented_band_masks(size: int, height: int, nbands: int) -> List[List[np.ndarray]]:
    """Cumulative real mask (fftshifted order) of every oriented band: result[level-1][band],
    each of shape (s_level, s_level) on the level's own (cropped) frequency grid."""
    if nbands < 2:
        raise RecursionError("nbands must be >= 2 (the reference's factorial(0) never terminates, "
                             "api/steerable/math_utils.py:79-84)")
    log_rad, angle = _polar_grid(size, size)
    xr, yr = _rcos()
    low = _interp(log_rad, np.sqrt(1 - yr ** 2), xr)              # lo0mask
    lutsize = 1024
    xcosn = np.pi * np.arange(-(2 * lutsize + 1), lutsize + 2) / lutsize
    alpha = (xcosn + np.pi) % (2 * np.pi) - np.pi
    order = nbands - 1
    const = (2 ** (2 * order)) * math.factorial(order) ** 2 / (nbands * math.factorial(2 * order))
    ycosn = 2 * np.sqrt(const) * np.cos(xcosn) ** order * (np.abs(alpha) < np.pi / 2)
    levels = []
    for _ in range(height - 2):
        xr = xr - 1.0
        hi = _interp(log_rad, yr, xr)
        levels.append([low * hi * _interp(angle, ycosn, xcosn + np.pi * b / nbands)
                       for b in range(nbands)])
        r0, r1 = _crop(log_rad.shape[0])
        c0, c1 = _crop(log_rad.shape[1])
        log_rad, angle = log_rad[r0:r1, c0:c1], angle[r0:r1, c0:c1]
        low = low[r0:r1, c0:c1] * _interp(log_rad, np.abs(np.sqrt(1 - yr ** 2)), xr)
    return levels


# ---- folding -------------------------------------------------------------------------------


@dataclass
class LevelTables:
    level: int                 # index into the reference's coeff list (1 = finest oriented level)
    s: int                     # transform size of this level (extended domain)
    c: int                     # kept crop: outputs y, x in [0, c)
    h: int                     # folded frequency count (k, l in [0, h))
    hp: int                    # h padded to a multiple of 8
    cp: int                    # c padded to a multiple of 8
    trig: np.ndarray           # float32 [2][hp][cp]: cos / sin of pi k / S + 2 pi k y / s
    masks: np.ndarray          # float32 [nb][2 ch][2 half][hp (l)][hp (k)]  (transposed: [l][k])
    inner_sel: np.ndarray      # int32  [2 ch][2 half]: 0 -> cos table, 1 -> sin table for b_l(x)


@dataclass
class PyramidTables:
    H: int
    height: int
    nbands: int
    Hp: int
    Kp: int                    # padded number of DCT frequencies kept (max h over levels)
    dct_t: np.ndarray          # float32 [Hp][Kp]: dct_t[n][k] = 2 cos(pi k (2n+1) / (2H))
    levels: List[LevelTables] = field(default_factory=list)


def _fold(mask: np.ndarray):
    """Fold a fftshifted (s,s) mask over the signs of both frequencies.
    Returns (Mcc, Mcs, Msc, Mss) indexed [k][l] for k,l in [0, s//2]."""
    s = mask.shape[0]
    half = s // 2                      # index of frequency 0
    n = half + 1

    def signed(sign_k, sign_l):
        out = np.zeros((n, n))
        for k in range(n):
            fk = sign_k * k
            if (sign_k < 0 and k == 0) or not (-half <= fk <= s - half - 1):
                continue
            for l in range(n):
                fl = sign_l * l
                if (sign_l < 0 and l == 0) or not (-half <= fl <= s - half - 1):
                    continue
                out[k, l] = mask[fk + half, fl + half]
        return out

    pp, pm, mp, mm = signed(1, 1), signed(1, -1), signed(-1, 1), signed(-1, -1)
    return pp + pm + mp + mm, pp - pm + mp - mm, pp + pm - mp - mm, pp - pm - mp + mm


def build_tables(H: int, height: int, nbands: int, levels: Sequence[int]) -> PyramidTables:
    """Tables for mirror-extended HxH frames (S = 2H) and the requested coeff levels."""
    S = 2 * H
    if height > max_height(S):
        raise RuntimeError("Cannot build {} levels, image too small.".format(height))
    for lv in levels:
        if not 1 <= lv <= height - 2:
            raise TypeError("extract_level {} does not index an oriented level (height={})".format(lv, height))
    all_masks = oriented_band_masks(S, height, nbands)
    twist = (nbands - 1) % 4            # (-i)^(nb-1), SCFpyr_PyTorch.py:64,165-168
    out_levels: List[LevelTables] = []
    for lv in levels:
        bands = all_masks[lv - 1]
        s = bands[0].shape[0]
        c = s // 2
        folded = [_fold(m / float(s * s)) for m in bands]
        # trim all-zero trailing frequencies (e.g. the Nyquist row every low mask kills)
        h = 1
        for f4 in folded:
            for m in f4:
                nzk = np.nonzero(np.abs(m).sum(1))[0]
                nzl = np.nonzero(np.abs(m).sum(0))[0]
                h = max(h, (nzk.max() + 1) if nzk.size else 1, (nzl.max() + 1) if nzl.size else 1)
        h = min(h, H)       # DCT row H (the extended image's Nyquist) is identically zero
        hp, cp = _pad4(h), _pad4(c)
        k = np.arange(hp)[:, None]
        y = np.arange(cp)[None, :]
        ang = np.pi * k * (2 * y + s / S) / s      # = pi k / S  (mirror phase)  +  2 pi k y / s
        trig = np.stack([np.cos(ang), np.sin(ang)])
        trig[:, h:, :] = 0
        trig[:, :, c:] = 0
        masks = np.zeros((nbands, 2, 2, hp, hp))
        inner_sel = np.zeros((2, 2), np.int32)
        for b, (mcc, mcs, msc, mss) in enumerate(folded):
            re_type = ((mcc, 0), (-mss, 1))        # (mask for cos a_k rows, mask for sin a_k rows)
            im_type = ((mcs, 1), (msc, 0))
            neg = lambda t: tuple((-m, sel) for m, sel in t)
            ch = {0: (re_type, im_type), 1: (im_type, neg(re_type)),
                  2: (neg(re_type), neg(im_type)), 3: (neg(im_type), re_type)}[twist]
            for ci in range(2):
                for half in range(2):
                    m, sel = ch[ci][half]
                    masks[b, ci, half, :h, :h] = m[:h, :h].T       # stored [l][k]
                    inner_sel[ci, half] = sel
        out_levels.append(LevelTables(lv, s, c, h, hp, cp, trig.astype(np.float32),
                                      masks.astype(np.float32), inner_sel))
    kmax = max(l.h for l in out_levels)
    Hp, Kp = _pad4(H), _pad4(kmax)
    n = np.arange(Hp)[:, None]
    kk = np.arange(Kp)[None, :]
    dct_t = 2 * np.cos(np.pi * kk * (2 * n + 1) / S)
    dct_t[H:, :] = 0
    dct_t[:, kmax:] = 0
    return PyramidTables(H, height, nbands, Hp, Kp, dct_t.astype(np.float32), out_levels)


def emulate(tables: PyramidTables, frames: np.ndarray, dtype=np.float64) -> List[np.ndarray]:
    """NumPy statement of exactly what the CUDA kernel computes (same tables, same products),
    used by the CPU tests to check the folding against the oracle.  frames (N,H,H) ->
    per level (N, nb, c, c, 2)."""
    H = tables.H
    x = np.zeros((frames.shape[0], tables.Hp, tables.Hp), dtype)
    x[:, :H, :H] = frames - frames.mean(axis=(1, 2), keepdims=True)   # DC never reaches a band
    d = tables.dct_t.astype(dtype)
    ct = np.einsum("fmn,nl,mk->flk", x, d, d)                         # Ct[l][k]
    outs = []
    for lv in tables.levels:
        trig = lv.trig.astype(dtype)
        res = np.zeros((frames.shape[0], tables.nbands, lv.c, lv.c, 2), dtype)
        for b in range(tables.nbands):
            for ch in range(2):
                acc = 0
                for half in range(2):
                    v = ct[:, :lv.hp, :lv.hp] * lv.masks[b, ch, half].astype(dtype)   # [l][k]
                    u = np.einsum("flk,lx->fkx", v, trig[lv.inner_sel[ch, half]])
                    acc = acc + np.einsum("ky,fkx->fyx", trig[half], u)
                res[:, b, :, :, ch] = acc[:, :lv.c, :lv.c]
        outs.append(res)
    return outs


# ---- full pyramid of un-mirrored images (SCFpyr_PyTorch.build / reconstruct) ----------------------


@dataclass
class ScfUnit:
    s: int                     # unit size (s x s)
    planes: int                # 1 (hi0 / lo) or nbands (an oriented level)
    is_real: bool              # the reference keeps only the real part (hi0, lo)
    twist_build: int           # band spectrum times (-i)^twist (SCFpyr_PyTorch.py:64,165-168)
    twist_recon: int           # (i)^(nb-1) = (-i)^(3 (nb-1)) when reconstructing (:65,284-287)
    src_index: np.ndarray      # int32 [s]: natural-order frequency u of the unit -> natural-order index in the image spectrum
    build_mask: np.ndarray     # float32 [planes][s][s], natural order, cumulative (lo0 . lomasks . himask . anglemask)
    recon_mask: np.ndarray     # float32 [planes][s][s]


def _natural(mask_shifted: np.ndarray) -> np.ndarray:
    """ifftshift as the reference does it (roll by floor(n/2), api/steerable/math_utils.py:42-47)."""
    s0, s1 = mask_shifted.shape[-2:]
    return np.roll(mask_shifted, shift=(-(s0 // 2), -(s1 // 2)), axis=(-2, -1))


def full_pyramid_tables(size: int, height: int, nbands: int) -> List[ScfUnit]:
    """Units [hi0, level 1, ..., level height-2, lo] of SCFpyr_PyTorch.build for size x size images.

    Every fftshift (out[i] = in[(i + ceil(n/2)) % n], math_utils.py:32-40), centre crop (SCFpyr_PyTorch.py:179-190)
    and ifftshift of the reference is an index permutation; they are composed here into one gather index per unit,
    and the real masks the spectrum meets on its way to a unit are multiplied up in float64."""
    if height > max_height(size):
        raise RuntimeError("Cannot build {} levels, image too small.".format(height))
    if nbands < 2:
        raise RecursionError("nbands must be >= 2 (the reference's factorial(0) never terminates, "
                             "api/steerable/math_utils.py:79-84)")
    log_rad, angle = _polar_grid(size, size)
    xr, yr = _rcos()
    lo0 = _interp(log_rad, np.sqrt(1 - yr ** 2), xr)
    hi0 = _interp(log_rad, yr, xr)
    lutsize = 1024
    xcosn = np.pi * np.arange(-(2 * lutsize + 1), lutsize + 2) / lutsize
    alpha = (xcosn + np.pi) % (2 * np.pi) - np.pi
    order = nbands - 1
    const = (2 ** (2 * order)) * math.factorial(order) ** 2 / (nbands * math.factorial(2 * order))
    ycosn_build = 2 * np.sqrt(const) * np.cos(xcosn) ** order * (np.abs(alpha) < np.pi / 2)
    ycosn_recon = np.sqrt(const) * np.cos(xcosn) ** order                    # SCFpyr_PyTorch.py:268
    tb, tr = (nbands - 1) % 4, (3 * (nbands - 1)) % 4

    def unit(idx_shifted, planes, is_real, bm, rm, twb, twr):
        s = idx_shifted.shape[0]
        src = np.roll(idx_shifted, -(s // 2)).astype(np.int32)               # ifftshift of the index map
        return ScfUnit(s, planes, is_real, twb, twr, src, np.ascontiguousarray(_natural(bm), np.float32),
                       np.ascontiguousarray(_natural(rm), np.float32))

    idx = (np.arange(size) + (size + 1) // 2) % size                         # shifted position -> natural index
    units = [unit(idx, 1, True, hi0[None], hi0[None], 0, 0)]
    low = lo0
    for _ in range(height - 2):
        xr = xr - 1.0
        hi = _interp(log_rad, yr, xr)
        bm = np.stack([low * hi * _interp(angle, ycosn_build, xcosn + np.pi * b / nbands) for b in range(nbands)])
        rm = np.stack([low * hi * _interp(angle, ycosn_recon, xcosn + np.pi * b / nbands) for b in range(nbands)])
        units.append(unit(idx, nbands, False, bm, rm, tb, tr))
        r0, r1 = _crop(log_rad.shape[0])
        c0, c1 = _crop(log_rad.shape[1])
        log_rad, angle = log_rad[r0:r1, c0:c1], angle[r0:r1, c0:c1]
        idx = idx[r0:r1]
        low = low[r0:r1, c0:c1] * _interp(log_rad, np.abs(np.sqrt(1 - yr ** 2)), xr)
    units.append(unit(idx, 1, True, low[None], low[None], 0, 0))
    return units


def emulate_full(units: List[ScfUnit], images: np.ndarray):
    """NumPy statement of mimamo_scf_build (float64): images (N,S,S) -> coeff list shaped like the reference's."""
    F = np.fft.fft2(images.astype(np.float64))
    out = []
    for u in units:
        z = F[:, u.src_index][:, :, u.src_index][None] * u.build_mask.astype(np.float64)[:, None]
        z = z * ((-1j) ** u.twist_build)
        c = np.fft.ifft2(z)
        out.append(c[0].real if u.is_real else [c[b] for b in range(u.planes)])
    return out


def emulate_reconstruct(units: List[ScfUnit], coeff, size: int):
    """NumPy statement of mimamo_scf_reconstruct (float64)."""
    acc = None
    for u, c in zip(units, coeff):
        planes = np.stack(c) if not u.is_real else np.asarray(c)[None]
        B = np.fft.fft2(planes.astype(np.complex128))
        contrib = (B * u.recon_mask.astype(np.float64)[:, None]).sum(0) * ((-1j) ** u.twist_recon)
        if acc is None:
            acc = np.zeros((contrib.shape[0], size, size), np.complex128)
        acc[np.ix_(np.arange(acc.shape[0]), u.src_index, u.src_index)] += contrib
    return np.fft.ifft2(acc).real
