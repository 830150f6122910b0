"""CPU oracle: a plain restatement of MIMAMO-Net's per-window inference hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module, and
only as the checker.  The product (`mimamo-net_b200/`) never routes through it.

Parity status
-------------
* Rows P0-P3, H, T of SURVEY.md section 8 are PINNED: `oracle/make_golden.py`
  runs the unmodified reference (through `oracle/ref_shim.py`) on seeded inputs
  in the build container, checks this restatement against it, and commits the
  reference's outputs under `tests/golden/`.
* Row R (ResNet50 `resnet50_ferplus_dag`): the architecture file and weights are
  a third-party download (albanie/pytorch-benchmarks, unpinned master,
  reference api/readme.md:60-74) that is absent from /root/reference, so
  `FerPlusResNet50` restates the published Caffe-style architecture and
  **parity with the upstream weights is unpinned**.

Every function cites the reference lines (relative to /root/reference) it follows.
All pyramid math is written against modern `torch.fft`; `dtype` selects fp32
(the reference's arithmetic) or fp64 (used to measure the fp32 noise floor).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# P1a: data-independent masks (host, float64)
# --------------------------------------------------------------------------------------


def polar_grid(m: int, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """log2-radius and angle grids, api/steerable/math_utils.py:52-60."""
    gx = np.linspace(-(m // 2) / (m / 2), (m // 2) / (m / 2) - (1 - m % 2) * 2 / m, num=m)
    gy = np.linspace(-(n // 2) / (n / 2), (n // 2) / (n / 2) - (1 - n % 2) * 2 / n, num=n)
    col, row = np.meshgrid(gy, gx)          # col[i, j] = gy[j], row[i, j] = gx[i]
    angle = np.arctan2(row, col)
    rad = np.sqrt(col ** 2 + row ** 2)
    rad[m // 2][n // 2] = rad[m // 2][n // 2 - 1]     # patch log2(0) at DC
    return np.log2(rad), angle


def raised_cosine(width: float = 1.0, position: float = -0.5) -> Tuple[np.ndarray, np.ndarray]:
    """259-point raised-cosine LUT, api/steerable/math_utils.py:62-69."""
    n = 256
    x = np.pi * np.arange(-n - 1, 2) / 2 / n
    y = np.cos(x) ** 2
    y[0] = y[1]
    y[n + 2] = y[n + 1]
    return position + 2 * width / np.pi * (x + np.pi / 4), y


def lut(values: np.ndarray, ys: np.ndarray, xs: np.ndarray) -> np.ndarray:
    """pointOp == np.interp, api/steerable/math_utils.py:71-73."""
    return np.interp(values.ravel(), xs, ys).reshape(values.shape)


def crop_bounds(dim: int) -> Tuple[int, int]:
    """Centre-crop indices of the low band, api/steerable/SCFpyr_PyTorch.py:179-183."""
    start = int(np.ceil((dim + 0.5) / 2) - np.ceil((np.ceil((dim - 0.5) / 2) + 0.5) / 2))
    return start, start + int(np.ceil((dim - 0.5) / 2))


def angular_lut(nbands: int) -> Tuple[np.ndarray, np.ndarray]:
    """Xcosn / Ycosn, api/steerable/SCFpyr_PyTorch.py:61-63,148-150."""
    lutsize = 1024
    xcosn = np.pi * np.arange(-(2 * lutsize + 1), lutsize + 2) / lutsize
    alpha = (xcosn + np.pi) % (2 * np.pi) - np.pi
    order = nbands - 1
    const = (2 ** (2 * order)) * (math.factorial(order) ** 2) / (nbands * math.factorial(2 * order))
    ycosn = 2 * np.sqrt(const) * np.cos(xcosn) ** order * (np.abs(alpha) < np.pi / 2)
    return xcosn, ycosn


def pyramid_masks(size_rows: int, size_cols: int, height: int, nbands: int) -> Dict:
    """All masks in fftshifted index order, following SCFpyr_PyTorch.py:94-107,139-158,193-196.

    Returns {'lo0','hi0', 'levels': [{'hi','angle':[nb],'lo','crop':((r0,r1),(c0,c1))}, ...]}.
    The reference calls prepare_grid(shape[2], shape[1]) (SCFpyr_PyTorch.py:87,94), which only
    broadcasts for square inputs; we keep that call order.
    """
    log_rad, angle = polar_grid(size_cols, size_rows)
    xr, yr = raised_cosine(1, -0.5)
    yr = np.sqrt(yr)
    yir = np.sqrt(1 - yr ** 2)
    out = {"lo0": lut(log_rad, yir, xr), "hi0": lut(log_rad, yr, xr), "levels": []}
    xcosn, ycosn = angular_lut(nbands)
    for _ in range(height - 2):
        xr = xr - 1.0                                          # log2(scale_factor = 2)
        lvl = {"hi": lut(log_rad, yr, xr),
               "angle": [lut(angle, ycosn, xcosn + np.pi * b / nbands) for b in range(nbands)]}
        r0, r1 = crop_bounds(log_rad.shape[0])
        c0, c1 = crop_bounds(log_rad.shape[1])
        log_rad = log_rad[r0:r1, c0:c1]
        angle = angle[r0:r1, c0:c1]
        lvl["crop"] = ((r0, r1), (c0, c1))
        lvl["lo"] = lut(log_rad, np.abs(np.sqrt(1 - yr ** 2)), xr)
        out["levels"].append(lvl)
    return out


# --------------------------------------------------------------------------------------
# P0 / P1b-d: symmetric extension and the pyramid itself
# --------------------------------------------------------------------------------------


def symmetric_extension(x: torch.Tensor) -> torch.Tensor:
    """[[x, x mirrored in w], [x mirrored in h, both]], api/utils/phase_utils.py:116-129."""
    top = torch.cat([x, x.flip(-1)], dim=-1)
    return torch.cat([top, top.flip(-2)], dim=-2)


def _shift(x: torch.Tensor) -> torch.Tensor:
    """fftshift over dims 1,2: out[i] = in[i + ceil(n/2)], api/steerable/math_utils.py:32-40."""
    return torch.roll(x, shifts=(-((x.shape[1] + 1) // 2), -((x.shape[2] + 1) // 2)), dims=(1, 2))


def _unshift(x: torch.Tensor) -> torch.Tensor:
    """ifftshift over dims 1,2 (roll by floor(n/2)), api/steerable/math_utils.py:42-47."""
    return torch.roll(x, shifts=(-(x.shape[1] // 2), -(x.shape[2] // 2)), dims=(1, 2))


def pyramid_build(images: torch.Tensor, height: int, nbands: int,
                  dtype: torch.dtype = torch.float32) -> List:
    """SCFpyr_PyTorch.build + _build_levels, api/steerable/SCFpyr_PyTorch.py:70-208.

    images (N,1,S,S) -> [hi0 (N,S,S), [nb x (N,S,S,2)], [nb x (N,S/2,S/2,2)], ..., lo (N,s,s)].
    """
    assert images.dim() == 4 and images.shape[1] == 1
    x = images[:, 0].to(dtype)
    rows, cols = x.shape[1], x.shape[2]
    if height > int(np.floor(np.log2(min(rows, cols))) - 2):
        raise RuntimeError("Cannot build {} levels, image too small.".format(height))
    cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
    masks = pyramid_masks(rows, cols, height, nbands)
    as_t = lambda m: torch.from_numpy(m).to(dtype)[None]
    dft = _shift(torch.fft.fft2(x).to(cdtype))
    lodft = dft * as_t(masks["lo0"])
    twist = complex(0, -1) ** (nbands - 1)                      # SCFpyr_PyTorch.py:64
    coeff: List = []
    for lvl in masks["levels"]:
        bands = []
        for b in range(nbands):
            banddft = lodft * as_t(lvl["angle"][b]) * as_t(lvl["hi"])
            banddft = torch.complex(twist.real * banddft.real - twist.imag * banddft.imag,
                                    twist.real * banddft.imag + twist.imag * banddft.real)
            bands.append(torch.view_as_real(torch.fft.ifft2(_unshift(banddft))))
        coeff.append(bands)
        (r0, r1), (c0, c1) = lvl["crop"]
        lodft = lodft[:, r0:r1, c0:c1] * as_t(lvl["lo"])
    coeff.append(torch.fft.ifft2(_unshift(lodft)).real)
    hi0 = torch.fft.ifft2(_unshift(dft * as_t(masks["hi0"]))).real
    coeff.insert(0, hi0)
    return coeff


def pyramid_reconstruct(coeff: List, nbands: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """SCFpyr_PyTorch.reconstruct + _reconstruct_levels, api/steerable/SCFpyr_PyTorch.py:214-314.

    coeff as returned by pyramid_build -> images (N,S,S).  Reconstruction uses its own angular LUT
    (sqrt(const) cos^order, no half-plane indicator, :266-268) and the factor (i)^(nbands-1) (:65,284-287)."""
    if nbands != len(coeff[1]):
        raise Exception("Unmatched number of orientations")
    rows, cols = coeff[0].shape[1], coeff[0].shape[2]
    cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
    log_rad, angle = polar_grid(cols, rows)
    xr, yr = raised_cosine(1, -0.5)
    yr = np.sqrt(yr)
    yir = np.sqrt(np.abs(1 - yr ** 2))
    as_t = lambda m: torch.from_numpy(m).to(dtype)[None]
    lutsize = 1024
    xcosn = np.pi * np.arange(-(2 * lutsize + 1), lutsize + 2) / lutsize
    order = nbands - 1
    const = (2 ** (2 * order)) * (math.factorial(order) ** 2) / (nbands * math.factorial(2 * order))
    ycosn = np.sqrt(const) * np.cos(xcosn) ** order
    twist = complex(0, 1) ** (nbands - 1)

    def levels(cf, log_rad, angle, xr):
        if len(cf) == 1:
            return _shift(torch.fft.fft2(cf[0].to(dtype)).to(cdtype))
        xr = xr - 1.0
        himask = as_t(lut(log_rad, yr, xr))
        orient = torch.zeros(cf[0][0].shape[:-1], dtype=cdtype)
        for b in range(nbands):
            anglemask = as_t(lut(angle, ycosn, xcosn + np.pi * b / nbands))
            band = torch.view_as_complex(cf[0][b].to(dtype).contiguous()).to(cdtype)
            banddft = _shift(torch.fft.fft2(band)) * anglemask * himask
            orient = orient + torch.complex(twist.real * banddft.real - twist.imag * banddft.imag,
                                            twist.real * banddft.imag + twist.imag * banddft.real)
        (r0, r1), (c0, c1) = crop_bounds(log_rad.shape[0]), crop_bounds(log_rad.shape[1])
        nlog_rad, nangle = log_rad[r0:r1, c0:c1], angle[r0:r1, c0:c1]
        lomask = as_t(lut(nlog_rad, yir, xr))
        nres = levels(cf[1:], nlog_rad, nangle, xr)
        res = torch.zeros_like(orient)
        res[:, r0:r1, c0:c1] = nres * lomask
        return res + orient

    temp = levels(coeff[1:], log_rad, angle, xr)
    hidft = _shift(torch.fft.fft2(coeff[0].to(dtype)).to(cdtype))
    out = temp * as_t(lut(log_rad, yir, xr)) + hidft * as_t(lut(log_rad, yr, xr))
    return torch.fft.ifft2(_unshift(out)).real


def build_pyramid(im_batch: torch.Tensor, height: int, nbands: int, extract_level,
                  symmetry: bool = True, dtype: torch.dtype = torch.float32):
    """Phase_Difference_Extractor.build_pyramid, api/phase_difference_extractor.py:38-87.

    im_batch (bs,T,H,H) -> per requested level (bs,nb,T,c,c,2), c = s_level/2 when symmetry.
    """
    bs, t, w, h = im_batch.shape
    flat = im_batch.reshape(bs * t, 1, w, h).to(dtype)
    if symmetry:
        flat = symmetric_extension(flat)
    coeff = pyramid_build(flat, height, nbands, dtype)

    def one(level):
        stacked = torch.stack(coeff[level], 0)                  # (nb, bs*T, s, s, 2)
        s0, s1 = stacked.shape[-3], stacked.shape[-2]
        out = stacked.view(nbands, bs, t, s0, s1, 2).permute(1, 0, 2, 3, 4, 5).contiguous()
        return out[..., : s0 // 2, : s1 // 2, :] if symmetry else out

    if isinstance(extract_level, int):
        return one(extract_level)
    return [one(l) for l in extract_level]


# --------------------------------------------------------------------------------------
# P2: phase tail
# --------------------------------------------------------------------------------------


def unwrap_positive_jumps(phase: torch.Tensor, dim: int) -> torch.Tensor:
    """torch_unwrap, api/utils/phase_utils.py:5-20.  Uses C fmod, so only jumps > +pi are
    corrected; kept bug-for-bug."""
    pi = math.pi
    n = phase.shape[dim]
    dd = phase.narrow(dim, 1, n - 1) - phase.narrow(dim, 0, n - 1)
    ddmod = torch.fmod(dd + pi, 2 * pi) - pi
    ddmod = torch.where((ddmod == -pi) & (dd > 0), torch.full_like(ddmod, pi), ddmod)
    corr = ddmod - dd
    corr = torch.where(dd.abs() < pi, torch.zeros_like(corr), corr)
    out = phase.clone()
    out.narrow(dim, 1, n - 1).add_(corr.cumsum(dim=dim))
    return out


def gaussian_taps(std: float = 2.0, tap: int = 11) -> np.ndarray:
    """Unnormalised 2-D Gaussian, api/utils/phase_utils.py:108-115."""
    r = np.arange(tap) - tap // 2
    return np.exp(-(r[:, None] ** 2 + r[None, :] ** 2) / (2 * std ** 2))


def amplitude_weighted_blur(mag: torch.Tensor, phase: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """(G*(mag.phase))/(G*mag), zero padded, api/utils/phase_utils.py:78-90."""
    ch = phase.shape[1]
    k = kernel.to(phase.dtype)[None, None].expand(ch, 1, -1, -1).contiguous()
    pad = kernel.shape[0] // 2
    num = F.conv2d(mag * phase, k, groups=ch, padding=pad)
    den = F.conv2d(mag, k, groups=ch, padding=pad)
    return num / den


def extract(coeff: torch.Tensor) -> torch.Tensor:
    """Phase_Difference_Extractor.extract, api/phase_difference_extractor.py:93-134.

    coeff (bs,nb,T,w,h,2) -> (bs,nb,T-1,w,h): atan2, |.|+1e-10, unwrap over T, amplitude
    weighted blur, temporal difference, spatial-mean removal, clamp to +-5pi.
    """
    bs, nb, t, w, h, _ = coeff.shape
    re, im = coeff[..., 0], coeff[..., 1]
    phase = torch.atan2(im, re).reshape(bs * nb, t, w, h)
    mag = torch.sqrt(im ** 2 + re ** 2).reshape(bs * nb, t, w, h) + 1e-10
    phase = unwrap_positive_jumps(phase, dim=-3)
    smooth = amplitude_weighted_blur(mag, phase, torch.from_numpy(gaussian_taps(2, 11)))
    smooth = smooth.view(bs, nb, t, w, h)
    delta = smooth[:, :, 1:] - smooth[:, :, :-1]
    delta = delta - delta.mean(-1).mean(-1)[..., None, None]
    return torch.clamp(delta, -5 * math.pi, 5 * math.pi)


def extract_phase(coeff: torch.Tensor, return_phase: bool = False, return_both: bool = False) -> torch.Tensor:
    """Steerable_Pyramid_Phase.extract_phase, Aff-wild-exps/utils.py:367-418 (the training-side superset of `extract`).

    coeff (bs,nb,T,w,h,2).  default -> phase differences (bs,nb,T-1,w,h) (same as `extract`); return_phase ->
    denoised phase minus its spatial mean (bs,nb,T,w,h); return_both -> insert_tensors(differences, phases[:, :, 1:])
    (bs,nb,2(T-1),w,h), where insert_tensors (:419-432) loops over range(T-1) only: slot i < T-1 holds difference i//2
    (i even) or phase 1 + i//2 (i odd) and the remaining T-1 slots stay zero -- kept bug for bug."""
    bs, nb, t, w, h, _ = coeff.shape
    re, im = coeff[..., 0], coeff[..., 1]
    phase = torch.atan2(im, re).reshape(bs * nb, t, w, h)
    mag = torch.sqrt(im ** 2 + re ** 2).reshape(bs * nb, t, w, h) + 1e-10
    phase = unwrap_positive_jumps(phase, dim=-3)
    smooth = amplitude_weighted_blur(mag, phase, torch.from_numpy(gaussian_taps(2, 11))).view(bs, nb, t, w, h)
    delta = smooth[:, :, 1:] - smooth[:, :, :-1]
    smooth = smooth - smooth.mean(-1).mean(-1)[..., None, None]
    delta = delta - delta.mean(-1).mean(-1)[..., None, None]
    delta = torch.clamp(delta, -5 * math.pi, 5 * math.pi)
    if return_both:
        rest = smooth[:, :, 1:]
        out = torch.zeros(bs, nb, 2 * (t - 1), w, h, dtype=delta.dtype)
        for i in range(t - 1):
            out[:, :, i] = delta[:, :, i // 2] if i % 2 == 0 else rest[:, :, i // 2]
        return out
    return smooth if return_phase else delta


def phase_diff_output(phase_batch: torch.Tensor, height: int = 4, nbands: int = 2,
                      extract_level: Sequence[int] = (1, 2), dtype: torch.dtype = torch.float32):
    """Tester.phase_diff_output, api/tester.py:122-139: (B,F,T,H,H) -> per level (B,F,nb*(T-1),c,c)."""
    b, f, t, w, h = phase_batch.shape
    coeffs = build_pyramid(phase_batch.reshape(b * f, t, w, h), height, nbands, list(extract_level),
                           dtype=dtype)
    outs = []
    for c in coeffs:
        d = extract(c)
        outs.append(d.reshape(b, f, -1, d.shape[-2], d.shape[-1]))
    return outs


# --------------------------------------------------------------------------------------
# T: window / snippet index rules and stitching (host logic)
# --------------------------------------------------------------------------------------


def snippet_ranges(n_frames: int, length: int = 64, stride: int = 64) -> List[List[int]]:
    """Snippet_Sampler.parse_video, api/sampler/snippet_sampler.py:107-128."""
    if n_frames < length:
        length = stride = n_frames
    out, start, end = [], 0, length
    while end <= n_frames and start < n_frames:
        out.append([start, end])
        start += stride
        end = start + length
    assert out, "No snippet is sampled."
    if out[-1][1] < n_frames:
        out.append([n_frames - length, n_frames])
    return out


def window_frame_ids(frame: int, n_frames: int, num_phase: int = 12) -> List[int]:
    """Clamped temporal window of a frame, api/sampler/snippet_sampler.py:144-152."""
    return [min(max(0, frame + i - num_phase // 2), n_frames - 1) for i in range(num_phase + 1)]


def gather_windows(gray: torch.Tensor, start: int, end: int, num_phase: int = 12) -> torch.Tensor:
    """gray (n,H,H) -> (end-start, num_phase+1, H, H) using the clamp-window rule."""
    n = gray.shape[0]
    idx = torch.tensor([window_frame_ids(f, n, num_phase) for f in range(start, end)])
    return gray[idx]


def stitch(ranges: Sequence[Sequence[int]], preds: Sequence[np.ndarray]) -> np.ndarray:
    """Per-video stitching, api/tester.py:104-118 (later snippets overwrite the overlap)."""
    max_len = max(r[-1] for r in ranges)
    out = np.zeros((max_len, preds[0].shape[-1]))
    lo, hi = 0, 0
    for (s, e), p in zip(ranges, preds):
        out[s:e, :] = p
        lo, hi = min(lo, s), max(hi, e)
    assert lo == 0 and hi == max_len
    return out


# --------------------------------------------------------------------------------------
# H: two-stream head, functional over a reference-keyed state_dict
# --------------------------------------------------------------------------------------


def _bn(x, sd, key, eps=1e-5):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                        sd[key + ".weight"], sd[key + ".bias"], training=False, eps=eps)


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one GRU layer over dim 0 (torch.nn.GRU gate order r,z,n)."""
    steps, batch, _ = x.shape
    hid = w_hh.shape[1]
    h = x.new_zeros(batch, hid)
    out = x.new_zeros(steps, batch, hid)
    order = range(steps - 1, -1, -1) if reverse else range(steps)
    for s in order:
        gi = x[s] @ w_ih.t() + b_ih
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :hid] + gh[:, :hid])
        z = torch.sigmoid(gi[:, hid:2 * hid] + gh[:, hid:2 * hid])
        n = torch.tanh(gi[:, 2 * hid:] + r * gh[:, 2 * hid:])
        h = (1 - z) * n + z * h
        out[s] = h
    return out


def mlp_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "mlp.mlp.") -> torch.Tensor:
    """MLP.forward in eval mode, api/mimamo_net.py:6-26: (rows, features) -> (rows, 256); [Dropout, Linear, BN, ReLU]
    per hidden layer (keys 4i+1 / 4i+2 of the nn.Sequential), any depth."""
    i = 0
    while prefix + "%d.weight" % (4 * i + 1) in sd:
        lin, bn = prefix + str(4 * i + 1), prefix + str(4 * i + 2)
        x = F.relu(_bn(F.linear(x, sd[lin + ".weight"], sd[lin + ".bias"]), sd, bn))
        i += 1
    return x


def phasenet_forward(sd: Dict[str, torch.Tensor], l0: torch.Tensor, l1: torch.Tensor, prefix: str = "phasenet.",
                     feature: bool = True) -> torch.Tensor:
    """PhaseNet.forward in eval mode, api/mimamo_net.py:79-95: (rows,C,S,S), (rows,C,S/2,S/2) with S in {48, 96, 112} ->
    (rows,256) when `feature`, else (rows,1) after the classifier Linear(256,1) + BatchNorm1d(1, eps=1e-6) (:62-64)."""
    def conv_block(t, blk, stride2):
        p = prefix + "conv_net.%d." % blk
        t = F.relu(_bn(F.conv2d(t, sd[p + "0.weight"], sd[p + "0.bias"], padding=1), sd, p + "1"))
        return F.relu(_bn(F.conv2d(t, sd[p + "3.weight"], sd[p + "3.bias"], padding=1, stride=stride2),
                          sd, p + "4"))

    t = torch.cat([conv_block(l0, 0, 2), l1], dim=1)
    blk = 1
    while prefix + "conv_net.%d.0.weight" % blk in sd:                          # 2 more blocks for 48x48 inputs, 3 for 96 / 112 (:33-40)
        t = conv_block(t, blk, 2)
        blk += 1
    t = F.avg_pool2d(t, kernel_size=t.shape[-1]).reshape(l0.shape[0], -1)
    for lin, bn in ((0, 2), (4, 6)):                                            # fc: Linear, ReLU, BN, Dropout
        t = _bn(F.relu(F.linear(t, sd[prefix + "fc.%d.weight" % lin], sd[prefix + "fc.%d.bias" % lin])),
                sd, prefix + "fc.%d" % bn)
    if feature:
        return t
    y = F.linear(t, sd[prefix + "classifier.0.weight"], sd[prefix + "classifier.0.bias"])
    return _bn(y, sd, prefix + "classifier.1", eps=1e-6)


def head_forward(sd: Dict[str, torch.Tensor], phase_0, phase_1, rgb) -> torch.Tensor:
    """Two_Stream_RNN.forward in eval mode, api/mimamo_net.py:129-143 (MLP :22-26,
    PhaseNet :79-95).  NOTE the GRU (built without batch_first, :119) recurs over dim 0 =
    the snippet axis, batch = frames."""
    bs, nf = rgb.shape[0], rgb.shape[1]
    spatial = mlp_forward(sd, rgb.reshape(bs * nf, -1))
    t = phasenet_forward(sd, phase_0.reshape(bs * nf, *phase_0.shape[2:]), phase_1.reshape(bs * nf, *phase_1.shape[2:]))
    feat = torch.cat([spatial, t], dim=-1)
    feat = _bn(F.relu(F.linear(feat, sd["transform.0.weight"], sd["transform.0.bias"])), sd, "transform.2")
    seq = feat.view(bs, nf, -1)
    for layer in range(2):
        halves = []
        for sfx, rev in (("", False), ("_reverse", True)):
            k = "_l%d%s" % (layer, sfx)
            halves.append(_gru_direction(seq, sd["rnns.weight_ih" + k], sd["rnns.weight_hh" + k],
                                         sd["rnns.bias_ih" + k], sd["rnns.bias_hh" + k], rev))
        seq = torch.cat(halves, dim=-1)
    y = F.linear(seq.reshape(bs * nf, -1), sd["classifier.1.weight"], sd["classifier.1.bias"])
    return _bn(y, sd, "classifier.2").view(bs, nf, -1)


def head_state_dict_spec(num_phase: int = 12) -> List[Tuple[str, Tuple[int, ...]]]:
    """Key names/shapes of Two_Stream_RNN.state_dict()."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def lin(p, o, i):
        spec.extend([(p + ".weight", (o, i)), (p + ".bias", (o,))])

    def bn(p, c):
        spec.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                     (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())])

    lin("mlp.mlp.1", 256, 2048); bn("mlp.mlp.2", 256); lin("mlp.mlp.5", 256, 256); bn("mlp.mlp.6", 256)
    chans = [(2 * num_phase, 64), (2 * num_phase + 64, 128), (128, 256)]
    for b, (ci, co) in enumerate(chans):
        p = "phasenet.conv_net.%d." % b
        spec.extend([(p + "0.weight", (co, ci, 3, 3)), (p + "0.bias", (co,))]); bn(p + "1", co)
        spec.extend([(p + "3.weight", (co, co, 3, 3)), (p + "3.bias", (co,))]); bn(p + "4", co)
    lin("phasenet.fc.0", 256, 256); bn("phasenet.fc.2", 256)
    lin("phasenet.fc.4", 256, 256); bn("phasenet.fc.6", 256)
    lin("phasenet.classifier.0", 1, 256); bn("phasenet.classifier.1", 1)
    lin("transform.0", 256, 512); bn("transform.2", 256)
    for layer in range(2):
        for sfx in ("", "_reverse"):
            k = "_l%d%s" % (layer, sfx)
            spec.extend([("rnns.weight_ih" + k, (384, 256)), ("rnns.weight_hh" + k, (384, 128)),
                         ("rnns.bias_ih" + k, (384,)), ("rnns.bias_hh" + k, (384,))])
    lin("classifier.1", 2, 256); bn("classifier.2", 2)
    return spec


def synthetic_state_dict(spec, seed: int, weight_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters: weights ~ He-normal, biases
    N(0,0.1), BN gamma N(1,0.1), beta N(0,0.1), running_mean N(0,0.1), running_var U[0.5,1.5]."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in spec:
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(1, dtype=torch.int64)
        elif name.endswith("running_var"):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1 and name.endswith(".weight"):            # BN gamma
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                                         # biases (incl. GRU)
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            sd[name] = torch.randn(shape, generator=g) * (weight_gain * math.sqrt(2.0 / fan_in))
    return sd


# --------------------------------------------------------------------------------------
# R: ResNet50 (resnet50_ferplus_dag architecture), restated -- parity with upstream unpinned
# --------------------------------------------------------------------------------------

RESNET_MEAN = (131.0912, 103.8827, 91.4953)      # meta['mean'], 0-255 scale, std = 1
RESNET_STAGES = ((2, 3, 64, 256, 1), (3, 4, 128, 512, 2), (4, 6, 256, 1024, 2), (5, 3, 512, 2048, 2))


class FerPlusResNet50(nn.Module):
    """Caffe-style ResNet-50 as exported by albanie/pytorch-benchmarks `resnet50_ferplus_dag`
    (SURVEY.md section 8(a) row R): stride 2 sits on the first 1x1 (`_reduce`) and on `_proj`
    of stages 3-5, pool1 is MaxPool(3,2,pad 0,ceil_mode), `pool5_7x7_s1` is the tapped layer
    (api/resnet50_extractor.py:74-83).  Module names follow the upstream file so the
    reference's forward-hook lookup `model._modules.get('pool5_7x7_s1')` works."""

    def __init__(self):
        super().__init__()
        self.meta = {"mean": list(RESNET_MEAN), "std": [1, 1, 1], "imageSize": [224, 224, 3]}
        self.conv1_7x7_s2 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.conv1_7x7_s2_bn = nn.BatchNorm2d(64)
        self.pool1_3x3_s2 = nn.MaxPool2d(3, stride=2, padding=0, ceil_mode=True)
        cin = 64
        for stage, blocks, mid, cout, stride in RESNET_STAGES:
            for blk in range(1, blocks + 1):
                p = "conv%d_%d_" % (stage, blk)
                s = stride if blk == 1 else 1
                self._add(p + "1x1_reduce", cin, mid, 1, s, 0)
                self._add(p + "3x3", mid, mid, 3, 1, 1)
                self._add(p + "1x1_increase", mid, cout, 1, 1, 0)
                if blk == 1:
                    self._add(p + "1x1_proj", cin, cout, 1, s, 0)
                cin = cout
        self.pool5_7x7_s1 = nn.AvgPool2d(7, stride=1, padding=0)
        self.classifier = nn.Conv2d(2048, 8, 1)

    def _add(self, name, cin, cout, k, s, p):
        setattr(self, name, nn.Conv2d(cin, cout, k, stride=s, padding=p, bias=False))
        setattr(self, name + "_bn", nn.BatchNorm2d(cout))

    def _cb(self, name, x, relu=True):
        y = getattr(self, name + "_bn")(getattr(self, name)(x))
        return F.relu(y) if relu else y

    def features(self, x):
        x = self.pool1_3x3_s2(self._cb("conv1_7x7_s2", x))
        for stage, blocks, *_ in RESNET_STAGES:
            for blk in range(1, blocks + 1):
                p = "conv%d_%d_" % (stage, blk)
                y = self._cb(p + "1x1_increase", self._cb(p + "3x3", self._cb(p + "1x1_reduce", x)), relu=False)
                skip = self._cb(p + "1x1_proj", x, relu=False) if blk == 1 else x
                x = F.relu(y + skip)
        return self.pool5_7x7_s1(x)

    def forward(self, x):
        return self.classifier(self.features(x))


def resnet_synthetic(seed: int = 1) -> FerPlusResNet50:
    """Seeded He-normal convs + non-identity BN statistics."""
    net = FerPlusResNet50().eval()
    spec = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = synthetic_state_dict(spec, seed)
    # keep the residual stream O(1): damp the last BN of every block, as zero-init-residual does
    for k in sd:
        if k.endswith("1x1_increase_bn.weight"):
            sd[k] = sd[k] * 0.25
    # inputs are 0-255 minus mean (|x| ~ 100): bring activations to O(1) like a trained net
    sd["conv1_7x7_s2.weight"] = sd["conv1_7x7_s2.weight"] * 0.02
    net.load_state_dict(sd)
    return net


def resnet_pool5(net: FerPlusResNet50, image: torch.Tensor) -> torch.Tensor:
    """Resnet50_Extractor.get_vec, api/resnet50_extractor.py:74-83: (bs,3,224,224) 0-255 minus
    mean -> relu(pool5) as (bs,2048).  (The reference's .squeeze() collapses bs==1; we keep 2-D.)"""
    with torch.no_grad():
        return F.relu(net.features(image).flatten(1))
