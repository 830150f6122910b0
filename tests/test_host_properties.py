"""CPU: property tests (hypothesis) of the host-side index rules and resampling tables against the oracle's restatement of
the reference (api/sampler/snippet_sampler.py:107-152, api/tester.py:104-118, Pillow's precompute_coeffs) -- ragged lengths,
clips shorter than a snippet, strides that do not divide the video, odd window widths."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import mimamo_b200  # noqa: F401  (puts the drop-in api/ on sys.path)
from oracle import mimamo_oracle as O
from oracle import pil_preproc as P


@settings(max_examples=300, deadline=None)
@given(n=st.integers(1, 700), length=st.sampled_from([8, 16, 64]), stride_frac=st.floats(0.05, 1.0))
def test_snippet_ranges(n, length, stride_frac):
    from sampler.snippet_sampler import snippet_ranges
    stride = max(1, int(round(length * stride_frac)))
    got = snippet_ranges(n, length, stride)
    assert got == O.snippet_ranges(n, length, stride)
    covered = np.zeros(n, dtype=bool)
    for s, e in got:
        assert 0 <= s < e <= n and e - s == min(length, n)
        covered[s:e] = True
    assert got[0][0] == 0 and got[-1][1] == n                  # the tail snippet always reaches the last frame
    if stride <= min(length, n):
        assert covered.all()


@settings(max_examples=300, deadline=None)
@given(n=st.integers(1, 400), num_phase=st.integers(1, 16), data=st.data())
def test_window_frame_ids(n, num_phase, data):
    from sampler.snippet_sampler import window_frame_ids, window_index
    f = data.draw(st.integers(0, n - 1))
    ids = window_frame_ids(f, n, num_phase)
    assert ids == O.window_frame_ids(f, n, num_phase)
    assert len(ids) == num_phase + 1 and all(0 <= i < n for i in ids) and ids == sorted(ids)
    assert f in ids
    lo = data.draw(st.integers(0, n - 1))
    hi = data.draw(st.integers(lo, n))
    idx = window_index(lo, hi, n, num_phase)
    assert tuple(idx.shape) == (hi - lo, num_phase + 1) or hi == lo
    for r, frame in enumerate(range(lo, hi)):
        assert idx[r].tolist() == O.window_frame_ids(frame, n, num_phase)


@settings(max_examples=100, deadline=None)
@given(lengths=st.lists(st.integers(1, 300), min_size=1, max_size=4), seed=st.integers(0, 2 ** 16))
def test_stitching_of_several_videos(lengths, seed):
    """Snippets of several videos interleaved in one prediction list: every video is stitched from its own ranges, the tail
    snippet overwrites the overlap (api/tester.py:104-118)."""
    from tester import stitch_predictions
    rs = np.random.RandomState(seed)
    names, ranges, preds = [], [], []
    per_video = {}
    for v, n in enumerate(lengths):
        rr = O.snippet_ranges(n)
        pp = [rs.rand(e - s, 2) for s, e in rr]
        per_video["v%d" % v] = (rr, pp)
        for r, p in zip(rr, pp):
            names.append("v%d" % v); ranges.append(r); preds.append(p)
    length = max(p.shape[0] for p in preds)
    if any(p.shape[0] != length for p in preds):
        return                                                 # the reference stacks equal-length snippets only (np.concatenate of batches)
    got = stitch_predictions(np.array(names), np.array(ranges), np.stack(preds))
    for name, (rr, pp) in per_video.items():
        assert np.array_equal(got[name], O.stitch(rr, pp))


@settings(max_examples=60, deadline=None)
@given(a=st.integers(8, 400), b=st.integers(4, 300), f=st.sampled_from(["lanczos", "bilinear"]))
def test_tap_tables(a, b, f):
    """Pillow's resampling coefficient tables (fixed-point taps, bounds) for arbitrary size pairs."""
    from utils.pil_tables import resample_table
    t, bounds, kk = resample_table(a, b, f)
    t2, bounds2, kk2 = P.precompute_coeffs(a, b, f)
    assert t == t2 and np.array_equal(bounds, bounds2) and np.array_equal(kk, kk2)
    assert (bounds[:, 0] >= 0).all() and (bounds.sum(1) <= a).all() and (bounds[:, 1] <= t).all()


@settings(max_examples=40, deadline=None)
@given(clips=st.integers(1, 4), frames=st.integers(1, 40), num_phase=st.sampled_from([4, 8, 12]))
def test_clip_window_index_selects_the_sampler_windows(clips, frames, num_phase):
    from sampler.snippet_sampler import window_index
    idx = window_index(0, frames, frames, num_phase)
    idx = (idx[None] + (torch.arange(clips) * frames)[:, None, None]).reshape(clips * frames, num_phase + 1)
    x = torch.arange(clips * frames, dtype=torch.float32)[:, None, None].expand(clips * frames, 2, 2).contiguous()
    want = torch.stack([O.gather_windows(x[b * frames:(b + 1) * frames], 0, frames, num_phase) for b in range(clips)])
    assert torch.equal(x[idx], want.reshape(clips * frames, num_phase + 1, 2, 2))


@settings(max_examples=60, deadline=None)
@given(h=st.integers(2, 13), w=st.integers(2, 13), k=st.sampled_from([3, 5, 7]), seed=st.integers(0, 2 ** 16))
def test_sub_lattice_tap_addressing(h, w, k, seed):
    """The addressing scheme of the stride-2 k x k layers (conv_engine.cu, conv_forward): input pixel 2y + kh - pad lies on the
    sub-lattice of parity (kh - pad) & 1 at index y + ((kh - pad) >> 1); indices outside a sub-lattice view read zero (TMA
    out-of-bounds fill = the convolution's padding).  Emulated with numpy slicing against torch's strided convolution."""
    import torch.nn.functional as F
    pad = k // 2
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 1, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(1, 1, k, k, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, wt, stride=2, padding=pad)[0, 0].numpy()
    ho, wo = ref.shape
    xn = x[0, 0].numpy()
    views = {(ph, pw): xn[ph::2, pw::2] for ph in (0, 1) for pw in (0, 1)}     # what the four tensor maps describe
    for (ph, pw), v in views.items():
        assert v.shape == ((h - ph + 1) // 2, (w - pw + 1) // 2)               # the extents conv_forward encodes
    out = np.zeros((ho, wo))
    for kh in range(k):
        for kw in range(k):
            oh, ow = kh - pad, kw - pad
            v = views[(oh & 1, ow & 1)]
            for y in range(ho):
                for xx in range(wo):
                    iy, ix = y + (oh >> 1), xx + (ow >> 1)                      # box coordinate of this output pixel in the view
                    if 0 <= iy < v.shape[0] and 0 <= ix < v.shape[1]:
                        out[y, xx] += wt[0, 0, kh, kw].item() * v[iy, ix]
    assert np.abs(out - ref).max() < 1e-12
