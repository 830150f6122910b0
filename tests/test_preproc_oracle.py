"""CPU: the PIL-preprocessing oracle against fixtures made by the real PIL / torchvision pipeline
(oracle/make_golden_preproc.py), against Pillow itself when importable, and the product-side tap
tables against the oracle's."""
import os

import numpy as np
import pytest
import torch

from oracle import mimamo_oracle as O
from oracle import pil_preproc as P


@pytest.fixture(scope="module")
def fixture(golden_dir):
    return np.load(os.path.join(golden_dir, "preproc_pil.npz"))


def test_gray_matches_pil_fixture(fixture):
    got = P.crops_to_gray(fixture["crops"])
    assert got.dtype == np.float32 and np.array_equal(got, fixture["gray"])


def test_rgb_matches_torchvision_fixture(fixture):
    got = P.crops_to_rgb(fixture["crops"])
    assert got.shape == (6, 3, 224, 224) and got.dtype == np.float32
    assert np.array_equal(got[:, :, :8, :], fixture["rgb_f32_rows"])
    # the whole image through PIL's uint8 result and the (u8/255)*255 - mean arithmetic pinned above
    u8 = fixture["rgb_u8"].astype(np.float32)
    full = (u8 / np.float32(255)) * np.float32(255) - np.asarray(P.RESNET_MEAN, np.float32).reshape(1, 3, 1, 1)
    assert np.array_equal(got, full)


def test_against_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for size, out, filt, pil_filter in [(112, 48, "lanczos", Image.LANCZOS), (112, 256, "bilinear", Image.BILINEAR),
                                        (97, 48, "lanczos", Image.LANCZOS), (64, 200, "bilinear", Image.BILINEAR)]:
        img = rng.integers(0, 256, (size, size), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img, "L").resize((out, out), pil_filter))
        assert np.array_equal(P.resize_u8(img, out, out, filt), want), (size, out, filt)
    rgb = rng.integers(0, 256, (40, 40, 3), dtype=np.uint8)
    assert np.array_equal(P.rgb_to_l(rgb), np.asarray(Image.fromarray(rgb, "RGB").convert("L")))


@pytest.mark.parametrize("a,b,f", [(112, 48, "lanczos"), (112, 256, "bilinear"), (100, 37, "lanczos"), (300, 48, "lanczos"),
                                   (64, 256, "bilinear")])
def test_product_tap_tables_match_oracle(a, b, f):
    from utils.pil_tables import resample_table
    t, bounds, kk = resample_table(a, b, f)
    t2, bounds2, kk2 = P.precompute_coeffs(a, b, f)
    assert t == t2 and np.array_equal(bounds, bounds2) and np.array_equal(kk, kk2)
    assert (bounds[:, 0] >= 0).all() and (bounds.sum(1) <= a).all() and (bounds[:, 1] <= t).all()


def test_clip_window_index_matches_oracle_gather():
    """The index the crop path hands to mimamo_pyr_phase_indexed selects exactly the windows the
    reference's sampler stacks (snippet_sampler.py:144-152), each clip being its own video."""
    from sampler.snippet_sampler import window_index
    B, F = 3, 20
    idx = window_index(0, F, F, 12)
    idx = (idx[None] + (torch.arange(B) * F)[:, None, None]).reshape(B * F, 13)
    frames = torch.arange(B * F, dtype=torch.float32)[:, None, None].expand(B * F, 2, 2).contiguous()
    want = torch.stack([O.gather_windows(frames[b * F:(b + 1) * F], 0, F) for b in range(B)]).reshape(B * F, 13, 2, 2)
    assert torch.equal(frames[idx], want)
