"""GPU parity: steerable pyramid + phase tail through the C ABI vs the reference's outputs
(tests/golden, written by the unmodified reference) and vs the oracle.

Tolerances (BASELINE.json north_star): phase maps within 1e-4 abs (fp32).  The reference's own
fp32 arithmetic sits 2e-5..1e-4 from the fp64 truth, so the fp64 distance
is checked as well.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import mimamo_oracle as O

pytestmark = pytest.mark.gpu
PHASE_TOL = 1e-4
COEFF_TOL = 2e-6
COEFF_TOL_LARGE = 1e-5      # 224x224 frames, relative to max(1, |coeff|max) of the level: block products as 3xTF32 tensor-core MMAs
                            # (22-bit operands; measured 2.1e-6 at level 1, 5.8e-6 at level 3; fp32 FMA path 1e-6); the phase tolerance stays 1e-4


def _pde(height, nbands, levels):
    from phase_difference_extractor import Phase_Difference_Extractor
    return Phase_Difference_Extractor(height=height, nbands=nbands, extract_level=levels)


def _as_list(x):
    return x if isinstance(x, list) else [x]


@pytest.mark.parametrize("name", ["pde_cfg1", "pde_tester", "pde_odd", "pde_3lvl"])
def test_against_reference_fixtures(cuda, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = torch.from_numpy(g["x"])
    levels = [int(l) for l in g["levels"]]
    pde = _pde(int(g["height"]), int(g["nbands"]), levels)
    coeffs = _as_list(pde.build_pyramid(x.to(cuda)))
    fused = _as_list(pde.phase_difference(x.to(cuda)))
    truth = O.build_pyramid(x.double(), int(g["height"]), int(g["nbands"]), levels, dtype=torch.float64)
    for i, (c, f) in enumerate(zip(coeffs, fused)):
        ref_d = torch.from_numpy(g["diff%d" % i])
        if "coeff%d" % i in g:
            ref_c = torch.from_numpy(g["coeff%d" % i])
            assert c.shape == ref_c.shape
            err_c = (c.cpu() - ref_c).abs().max().item()
            err_t = (c.cpu().double() - truth[i]).abs().max().item()
            ref_t = (ref_c.double() - truth[i]).abs().max().item()
            print("%s level %d: coeff |gpu-ref| %.2e  |gpu-fp64| %.2e  |ref-fp64| %.2e" % (name, i, err_c, err_t, ref_t))
            assert err_c < COEFF_TOL
            # the tail alone, fed the reference's own coefficients
            d = pde.extract(ref_c.to(cuda)).cpu()
            assert (d - ref_d).abs().max().item() < 2e-5
        assert f.shape == ref_d.shape
        err = (f.cpu() - ref_d).abs()
        truth_d = O.extract(truth[i])
        err64 = (f.cpu().double() - truth_d).abs().max().item()
        ref64 = (ref_d.double() - truth_d).abs().max().item()
        print("%s level %d: phase |gpu-ref| max %.2e, frac>1e-4 %.2e; |gpu-fp64| %.2e, |ref-fp64| %.2e"
              % (name, i, err.max().item(), (err > PHASE_TOL).float().mean().item(), err64, ref64))
        assert err.max().item() < PHASE_TOL
        # unfused path (build_pyramid -> extract) is the same computation
        assert torch.equal(pde.extract(c), f)


def test_tester_shapes_and_channel_order(cuda):
    """Tester.phase_diff_output layout: channel = band*12 + t (api/tester.py:131-138)."""
    from tester import Tester
    x = torch.rand(2, 3, 13, 48, 48, generator=torch.Generator().manual_seed(11))
    pde = _pde(4, 2, [1, 2])
    p0, p1 = Tester.phase_diff_output(None, x.to(cuda), pde)
    assert p0.shape == (2, 3, 24, 48, 48) and p1.shape == (2, 3, 24, 24, 24)
    r0, r1 = O.phase_diff_output(x)
    assert (p0.cpu() - r0).abs().max() < PHASE_TOL and (p1.cpu() - r1).abs().max() < PHASE_TOL


def test_properties_at_bench_size(cuda):
    """Full bench-size batch (2048 windows): size-independent properties instead of an oracle run."""
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(2048, 13, 48, 48, generator=gen).to(cuda)
    pde = _pde(4, 2, [1, 2])
    d0, d1 = pde.phase_difference(x)
    assert d0.shape == (2048, 2, 12, 48, 48) and d1.shape == (2048, 2, 12, 24, 24)
    for d in (d0, d1):
        assert torch.isfinite(d).all()
        assert d.abs().max().item() <= 5 * math.pi + 1e-5                     # clamp
        clamped = (d.abs() >= 5 * math.pi - 1e-6).flatten(3).any(-1)
        m = d.flatten(3).mean(-1)
        assert m[~clamped].abs().max().item() < 1e-4                           # spatial mean removed
    # windows are independent: a permuted batch gives the permuted result, bit for bit
    perm = torch.randperm(2048, generator=gen).to(cuda)
    e0, e1 = pde.phase_difference(x[perm])
    assert torch.equal(e0, d0[perm]) and torch.equal(e1, d1[perm])
    # spot-check 4 windows against the oracle
    idx = [0, 777, 1500, 2047]
    r0, r1 = [O.extract(c) for c in O.build_pyramid(x[idx].cpu(), 4, 2, [1, 2])]
    assert (d0[idx].cpu() - r0).abs().max() < PHASE_TOL and (d1[idx].cpu() - r1).abs().max() < PHASE_TOL


def test_constant_and_static_inputs(cuda):
    pde = _pde(4, 2, [1, 2])
    frame = torch.rand(1, 1, 48, 48, generator=torch.Generator().manual_seed(1))
    static = frame.expand(1, 13, 48, 48).contiguous().to(cuda)
    for d in _as_list(pde.phase_difference(static)):
        assert d.abs().max().item() == 0.0                                     # no motion -> no phase change
    ref = [O.extract(c) for c in O.build_pyramid(static.cpu(), 4, 2, [1, 2])]
    assert all(r.abs().max().item() == 0.0 for r in ref)


def test_edge_shapes(cuda):
    pde = _pde(4, 2, [1, 2])
    empty = pde.phase_difference(torch.zeros(0, 13, 48, 48, device=cuda))
    assert empty[0].shape == (0, 2, 12, 48, 48) and empty[1].shape == (0, 2, 12, 24, 24)
    x = torch.rand(3, 2, 48, 48, generator=torch.Generator().manual_seed(2))       # T = 2: a single difference
    got = pde.phase_difference(x.to(cuda))
    ref = [O.extract(c) for c in O.build_pyramid(x, 4, 2, [1, 2])]
    for a, b in zip(got, ref):
        assert a.shape == b.shape and (a.cpu() - b).abs().max() < PHASE_TOL
    single = _pde(4, 2, 1)                                                         # int extract_level -> tensor
    c = single.build_pyramid(x.to(cuda))
    assert torch.is_tensor(c) and c.shape == (3, 2, 2, 48, 48, 2)
    with pytest.raises(AssertionError):
        pde.build_pyramid(x)                                                       # CPU tensor: device mismatch


def test_tiled_tail_matches_whole_map(cuda):
    """Maps above 56x56 are split into 32x32 tiles with a two-phase mean; same numbers expected."""
    x = torch.rand(1, 4, 112, 112, generator=torch.Generator().manual_seed(9))
    c = O.build_pyramid(x, 4, 4, [1])[0]
    ref = O.extract(c)
    got = _pde(4, 4, [1]).extract(c.to(cuda)).cpu()
    assert (got - ref).abs().max() < 2e-5


def test_frame_dedup_is_exact(cuda, monkeypatch):
    """Sliding windows (12 of 13 frames shared with the neighbour, clamped repeats at clip ends):
    the fused path builds the pyramid once per distinct frame.  Must be bit-identical to
    transforming every copy, and equal the oracle."""
    gen = torch.Generator().manual_seed(17)
    clips = torch.rand(3, 40, 48, 48, generator=gen)
    clips[1, 5] = clips[1, 4]                                   # a genuine repeated frame inside a clip
    gray = torch.cat([O.gather_windows(clips[b], 0, 40) for b in range(3)])       # (120, 13, 48, 48)
    pde = _pde(4, 2, [1, 2])
    dedup = pde.phase_difference(gray.to(cuda))
    monkeypatch.setenv("MIMAMO_PYR_DEDUP", "0")
    plain = pde.phase_difference(gray.to(cuda))
    for a, b in zip(dedup, plain):
        assert torch.equal(a, b)
    idx = [0, 3, 39, 40, 45, 46, 119]
    ref = [O.extract(c) for c in O.build_pyramid(gray[idx], 4, 2, [1, 2])]
    for a, r in zip(dedup, ref):
        assert (a[idx].cpu() - r).abs().max() < PHASE_TOL


def test_config3_large_frames(cuda):
    """BASELINE config 3 shape: 224x224 frames, height=5 (3 oriented scales), 8 orientations.
    Frames this large do not fit in shared memory; the persistent large-frame path is used."""
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(23))
    pde = _pde(5, 8, [1, 2, 3])
    got_c = pde.build_pyramid(x.to(cuda))
    got_d = pde.phase_difference(x.to(cuda))
    ref_c = O.build_pyramid(x, 5, 8, [1, 2, 3])
    for i, (c, d, rc) in enumerate(zip(got_c, got_d, ref_c)):
        assert c.shape == rc.shape == (2, 8, 3, 224 >> i, 224 >> i, 2)
        assert (c.cpu() - rc).abs().max() < COEFF_TOL_LARGE * max(1.0, rc.abs().max().item())
        rd = O.extract(rc)
        err = (d.cpu() - rd).abs()
        print("config3 level %d: phase max|err| %.2e" % (i, err.max().item()))
        assert err.max() < PHASE_TOL
    with pytest.raises(RuntimeError, match="Cannot build 7 levels, image too small."):
        _pde(7, 8, [1]).build_pyramid(x.to(cuda))


def test_large_frames_tensor_core_products(cuda, monkeypatch):
    """Large frames run their block products as 3xTF32 tensor-core MMAs (tcgen05 kind::tf32, hi/lo split operands); MIMAMO_PYR_TC=0
    keeps the fp32-FMA blocks.  Both must agree with the oracle within the coefficient tolerance and with each other."""
    x = torch.rand(1, 2, 224, 224, generator=torch.Generator().manual_seed(29))
    ref_c = O.build_pyramid(x, 5, 8, [1, 2, 3])
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MIMAMO_PYR_TC", mode)
        got[mode] = [c.cpu() for c in _pde(5, 8, [1, 2, 3]).build_pyramid(x.to(cuda))]
    for i, rc in enumerate(ref_c):
        e_tc, e_fma = (got["1"][i] - rc).abs().max().item(), (got["0"][i] - rc).abs().max().item()
        print("level %d: |coeff| max %.3f, 3xTF32 max|err| %.2e, fp32 FMA max|err| %.2e" % (i, rc.abs().max().item(), e_tc, e_fma))
        scale = max(1.0, rc.abs().max().item())
        assert e_tc < COEFF_TOL_LARGE * scale and e_fma < COEFF_TOL * scale
        assert (got["1"][i] - got["0"][i]).abs().max() < COEFF_TOL_LARGE * scale


@pytest.mark.parametrize("name", ["scf_64", "scf_50"])
def test_full_pyramid_build_and_reconstruct(cuda, golden_dir, name):
    """SCFpyr_PyTorch.build (hi0, every band at full size, lo) and reconstruct through mimamo_scf_build /
    mimamo_scf_reconstruct vs the unmodified reference's outputs, plus the round-trip property."""
    from steerable.SCFpyr_PyTorch import SCFpyr_PyTorch
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = torch.from_numpy(g["x"]).to(cuda)
    height, nbands = int(g["height"]), int(g["nbands"])
    pyr = SCFpyr_PyTorch(height=height, nbands=nbands, scale_factor=2, device=cuda)
    coeff = pyr.build(x)
    assert isinstance(coeff, list) and len(coeff) == height
    worst = 0.0
    truth = O.pyramid_build(x.cpu().double(), height, nbands, dtype=torch.float64)
    for i, c in enumerate(coeff):
        ref = torch.from_numpy(g["c%d" % i])
        got = torch.stack(c, 0).cpu() if isinstance(c, list) else c.cpu()
        t64 = torch.stack(truth[i], 0) if isinstance(truth[i], list) else truth[i]
        assert got.shape == ref.shape
        scale = max(1.0, ref.abs().max().item())            # the low residual carries the image mean times (S/s)^2
        e_ref = (got - ref).abs().max().item() / scale
        print("  unit %d: |gpu-ref| %.2e  |gpu-fp64| %.2e  |ref-fp64| %.2e  (relative to %.1f)" % (
            i, e_ref, (got.double() - t64).abs().max().item() / scale, (ref.double() - t64).abs().max().item() / scale, scale))
        worst = max(worst, e_ref)
    rec = pyr.reconstruct(coeff).cpu()
    err_rec = (rec - torch.from_numpy(g["rec"])).abs().max().item()
    err_rt = (rec - x[:, 0].cpu()).abs().max().item()
    # the reference's own coefficients through the device reconstruct
    ref_coeff = []
    for i in range(height):
        r = torch.from_numpy(g["c%d" % i]).to(cuda)
        ref_coeff.append([r[b] for b in range(r.shape[0])] if r.dim() == 5 else r)
    err_rec2 = (pyr.reconstruct(ref_coeff).cpu() - torch.from_numpy(g["rec"])).abs().max().item()
    print("%s: build |gpu-ref| %.2e, reconstruct |gpu-ref| %.2e (ref coeff in: %.2e), round trip %.2e" % (name, worst, err_rec, err_rec2, err_rt))
    assert worst < COEFF_TOL and err_rec < 1e-5 and err_rec2 < 1e-5 and err_rt < 2e-5
    with pytest.raises(Exception):
        pyr.reconstruct([coeff[0], coeff[1][:-1]] + coeff[2:])       # "Unmatched number of orientations"
    with pytest.raises(RuntimeError):
        SCFpyr_PyTorch(height=9, nbands=2, device=cuda).build(x)      # 'Cannot build 9 levels, image too small.'


def test_build_pyramid_without_symmetry(cuda):
    """symmetry=False: no mirror extension, no quadrant crop (api/phase_difference_extractor.py:38-45,76-86)."""
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    pde = _pde(4, 2, [1, 2])
    got = pde.build_pyramid(x.to(cuda), symmetry=False)
    ref = O.build_pyramid(x, 4, 2, [1, 2], symmetry=False)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and (a.cpu() - b).abs().max().item() < COEFF_TOL
    assert got[0].shape == (2, 2, 3, 64, 64, 2) and got[1].shape == (2, 2, 3, 32, 32, 2)
    one = _pde(4, 2, 2).build_pyramid(x.to(cuda), symmetry=False)      # int extract_level -> a tensor
    assert torch.equal(one, got[1])


def test_extract_phase_variants(cuda, golden_dir):
    """Steerable_Pyramid_Phase.extract_phase(return_phase / return_both) vs the reference's outputs,
    on whole-map tiles (16x16) and on the tiled path (64x64 maps, fixed-order two-phase means)."""
    from phase_difference_extractor import Steerable_Pyramid_Phase
    g = np.load(os.path.join(golden_dir, "extract_phase.npz"))
    spp = Steerable_Pyramid_Phase(height=4, nbands=2, scale_factor=2, device=cuda, extract_level=2)
    coeff = torch.from_numpy(g["coeff"]).to(cuda)
    for key, kw in (("diff", {}), ("phase", {"return_phase": True}), ("both", {"return_both": True})):
        got = spp.extract_phase(coeff, **kw).cpu()
        ref = torch.from_numpy(g[key])
        err = (got - ref).abs().max().item()
        print("extract_phase %s: max|err| %.2e" % (key, err))
        assert got.shape == ref.shape and err < PHASE_TOL
    both = spp.extract_phase(coeff, return_both=True)
    assert float(both[:, :, coeff.shape[2] - 1:].abs().max()) == 0.0
    # maps larger than one CTA tile
    big = torch.randn(1, 2, 4, 64, 64, 2, generator=torch.Generator().manual_seed(6))
    for kw in ({}, {"return_phase": True}, {"return_both": True}):
        got = spp.extract_phase(big.to(cuda), **kw).cpu()
        ref = O.extract_phase(big, **kw)
        assert got.shape == ref.shape and (got - ref).abs().max().item() < PHASE_TOL


def test_map_kernel_matches_tiled_kernel(cuda, monkeypatch):
    """The software-pipelined whole-map tail kernel and the tiled kernel (forced with MIMAMO_TAIL=tiled) run the same
    per-pixel operations in the same order: identical bits, for every output mode."""
    from phase_difference_extractor import Steerable_Pyramid_Phase
    spp = Steerable_Pyramid_Phase(height=4, nbands=2, scale_factor=2, device=cuda, extract_level=1)
    gen = torch.Generator().manual_seed(12)
    for shape in ((3, 2, 13, 48, 48, 2), (2, 3, 5, 25, 25, 2), (1, 2, 4, 56, 56, 2)):
        coeff = torch.randn(*shape, generator=gen).to(cuda)
        for kw in ({}, {"return_phase": True}, {"return_both": True}):
            monkeypatch.delenv("MIMAMO_TAIL", raising=False)
            a = spp.extract_phase(coeff, **kw)
            monkeypatch.setenv("MIMAMO_TAIL", "tiled")
            b = spp.extract_phase(coeff, **kw)
            assert torch.equal(a, b), (shape, kw)
            assert (a.cpu() - O.extract_phase(coeff.cpu(), **kw)).abs().max().item() < PHASE_TOL


def test_size_specialised_tail_is_bit_identical(cuda, monkeypatch):
    """The whole-map tail kernel is instantiated with compile-time extents for the Tester configuration's 48x48 and 24x24 maps
    (constant strides / trip counts instead of run-time index arithmetic).  MIMAMO_TAIL_GENERIC=1 forces the run-time-extent
    instantiation: same operations in the same order, so the same bits -- fp32 outputs in every mode and the fp16 NHWC feed."""
    from phase_difference_extractor import Phase_Difference_Extractor, Steerable_Pyramid_Phase
    from sampler.snippet_sampler import window_index
    spp = Steerable_Pyramid_Phase(height=4, nbands=2, scale_factor=2, device=cuda, extract_level=1)
    gen = torch.Generator().manual_seed(21)
    for shape in ((5, 2, 13, 48, 48, 2), (5, 2, 13, 24, 24, 2), (3, 2, 2, 48, 48, 2)):
        coeff = torch.randn(*shape, generator=gen).to(cuda)
        for kw in ({}, {"return_phase": True}, {"return_both": True}):
            monkeypatch.delenv("MIMAMO_TAIL_GENERIC", raising=False)
            a = spp.extract_phase(coeff, **kw)
            monkeypatch.setenv("MIMAMO_TAIL_GENERIC", "1")
            b = spp.extract_phase(coeff, **kw)
            assert torch.equal(a, b), (shape, kw)
            assert (a.cpu() - O.extract_phase(coeff.cpu(), **kw)).abs().max().item() < PHASE_TOL
    pde = Phase_Difference_Extractor(height=4, nbands=2, extract_level=[1, 2])
    frames = torch.rand(40, 48, 48, generator=gen).to(cuda)
    idx = window_index(0, 40, 40, 12).to(cuda, torch.int32)
    monkeypatch.delenv("MIMAMO_TAIL_GENERIC", raising=False)
    fast = pde.phasenet_operands(frames, idx)
    fast32 = pde.phase_difference_indexed(frames, idx)
    monkeypatch.setenv("MIMAMO_TAIL_GENERIC", "1")
    slow = pde.phasenet_operands(frames, idx)
    slow32 = pde.phase_difference_indexed(frames, idx)
    for x, y in zip(list(fast) + list(fast32), list(slow) + list(slow32)):
        assert x.shape == y.shape and torch.equal(x, y)
