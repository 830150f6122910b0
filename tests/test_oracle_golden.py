"""CPU: the oracle restatement against the fixtures written by the UNMODIFIED reference
(oracle/make_golden.py) and against the reference's documented known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import mimamo_oracle as O


@pytest.mark.parametrize("name", ["pde_cfg1", "pde_tester", "pde_odd", "pde_3lvl"])
def test_pyramid_and_phase_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = torch.from_numpy(g["x"])
    levels = [int(l) for l in g["levels"]]
    coeffs = O.build_pyramid(x, int(g["height"]), int(g["nbands"]), levels)
    for i, c in enumerate(coeffs):
        if "coeff%d" % i in g:
            assert torch.equal(c, torch.from_numpy(g["coeff%d" % i]))
        assert torch.equal(O.extract(c), torch.from_numpy(g["diff%d" % i]))


def test_config1_shapes(golden_dir):
    g = np.load(os.path.join(golden_dir, "pde_cfg1.npz"))
    assert g["diff0"].shape == (1, 4, 7, 112, 112) and g["diff1"].shape == (1, 4, 7, 56, 56)


def test_unwrap_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "unwrap_kat.npz"))
    out = O.unwrap_positive_jumps(torch.from_numpy(g["x"]), dim=-1)
    assert torch.equal(out, torch.from_numpy(g["y"]))
    # negative jumps are NOT unwrapped, positive ones are
    assert torch.allclose(out[0], torch.tensor([2.0, -2.5, -2.6]))
    assert torch.allclose(out[1], torch.tensor([-2.0, -3.7832, -3.6832]), atol=1e-4)


@pytest.mark.parametrize("name", ["head_b3", "head_b1"])
def test_head_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    bs, nf = int(g["bs"]), int(g["nf"])
    p0 = torch.randn(bs, nf, 24, 48, 48, generator=gen)
    p1 = torch.randn(bs, nf, 24, 24, 24, generator=gen)
    rgb = torch.rand(bs, nf, 2048, generator=gen) * 4
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    assert abs(float(p0.double().sum() + p1.double().sum() + rgb.double().sum()) - float(g["input_checksum"])) < 1e-6
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["weight_checksum"])) < 1e-6
    with torch.no_grad():
        y = O.head_forward(sd, p0, p1, rgb)
    assert (y - torch.from_numpy(g["y"])).abs().max() < 5e-6


def test_gru_recurs_over_batch_axis():
    """snippets are coupled, frames are independent."""
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(3, 2, 24, 48, 48, generator=gen)
    p1 = torch.randn(3, 2, 24, 24, 24, generator=gen)
    rgb = torch.rand(3, 2, 2048, generator=gen)
    with torch.no_grad():
        joint = O.head_forward(sd, p0, p1, rgb)
        single = torch.cat([O.head_forward(sd, p0[i:i + 1], p1[i:i + 1], rgb[i:i + 1]) for i in range(3)])
        perm = O.head_forward(sd, p0[:, [1, 0]], p1[:, [1, 0]], rgb[:, [1, 0]])
    assert (joint - single).abs().max() > 1e-3
    assert torch.allclose(perm, joint[:, [1, 0]], atol=1e-6)


def test_fp32_noise_floor_vs_fp64():
    """The reference's own fp32 arithmetic sits ~1e-5..1e-4 from fp64 on the phase maps."""
    x = torch.rand(2, 13, 48, 48, generator=torch.Generator().manual_seed(7))
    d32 = O.phase_diff_output(x[None], dtype=torch.float32)
    d64 = O.phase_diff_output(x[None].double(), dtype=torch.float64)
    for a, b in zip(d32, d64):
        assert (a.double() - b).abs().max() < 5e-4


def test_snippet_and_window_rules():
    assert O.snippet_ranges(300) == [[0, 64], [64, 128], [128, 192], [192, 256], [236, 300]]
    assert O.snippet_ranges(128) == [[0, 64], [64, 128]]
    assert O.snippet_ranges(50) == [[0, 50]]
    assert O.snippet_ranges(65) == [[0, 64], [1, 65]]
    assert O.window_frame_ids(0, 300) == [0] * 7 + [1, 2, 3, 4, 5, 6]
    assert O.window_frame_ids(299, 300) == [293, 294, 295, 296, 297, 298] + [299] * 7
    assert O.window_frame_ids(100, 300) == list(range(94, 107))


def test_stitch_tail_overwrites():
    r = O.snippet_ranges(70)
    preds = [np.full((64, 2), 1.0), np.full((64, 2), 2.0)]
    out = O.stitch(r, preds)
    assert out.shape == (70, 2) and (out[:6] == 1).all() and (out[6:] == 2).all()


def test_resnet_structure():
    net = O.FerPlusResNet50()
    convs = [m for n, m in net.named_modules() if isinstance(m, torch.nn.Conv2d) and n != "classifier"]
    assert len(convs) == 53
    assert net._modules.get("pool5_7x7_s1") is not None
    f = O.resnet_pool5(O.resnet_synthetic(1), torch.zeros(1, 3, 224, 224))
    assert f.shape == (1, 2048)


@pytest.mark.parametrize("name", ["scf_64", "scf_50"])
def test_full_pyramid_and_reconstruction_match_reference(golden_dir, name):
    """SCFpyr_PyTorch.build / reconstruct of un-mirrored images: the oracle restatement and the host tables the CUDA
    path consumes (plan_tables.full_pyramid_tables, evaluated in NumPy) against the unmodified reference's outputs."""
    from steerable import plan_tables
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    x = torch.from_numpy(g["x"])
    height, nbands = int(g["height"]), int(g["nbands"])
    coeff = O.pyramid_build(x, height, nbands)
    units = plan_tables.full_pyramid_tables(x.shape[-1], height, nbands)
    emu = plan_tables.emulate_full(units, g["x"][:, 0])
    assert len(coeff) == len(units) == height
    for i, c in enumerate(coeff):
        ref = g["c%d" % i]
        if isinstance(c, list):
            assert torch.equal(torch.stack(c, 0), torch.from_numpy(ref))
            e = np.stack(emu[i])
            assert np.abs(e.real - ref[..., 0]).max() < 2e-6 and np.abs(e.imag - ref[..., 1]).max() < 2e-6
        else:
            assert torch.equal(c, torch.from_numpy(ref))
            assert np.abs(emu[i] - ref).max() < 2e-6
    rec = O.pyramid_reconstruct(coeff, nbands)
    assert (rec - torch.from_numpy(g["rec"])).abs().max() < 1e-6
    assert (rec - x[:, 0]).abs().max() < 2e-5                       # the reference's own round-trip property
    emu_rec = plan_tables.emulate_reconstruct(units, emu, x.shape[-1])
    assert np.abs(emu_rec - g["rec"]).max() < 1e-5


def test_extract_phase_variants_match_reference(golden_dir):
    """Steerable_Pyramid_Phase.extract_phase (Aff-wild-exps/utils.py:367-432): default / return_phase / return_both."""
    g = np.load(os.path.join(golden_dir, "extract_phase.npz"))
    coeff = torch.from_numpy(g["coeff"])
    assert torch.equal(O.extract_phase(coeff), torch.from_numpy(g["diff"]))
    assert torch.equal(O.extract(coeff), torch.from_numpy(g["diff"]))          # same tail as the api/ extractor
    assert torch.equal(O.extract_phase(coeff, return_phase=True), torch.from_numpy(g["phase"]))
    both = O.extract_phase(coeff, return_both=True)
    assert torch.equal(both, torch.from_numpy(g["both"]))
    t = coeff.shape[2]
    assert both.shape[2] == 2 * (t - 1) and float(both[:, :, t - 1:].abs().max()) == 0.0   # insert_tensors fills only T-1 slots
