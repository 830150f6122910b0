"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through ref_shim).

Run in the build container (needs /root/reference):  python oracle/make_golden.py
It also asserts that the oracle restatement reproduces the reference on every vector it
writes, which is what "pins" the oracle.  The GPU box never runs this script; it only reads
the committed fixtures.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mimamo_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pde_case(R, name, shape, height, nbands, levels, seed, keep_coeff=None):
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(seed))
    pde = R.Phase_Difference_Extractor(height=height, nbands=nbands, extract_level=levels)
    coeff = pde.build_pyramid(x)
    diffs = [pde.extract(c) for c in coeff]
    mine_c = O.build_pyramid(x, height, nbands, levels)
    mine_d = [O.extract(c) for c in mine_c]
    for a, b in zip(coeff + diffs, mine_c + mine_d):
        assert torch.equal(a, b), "oracle restatement drifted from the reference (%s)" % name
    out = {"x": x.numpy(), "height": height, "nbands": nbands, "levels": np.array(levels)}
    for i, (c, d) in enumerate(zip(coeff, diffs)):
        if keep_coeff is None or i in keep_coeff:
            out["coeff%d" % i] = c.numpy()
        out["diff%d" % i] = d.numpy()
    np.savez(os.path.join(GOLD, name + ".npz"), **out)
    print(name, [tuple(d.shape) for d in diffs])


def head_case(R, name, bs, nf, seed):
    g = torch.Generator().manual_seed(seed)
    p0 = torch.randn(bs, nf, 24, 48, 48, generator=g)
    p1 = torch.randn(bs, nf, 24, 24, 24, generator=g)
    rgb = torch.rand(bs, nf, 2048, generator=g) * 4
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    model = R.Two_Stream_RNN().eval()
    model.load_state_dict(sd)                        # strict: key set must match the reference
    with torch.no_grad():
        y = model([p0, p1], rgb)
        mine = O.head_forward(sd, p0, p1, rgb)
    assert (y - mine).abs().max() < 5e-6
    chk = float(p0.double().sum() + p1.double().sum() + rgb.double().sum())
    wchk = float(sum(v.double().sum() for v in sd.values()))
    np.savez(os.path.join(GOLD, name + ".npz"), y=y.numpy(), input_checksum=chk, weight_checksum=wchk,
             bs=bs, nf=nf, seed=seed)
    print(name, tuple(y.shape), chk, wchk)


def scf_case(R, name, size, height, nbands, seed):
    """SCFpyr_PyTorch.build / reconstruct of un-mirrored images (api/steerable/SCFpyr_PyTorch.py:70-125,214-245)."""
    x = torch.rand(2, 1, size, size, generator=torch.Generator().manual_seed(seed))
    pyr = R.SCFpyr_PyTorch(height=height, nbands=nbands, scale_factor=2, device=torch.device("cpu"))
    coeff = pyr.build(x)
    rec = pyr.reconstruct(coeff)
    mine = O.pyramid_build(x, height, nbands)
    out = {"x": x.numpy(), "height": height, "nbands": nbands, "rec": rec.numpy()}
    for i, (a, b) in enumerate(zip(coeff, mine)):
        if isinstance(a, list):
            for j, (aa, bb) in enumerate(zip(a, b)):
                assert torch.equal(aa, bb), "oracle pyramid_build drifted from the reference (%s)" % name
            out["c%d" % i] = torch.stack(a, 0).numpy()
        else:
            assert torch.equal(a, b), "oracle pyramid_build drifted from the reference (%s)" % name
            out["c%d" % i] = a.numpy()
    mine_rec = O.pyramid_reconstruct(coeff, nbands)
    assert (mine_rec - rec).abs().max() < 1e-6, "oracle pyramid_reconstruct drifted from the reference (%s)" % name
    print(name, "levels", len(coeff), "round trip max|err| %.2e" % (rec - x[:, 0]).abs().max().item(),
          "oracle reconstruct vs reference %.1e" % (mine_rec - rec).abs().max().item())
    np.savez(os.path.join(GOLD, name + ".npz"), **out)


def extract_phase_case(R, name, seed):
    """Steerable_Pyramid_Phase.extract_phase(return_phase / return_both), Aff-wild-exps/utils.py:367-432."""
    U = ref_shim.load_training_utils()
    x = torch.rand(2, 5, 32, 32, generator=torch.Generator().manual_seed(seed))
    spp = U.Steerable_Pyramid_Phase(height=4, nbands=2, scale_factor=2, device=torch.device("cpu"), extract_level=2)
    coeff = spp.build_pyramid(x)
    outs = {"diff": spp.extract_phase(coeff), "phase": spp.extract_phase(coeff, return_phase=True),
            "both": spp.extract_phase(coeff, return_both=True)}
    mine = {"diff": O.extract_phase(coeff), "phase": O.extract_phase(coeff, return_phase=True),
            "both": O.extract_phase(coeff, return_both=True)}
    for k in outs:
        assert torch.equal(outs[k], mine[k]), "oracle extract_phase(%s) drifted from the reference" % k
    np.savez(os.path.join(GOLD, name + ".npz"), coeff=coeff.numpy(), **{k: v.numpy() for k, v in outs.items()})
    print(name, {k: tuple(v.shape) for k, v in outs.items()})


def main():
    R = ref_shim.load()
    os.makedirs(GOLD, exist_ok=True)
    # BASELINE config 1: one 8-frame 112x112 stack, 4 orientations x 2 scales
    pde_case(R, "pde_cfg1", (1, 8, 112, 112), 4, 4, [1, 2], seed=0, keep_coeff=(1,))
    # Tester defaults (api/tester.py:27-33): 13-frame 48x48 windows, height 4, 2 bands
    pde_case(R, "pde_tester", (2, 13, 48, 48), 4, 2, [1, 2], seed=2)
    # odd crop sizes / three levels / int extract_level
    pde_case(R, "pde_odd", (1, 3, 44, 44), 4, 6, [2], seed=3)
    pde_case(R, "pde_3lvl", (1, 4, 64, 64), 5, 3, [1, 2, 3], seed=4)
    # head: GRU recurs over dim 0 (bs) -- bs=3 exercises the recurrence, bs=1 the degenerate case
    head_case(R, "head_b3", 3, 4, seed=5)
    head_case(R, "head_b1", 1, 8, seed=6)
    # full pyramid of un-mirrored images + reconstruction: two oriented levels; an odd level size (50 -> 25)
    scf_case(R, "scf_64", 64, 4, 2, seed=7)
    scf_case(R, "scf_50", 50, 3, 3, seed=8)
    # training-side tail variants
    extract_phase_case(R, "extract_phase", seed=9)
    # unwrap known answers straight from the reference function
    pu = R.phase_utils
    kat_in = torch.tensor([[2.0, -2.5, -2.6], [-2.0, 2.5, 2.6], [0.0, 3.0, -3.0], [3.0, -3.0, 3.0]])
    np.savez(os.path.join(GOLD, "unwrap_kat.npz"), x=kat_in.numpy(),
             y=pu.torch_unwrap(kat_in, dim=-1).numpy())


if __name__ == "__main__":
    main()
