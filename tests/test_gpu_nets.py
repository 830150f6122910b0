"""GPU parity: ResNet50 pool5, the two-stream head and the end-to-end clip path vs the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import mimamo_oracle as O

pytestmark = pytest.mark.gpu
VA_TOL = 1e-3            # BASELINE.json north_star: valence/arousal within 1e-3 abs (identical inputs)


def _rgb(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, 3, 224, 224), generator=g).float() - torch.tensor(O.RESNET_MEAN)[None, :, None, None]


# relative tolerances = 2x the error measured on B200 (fp16 5.3e-4, bf16 3.6e-3 of the feature scale); fp16 is the default
@pytest.mark.parametrize("dtype,rel_tol,conv1", [("fp16", 6e-4, "line"), ("bf16", 7.5e-3, "line"), ("fp16", 6e-4, "im2col"),
                                                 ("fp16", 6e-4, "line_unfused_pool")])
def test_resnet50_pool5(cuda, dtype, rel_tol, conv1, monkeypatch):
    """Row R: parity against the restated architecture with seeded synthetic weights (parity with
    the published checkpoint is unpinned: the third-party definition/weights are absent).
    16-bit activations over 53 layers: tolerance is relative to the feature scale."""
    from resnet50_extractor import Resnet50_Extractor
    monkeypatch.setenv("MIMAMO_RESNET_DTYPE", dtype)
    # conv1 lowerings: line kernel over the space-to-depth'ed input (default), explicit im2col GEMM (cross-check)
    monkeypatch.setenv("MIMAMO_CONV1", "im2col" if conv1 == "im2col" else "s2d")
    monkeypatch.setenv("MIMAMO_CONV1_POOL", "0" if conv1 == "line_unfused_pool" else "1")    # pool1 fused into conv1's epilogue (default) or separate
    net = O.resnet_synthetic(1)
    x = _rgb(5, 21)
    ref = O.resnet_pool5(net, x)
    ext = Resnet50_Extractor(model=net)
    got = ext.features(x.to(cuda)).cpu()
    assert got.shape == (5, 2048)
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    print("resnet50 %s: max|err| %.3e (%.2e of scale %.3f), mean rel %.2e" % (dtype, err.max().item(), err.max().item() / scale, scale, (err.mean() / ref.abs().mean()).item()))
    assert err.max().item() < rel_tol * scale
    assert torch.isfinite(got).all()                        # fp16 activations: nothing overflowed with 0-255-scale inputs
    vec = ext.get_vec(x[:1].to(cuda))                       # reference quirk: bs == 1 squeezes to (2048,)
    assert vec.shape == (2048,) and vec.device.type == "cpu"
    # batch composition must not matter (chunking / tile boundaries)
    again = ext.features(x[1:4].to(cuda)).cpu()
    assert torch.equal(again, got[1:4])


def test_weight_rounding_calibration(cuda, monkeypatch):
    """conv_layer_quantize: mean-compensated weight rounding vs plain round-to-nearest (MIMAMO_RESNET_CALIB 2 / 1 / 0),
    same fp16 kernels.  The compensated rounding removes the error component that survives pool5's spatial average."""
    from resnet50_extractor import Resnet50_Extractor
    import time
    net = O.resnet_synthetic(1)
    x = _rgb(5, 21)
    ref = O.resnet_pool5(net, x)
    scale = ref.abs().max().item()
    errs = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("MIMAMO_RESNET_CALIB", mode)
        t0 = time.time()
        ext = Resnet50_Extractor(model=net)
        got = ext.features(x.to(cuda)).cpu()
        errs[mode] = (got - ref).abs().max().item() / scale
        print("resnet50 fp16, weight rounding mode %s: max|err| %.2e of scale (create + run %.1f s)" % (mode, errs[mode], time.time() - t0))
    assert errs["2"] < 0.6 * errs["0"] and errs["1"] < errs["0"]


def test_chained_layers_are_bit_identical(cuda, monkeypatch):
    """Stages 2-3 run `_increase` (+ residual) and the next block's `_reduce` as one launch (conv_chain_kernel); the chained
    layer reads the same 16-bit rows the separate launch would read, so pool5 is bit-identical to the layer-by-layer pass."""
    from resnet50_extractor import Resnet50_Extractor
    net = O.resnet_synthetic(1)
    x = _rgb(5, 34).to(cuda)
    monkeypatch.setenv("MIMAMO_CHAIN", "1")
    chained = Resnet50_Extractor(model=net).features(x)
    monkeypatch.setenv("MIMAMO_CHAIN", "0")
    separate = Resnet50_Extractor(model=net).features(x)
    assert torch.isfinite(chained).all()
    assert torch.equal(chained, separate)


def test_fused_pool1_is_bit_identical(cuda, monkeypatch):
    """pool1 fused into conv1's epilogue takes the maximum of the same 16-bit values a separate pooling kernel reads."""
    from resnet50_extractor import Resnet50_Extractor
    net = O.resnet_synthetic(1)
    x = _rgb(3, 33).to(cuda)
    monkeypatch.setenv("MIMAMO_CONV1_POOL", "1")
    fused = Resnet50_Extractor(model=net).features(x)
    monkeypatch.setenv("MIMAMO_CONV1_POOL", "0")
    unfused = Resnet50_Extractor(model=net).features(x)
    assert torch.equal(fused, unfused)


@pytest.mark.parametrize("name", ["head_b3", "head_b1"])
def test_head_against_reference_fixture(cuda, golden_dir, name):
    from mimamo_net import Two_Stream_RNN
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    bs, nf = int(g["bs"]), int(g["nf"])
    p0 = torch.randn(bs, nf, 24, 48, 48, generator=gen)
    p1 = torch.randn(bs, nf, 24, 24, 24, generator=gen)
    rgb = torch.rand(bs, nf, 2048, generator=gen) * 4
    model = Two_Stream_RNN().eval()
    model.load_state_dict(O.synthetic_state_dict(O.head_state_dict_spec(), seed=1))
    out = model([p0.to(cuda), p1.to(cuda)], rgb.to(cuda)).cpu()
    err = (out - torch.from_numpy(g["y"])).abs().max().item()
    print("%s: valence/arousal max|err| %.3e" % (name, err))
    assert out.shape == (bs, nf, 2) and err < VA_TOL


def test_head_batch32_recurrence(cuda):
    """BASELINE config 2 grouping: one forward with B = 32 snippets (GRU sequence length 32)."""
    from mimamo_net import Two_Stream_RNN
    gen = torch.Generator().manual_seed(8)
    bs, nf = 32, 64
    p0 = torch.randn(bs, nf, 24, 48, 48, generator=gen)
    p1 = torch.randn(bs, nf, 24, 24, 24, generator=gen)
    rgb = torch.rand(bs, nf, 2048, generator=gen) * 4
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    model = Two_Stream_RNN().eval()
    model.load_state_dict(sd)
    out = model([p0.to(cuda), p1.to(cuda)], rgb.to(cuda)).cpu()
    with torch.no_grad():
        ref = O.head_forward(sd, p0, p1, rgb)
    err = (out - ref).abs().max().item()
    print("head B=32: max|err| %.3e" % err)
    assert err < VA_TOL
    # frames are independent GRU batch rows: permuting frames permutes outputs
    perm = torch.randperm(nf, generator=gen)
    out_p = model([p0[:, perm].to(cuda), p1[:, perm].to(cuda)], rgb[:, perm].to(cuda)).cpu()
    assert (out_p - out[:, perm]).abs().max().item() < 1e-5


def test_end_to_end_clip_path(cuda):
    """Gray windows + RGB frames -> valence/arousal through Tester.infer_clips vs the oracle chain: the whole
    chain (pyramid + phase, fp16 ResNet50, head) inside the north_star's 1e-3 valence/arousal budget."""
    from tester import Tester
    B, Fr = 2, 8
    gen = torch.Generator().manual_seed(13)
    clip = torch.rand(B, 40, 48, 48, generator=gen)
    gray = torch.stack([O.gather_windows(clip[b], 10, 10 + Fr) for b in range(B)])
    rgb = _rgb(B * Fr, 14)
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=B, resnet_model=net, head_state_dict=sd)
    out = t.infer_clips(gray.to(cuda), rgb.to(cuda)).cpu()
    p0, p1 = O.phase_diff_output(gray)
    with torch.no_grad():
        ref = O.head_forward(sd, p0, p1, O.resnet_pool5(net, rgb).view(B, Fr, 2048))
    err = (out - ref).abs().max().item()
    print("end-to-end (fp16 ResNet): valence/arousal max|err| %.3e" % err)
    assert out.shape == (B, Fr, 2) and err < VA_TOL
    host = t.infer_clips_host(gray.pin_memory(), rgb.pin_memory(), copy_chunk=5)     # ragged last chunk
    assert torch.equal(host, out)


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_end_to_end_batch32_grouping(cuda, dtype, monkeypatch):
    """BASELINE configs[1] grouping: ONE forward over B = 32 snippets (GRU sequence length 32, api/tester.py:84-93), whole
    chain vs the oracle.  Frames are independent GRU batch rows, so 8 frames per snippet keep the CPU oracle's ResNet50
    affordable without changing the recurrence.  fp16 (the default) must meet the 1e-3 budget; bf16 is measured and
    reported (it does not: that is why it is not the default)."""
    from tester import Tester
    monkeypatch.setenv("MIMAMO_RESNET_DTYPE", dtype)
    B, Fr = 32, 8
    gen = torch.Generator().manual_seed(29)
    clip = torch.rand(B, Fr + 12, 48, 48, generator=gen)
    gray = torch.stack([O.gather_windows(clip[b], 6, 6 + Fr) for b in range(B)])
    rgb = _rgb(B * Fr, 30)
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=B, resnet_model=net, head_state_dict=sd)
    out = t.infer_clips(gray.to(cuda), rgb.to(cuda)).cpu()
    p0, p1 = O.phase_diff_output(gray)
    with torch.no_grad():
        ref = O.head_forward(sd, p0, p1, O.resnet_pool5(net, rgb).view(B, Fr, 2048))
    err = (out - ref).abs()
    print("end-to-end B=32 grouping (%s ResNet): valence/arousal max|err| %.3e, mean %.3e" % (dtype, err.max().item(), err.mean().item()))
    assert out.shape == (B, Fr, 2)
    assert err.max().item() < (VA_TOL if dtype == "fp16" else 3e-2)


def test_mlp_and_phasenet_on_their_own(cuda):
    """MLP.forward / PhaseNet.forward (api/mimamo_net.py:22-26,79-95) as stand-alone modules, feature=True and the
    classifier output (feature=False), and a deeper MLP than the published one."""
    from mimamo_net import MLP, PhaseNet
    gen = torch.Generator().manual_seed(41)
    for hidden in ([2048, 256, 256], [512, 384, 320, 256]):
        mlp = MLP(hidden).eval()
        sd = O.synthetic_state_dict([(k, tuple(v.shape)) for k, v in mlp.state_dict().items()], seed=3)
        mlp.load_state_dict(sd)
        x = torch.rand(3, 5, hidden[0], generator=gen) * 2
        got = mlp(x.to(cuda)).cpu()
        ref = O.mlp_forward(sd, x.reshape(15, -1), prefix="mlp.").view(3, 5, 256)
        err = (got - ref).abs().max().item()
        print("MLP %s: max|err| %.2e" % (hidden, err))
        assert got.shape == (3, 5, 256) and err < 1e-4
    for feature in (True, False):
        net = PhaseNet(48, 24, hidden_units=[256, 256, 1], dropout=0.3, feature=feature).eval()
        sd = O.synthetic_state_dict([(k, tuple(v.shape)) for k, v in net.state_dict().items()], seed=4)
        net.load_state_dict(sd)
        p0 = torch.randn(2, 3, 24, 48, 48, generator=gen)
        p1 = torch.randn(2, 3, 24, 24, 24, generator=gen)
        got = net(p0.to(cuda), p1.to(cuda)).cpu()
        ref = O.phasenet_forward(sd, p0.reshape(6, 24, 48, 48), p1.reshape(6, 24, 24, 24), prefix="", feature=feature)
        err = (got - ref).abs().max().item()
        print("PhaseNet feature=%s: max|err| %.2e (scale %.2f)" % (feature, err, ref.abs().max().item()))
        assert got.shape == ref.shape and err < 2e-3 * max(1.0, ref.abs().max().item())
    for size in (96, 112):                                      # the four-block variants (api/mimamo_net.py:33-40)
        net = PhaseNet(size, 24, hidden_units=[256, 256, 1], dropout=0.3, feature=True).eval()
        sd = O.synthetic_state_dict([(k, tuple(v.shape)) for k, v in net.state_dict().items()], seed=5)
        net.load_state_dict(sd)
        p0 = torch.randn(1, 3, 24, size, size, generator=gen)
        p1 = torch.randn(1, 3, 24, size // 2, size // 2, generator=gen)
        got = net(p0.to(cuda), p1.to(cuda)).cpu()
        ref = O.phasenet_forward(sd, p0.reshape(3, 24, size, size), p1.reshape(3, 24, size // 2, size // 2), prefix="", feature=True)
        err = (got - ref).abs().max().item()
        print("PhaseNet(%d): max|err| %.2e (scale %.2f)" % (size, err, ref.abs().max().item()))
        assert got.shape == (3, 256) and err < 2e-3 * max(1.0, ref.abs().max().item())
    with pytest.raises(ValueError):
        PhaseNet(50, 24)                                        # "Incorrect input size"


def test_head_other_num_phase(cuda):
    """Two_Stream_RNN(num_phase=8): 16 phase channels per level (the Tester passes its num_phase through, api/tester.py:42)."""
    from mimamo_net import Two_Stream_RNN
    gen = torch.Generator().manual_seed(43)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(num_phase=8), seed=2)
    model = Two_Stream_RNN(num_phase=8).eval()
    model.load_state_dict(sd)
    p0 = torch.randn(3, 4, 16, 48, 48, generator=gen)
    p1 = torch.randn(3, 4, 16, 24, 24, generator=gen)
    rgb = torch.rand(3, 4, 2048, generator=gen) * 4
    out = model([p0.to(cuda), p1.to(cuda)], rgb.to(cuda)).cpu()
    with torch.no_grad():
        ref = O.head_forward(sd, p0, p1, rgb)
    err = (out - ref).abs().max().item()
    print("head num_phase=8: max|err| %.3e" % err)
    assert err < VA_TOL
