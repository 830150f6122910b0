"""GPU parity: the tcgen05 implicit-GEMM convolution engine (TMA boxes, TMEM accumulators)
against a plain fp32 convolution of the same bf16-rounded operands."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, H, W, Cin, Cout, k, stride, pad, relu, residual
    (2, 56, 56, 64, 64, 1, 1, 0, 1, False),      # flat GEMM, one K block
    (2, 56, 56, 64, 256, 1, 1, 0, 0, False),     # BLOCK_N = 256
    (3, 28, 28, 512, 128, 1, 1, 0, 1, False),    # 8 K blocks, BLOCK_N = 128, M not a tile multiple
    (2, 56, 56, 64, 64, 3, 1, 1, 1, False),      # 3x3: shifted 4-D boxes, TMA zero fill = padding
    (2, 28, 28, 128, 128, 3, 1, 1, 1, False),
    (3, 14, 14, 256, 256, 3, 1, 1, 1, False),
    (5, 7, 7, 512, 512, 3, 1, 1, 1, False),      # two images per box
    (2, 56, 56, 256, 128, 1, 2, 0, 1, False),    # strided 1x1: dense boxes over a strided view of the tensor
    (2, 56, 56, 256, 512, 1, 2, 0, 0, False),
    (3, 7, 9, 256, 512, 1, 2, 0, 0, False),      # odd extents: the view ends on the last sampled pixel
    (5, 14, 14, 512, 256, 1, 2, 0, 1, False),    # even extents: sampled rows merge across images (126-row tiles)
    (40, 28, 28, 256, 512, 1, 2, 0, 0, False),   # ... with more tiles than one wave of CTA pairs
    (2, 48, 48, 64, 64, 3, 2, 1, 1, False),      # PhaseNet stride-2 3x3: every tap a dense box over one of four sub-lattice views
    (3, 12, 12, 256, 256, 3, 2, 1, 1, False),
    (40, 24, 24, 128, 128, 3, 2, 1, 1, False),   # ... more tiles than CTAs
    (3, 9, 7, 64, 128, 3, 2, 1, 0, False),       # odd extents: the odd sub-lattices are one row / column shorter
    (2, 14, 14, 64, 64, 5, 2, 2, 1, False),      # 5x5: tap offsets -2 .. 2
    (2, 14, 14, 256, 1024, 1, 1, 0, 1, True),    # residual add + ReLU epilogue
    (1, 7, 7, 512, 2048, 1, 1, 0, 1, True),
    (300, 7, 7, 64, 64, 3, 1, 1, 0, False),      # more tiles than SMs: persistent loop + TMEM double buffer
]


# MIMAMO_PAIR "<mode><force>": "0" single CTAs; "11" forces 2-CTA clusters with multicast weight boxes on every 256-wide
# layer; "21" forces the cta_group::2 UMMA kernel (shapes this small would not select them on their own)
def _cases_with_modes():
    out = []
    for c in CASES:
        for pair in ("0", "11", "21"):
            if pair != "0" and c[4] % 256 != 0:
                continue                                  # pair modes only exist for 256-wide tiles
            out.append(pytest.param(c, pair, "1", id="x".join(str(v) for v in c) + "-pair" + pair))
        if c[6] == 2:                                     # strided layers: also the element-strided boxes (MIMAMO_STRIDED_VIEW=0)
            out.append(pytest.param(c, "0", "0", id="x".join(str(v) for v in c) + "-elemstride"))
    return out


@pytest.mark.parametrize("case,pair,view", _cases_with_modes())
def test_conv_engine(cuda, case, pair, view, monkeypatch):
    import _native
    monkeypatch.setenv("MIMAMO_PAIR", pair)
    monkeypatch.setenv("MIMAMO_STRIDED_VIEW", view)
    B, H, W, Cin, Cout, k, s, p, relu, use_res = case
    gen = torch.Generator().manual_seed(sum(case[:8]))
    x = torch.randn(B, H, W, Cin, generator=gen).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5).to(torch.bfloat16).float()
    scale = (1 + 0.1 * torch.randn(Cout, generator=gen)).float()
    shift = (0.1 * torch.randn(Cout, generator=gen)).float()
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn(B, Ho, Wo, Cout, generator=gen).to(torch.bfloat16) if use_res else None
    xd = x.to(cuda)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), dtype=torch.bfloat16, device=cuda)
    resd = res.to(cuda) if use_res else None
    wn, sn, tn = (np.ascontiguousarray(t.numpy()) for t in (w, scale, shift))
    rc = _native.lib().mimamo_conv_bf16(_native.dptr(xd), B, H, W, Cin, _native.f32_host_ptr(wn), _native.f32_host_ptr(sn),
                                        _native.f32_host_ptr(tn), Cout, k, s, p, relu,
                                        _native.dptr(resd) if use_res else None, _native.dptr(out),
                                        _native.stream_ptr(cuda))
    _native.check(rc)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2).to(cuda), w.to(cuda), stride=s, padding=p)
    ref = ref * scale.to(cuda)[None, :, None, None] + shift.to(cuda)[None, :, None, None]
    ref = ref.permute(0, 2, 3, 1)
    if use_res:
        ref = ref + resd.float()
    if relu:
        ref = ref.clamp_min(0)
    got = out.float()
    assert torch.isfinite(got).all(), "unwritten / non-finite outputs"
    err = (got - ref).abs()
    tol = 2e-2 + 1e-2 * ref.abs()                 # bf16 output rounding (2^-8 relative) + fp32 sum order
    assert (err <= tol).all(), "max err %.4f at ref %.4f" % (err.max().item(), ref.flatten()[err.argmax()].item())


CHAIN_CASES = [
    # M, K1, N1, N2
    (2 * 56 * 56, 64, 256, 64),        # ResNet50 stage 2: one K block, two 128-wide sub-tiles per M tile, 64-wide chained layer
    (3 * 28 * 28 + 5, 128, 512, 128),  # stage 3, M not a multiple of 128 (TMA clips the last tile)
    (100, 64, 128, 64),                # a single partial M tile
    (100, 64, 256, 64),                # ... through the on-chip hand-over kernels (one CTA, one tile, TMA clips rows >= M)
    (60, 128, 512, 128),
    (300 * 128, 64, 256, 64),          # more M tiles than SMs: the software-pipelined sequence over several tiles per CTA
    (149 * 128, 128, 512, 128),        # 149 tiles: one CTA owns two M tiles, the others one
]


def _chain_cases():
    # rw = "1": on-chip hand-over where a variant exists (stage-2 shape: resident weights, stage-3 shape: streamed), "0": hand-over
    # through L2 -- only distinct for the two ResNet50 shapes that have an on-chip variant to switch off
    out = []
    for c in CHAIN_CASES:
        out.append(pytest.param(c, "1", id="x".join(str(v) for v in c) + "-1"))
        if c[1:] in ((64, 256, 64), (128, 512, 128)):
            out.append(pytest.param(c, "0", id="x".join(str(v) for v in c) + "-0"))
    return out


@pytest.mark.parametrize("case,rw", _chain_cases())
def test_conv_chain(cuda, case, rw, monkeypatch):
    """conv_chain_kernel (increase + residual + ReLU, then the next block's reduce + ReLU in one launch) against the two
    layers computed separately in fp32 on the same bf16-rounded operands (the chained layer reads the bf16-rounded
    output of the first, exactly like the layer-by-layer path)."""
    import _native
    M, K1, N1, N2 = case
    monkeypatch.setenv("MIMAMO_CHAIN_SMEM", rw)
    gen = torch.Generator().manual_seed(sum(case))
    x = torch.randn(M, K1, generator=gen).to(torch.bfloat16)
    res = torch.randn(M, N1, generator=gen).to(torch.bfloat16)
    w1 = (torch.randn(N1, K1, generator=gen) * (2.0 / K1) ** 0.5).to(torch.bfloat16).float()
    w2 = (torch.randn(N2, N1, generator=gen) * (2.0 / N1) ** 0.5).to(torch.bfloat16).float()
    s1, s2 = (1 + 0.1 * torch.randn(N1, generator=gen)).float(), (1 + 0.1 * torch.randn(N2, generator=gen)).float()
    t1, t2 = (0.1 * torch.randn(N1, generator=gen)).float(), (0.1 * torch.randn(N2, generator=gen)).float()
    xd, resd = x.to(cuda), res.to(cuda)
    out1 = torch.full((M, N1), float("nan"), dtype=torch.bfloat16, device=cuda)
    out2 = torch.full((M, N2), float("nan"), dtype=torch.bfloat16, device=cuda)
    arrs = [np.ascontiguousarray(t.numpy()) for t in (w1, s1, t1, w2, s2, t2)]
    ptr = [_native.f32_host_ptr(a) for a in arrs]
    rc = _native.lib().mimamo_conv_chain_bf16(_native.dptr(xd), M, K1, ptr[0], ptr[1], ptr[2], N1, _native.dptr(resd),
                                              ptr[3], ptr[4], ptr[5], N2, _native.dptr(out1), _native.dptr(out2),
                                              _native.stream_ptr(cuda))
    _native.check(rc)
    torch.cuda.synchronize()
    ref1 = (xd.float() @ w1.to(cuda).t()) * s1.to(cuda) + t1.to(cuda) + resd.float()
    ref1 = ref1.clamp_min(0)
    got1 = out1.float()
    assert torch.isfinite(got1).all() and torch.isfinite(out2.float()).all(), "unwritten / non-finite outputs"
    err1 = (got1 - ref1).abs()
    assert (err1 <= 2e-2 + 1e-2 * ref1.abs()).all(), "layer 1: max err %.4f" % err1.max().item()
    ref2 = ((got1 @ w2.to(cuda).t()) * s2.to(cuda) + t2.to(cuda)).clamp_min(0)       # from the kernel's own 16-bit rows
    err2 = (out2.float() - ref2).abs()
    assert (err2 <= 2e-2 + 1e-2 * ref2.abs()).all(), "chained layer: max err %.4f" % err2.max().item()
