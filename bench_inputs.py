"""Synthetic inputs / weights for bench.py: seeded CPU generators so every
run sees the same bits; nothing here touches the oracle or the reference."""
import math

import torch

RESNET_MEAN = (131.0912, 103.8827, 91.4953)
STAGES = ((2, 3, 64, 256), (3, 4, 128, 512), (4, 6, 256, 1024), (5, 3, 512, 2048))


def make_inputs(seed, clips, frames, t, size):
    """gray windows (clips, frames, t, size, size) in [0,1) built with the clamp-window rule
    (api/sampler/snippet_sampler.py:144-152) over a (frames)-long clip, and RGB (clips*frames,3,224,224)
    = uint8 value minus the channel mean.  Returned as pinned host tensors."""
    g = torch.Generator().manual_seed(seed)
    clip = torch.rand(clips, frames, size, size, generator=g)
    half = (t - 1) // 2
    idx = (torch.arange(frames)[:, None] + torch.arange(-half, half + 1)[None, :]).clamp_(0, frames - 1)
    gray = clip[:, idx].contiguous()                                   # (clips, frames, t, size, size)
    rgb = torch.randint(0, 256, (clips * frames, 3, 224, 224), generator=g, dtype=torch.uint8).float()
    rgb -= torch.tensor(RESNET_MEAN)[None, :, None, None]
    if torch.cuda.is_available():
        gray, rgb = gray.pin_memory(), rgb.pin_memory()
    return gray, rgb


def make_crops(seed, clips, frames, size=112):
    """uint8 face crops (clips, frames, size, size, 3), uniform 0-255: what OpenFace's aligned bmp files
    decode to (BASELINE.json configs[1]: synthetic 64-frame 112x112 clips).  Pinned host tensor."""
    g = torch.Generator().manual_seed(seed)
    crops = torch.randint(0, 256, (clips, frames, size, size, 3), generator=g, dtype=torch.uint8)
    return crops.pin_memory() if torch.cuda.is_available() else crops


def _fill(spec, seed, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in spec:
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(1, dtype=torch.int64)
        elif name.endswith("running_var"):
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1 and name.endswith(".weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[name] = torch.randn(shape, generator=g) * (gain * math.sqrt(2.0 / fan_in))
    return sd


def resnet_spec():
    spec = []

    def conv_bn(name, cout, cin, k):
        spec.append((name + ".weight", (cout, cin, k, k)))
        for s in ("weight", "bias", "running_mean", "running_var"):
            spec.append((name + "_bn." + s, (cout,)))

    conv_bn("conv1_7x7_s2", 64, 3, 7)
    cin = 64
    for stage, blocks, mid, cout in STAGES:
        for b in range(1, blocks + 1):
            p = "conv%d_%d_" % (stage, b)
            conv_bn(p + "1x1_reduce", mid, cin, 1)
            conv_bn(p + "3x3", mid, mid, 3)
            conv_bn(p + "1x1_increase", cout, mid, 1)
            if b == 1:
                conv_bn(p + "1x1_proj", cout, cin, 1)
            cin = cout
    return spec


def head_spec(num_phase=12):
    import mimamo_b200
    mimamo_b200.install()
    from mimamo_net import Two_Stream_RNN
    return [(k, tuple(v.shape)) for k, v in Two_Stream_RNN(num_phase=num_phase).state_dict().items()]


def synthetic_weights():
    rs = _fill(resnet_spec(), seed=1)
    for k in rs:
        if k.endswith("1x1_increase_bn.weight"):
            rs[k] = rs[k] * 0.25
    rs["conv1_7x7_s2.weight"] = rs["conv1_7x7_s2.weight"] * 0.02
    return rs, _fill(head_spec(), seed=2)
