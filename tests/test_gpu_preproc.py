"""GPU parity of the uint8 face-crop path: on-device PIL-exact preprocessing, the indexed
pyramid/phase entry and Tester.infer_crops, all through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import mimamo_oracle as O
from oracle import pil_preproc as P

pytestmark = pytest.mark.gpu


def _crops(n, seed, size=112):
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 256, (n, size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(0, n, 3):                                  # every third crop smooth, with a drifting pattern
        c[i] = np.stack([(127 + 110 * np.sin(xx / 9.0 + 0.3 * i + ch) * np.cos(yy / 13.0)).astype(np.uint8) for ch in range(3)], -1)
    return c


@pytest.mark.parametrize("generic", ["0", "1"])          # 3-tap word-wise RGB kernel / generic any-tap kernel
def test_preprocessing_is_bit_exact_with_pil(cuda, golden_dir, generic, monkeypatch):
    from utils.crop_preprocessor import Crop_Preprocessor
    monkeypatch.setenv("MIMAMO_PREPROC_GENERIC", generic)
    g = np.load(os.path.join(golden_dir, "preproc_pil.npz"))
    pre = Crop_Preprocessor()
    crops = torch.from_numpy(g["crops"]).to(cuda)
    gray = pre.gray(crops).cpu().numpy()
    rgb = pre.rgb(crops).cpu().numpy()
    assert np.array_equal(gray, g["gray"])                    # produced by PIL itself
    assert np.array_equal(rgb[:, :, :8, :], g["rgb_f32_rows"])  # produced by torchvision itself
    assert np.array_equal(rgb, P.crops_to_rgb(g["crops"]))
    more = _crops(9, 5)
    assert np.array_equal(pre.gray(torch.from_numpy(more).to(cuda)).cpu().numpy(), P.crops_to_gray(more))
    assert np.array_equal(pre.rgb(torch.from_numpy(more).to(cuda)).cpu().numpy(), P.crops_to_rgb(more))
    assert pre.gray(crops[:0]).shape == (0, 48, 48)
    with pytest.raises(ValueError):
        pre.gray(torch.zeros(2, 100, 112, 3, dtype=torch.uint8, device=cuda))


def test_other_crop_geometry(cuda):
    """save_size is a Tester argument (api/tester.py:18): a 96x96 crop must resample with its own tables."""
    from utils.crop_preprocessor import Crop_Preprocessor
    pre = Crop_Preprocessor(save_size=96, phase_size=32)
    c = _crops(3, 8, size=96)
    assert np.array_equal(pre.gray(torch.from_numpy(c).to(cuda)).cpu().numpy(), P.crops_to_gray(c, 32))
    assert np.array_equal(pre.rgb(torch.from_numpy(c).to(cuda)).cpu().numpy(), P.crops_to_rgb(c))


def test_indexed_pyramid_matches_window_path(cuda):
    from phase_difference_extractor import Phase_Difference_Extractor
    from sampler.snippet_sampler import window_index
    pde = Phase_Difference_Extractor(height=4, nbands=2, extract_level=[1, 2])
    gen = torch.Generator().manual_seed(3)
    frames = torch.rand(30, 48, 48, generator=gen).to(cuda)
    idx = window_index(0, 30, 30, 12).to(cuda, torch.int32)
    a = pde.phase_difference_indexed(frames, idx)
    b = pde.phase_difference(frames[idx.long()])
    for x, y in zip(a, b):
        assert x.shape == y.shape and torch.equal(x, y)
    ref0, ref1 = O.phase_diff_output(frames.cpu()[idx.long().cpu()][None])
    err = max((a[0].cpu().reshape(ref0.shape) - ref0).abs().max().item(), (a[1].cpu().reshape(ref1.shape) - ref1).abs().max().item())
    print("indexed pyramid+phase vs oracle: max|err| %.3e" % err)
    assert err < 1e-4                                         # north_star: phase maps within 1e-4 abs


def test_infer_crops_matches_clip_path_and_oracle(cuda):
    from tester import Tester
    B, Fr = 2, 16
    crops = _crops(B * Fr, 11).reshape(B, Fr, 112, 112, 3)
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=B, resnet_model=net, head_state_dict=sd)
    out = t.infer_crops(torch.from_numpy(crops).to(cuda)).cpu()
    # the same inputs prepared on the host exactly like the reference's samplers do
    gray = torch.from_numpy(P.crops_to_gray(crops.reshape(-1, 112, 112, 3))).view(B, Fr, 48, 48)
    windows = torch.stack([O.gather_windows(gray[b], 0, Fr) for b in range(B)])
    rgb = torch.from_numpy(P.crops_to_rgb(crops.reshape(-1, 112, 112, 3)))
    via_clips = t.infer_clips(windows.to(cuda), rgb.to(cuda)).cpu()
    assert torch.equal(out, via_clips)                        # same bits: only where the preprocessing ran differs
    p0, p1 = O.phase_diff_output(windows)
    with torch.no_grad():
        ref = O.head_forward(sd, p0, p1, O.resnet_pool5(net, rgb).view(B, Fr, 2048))
    err = (out - ref).abs().max().item()
    print("crop path end-to-end (fp16 ResNet): valence/arousal max|err| %.3e" % err)
    assert out.shape == (B, Fr, 2) and err < 1e-3             # north_star: valence/arousal within 1e-3 abs
    host = t.infer_crops_host(torch.from_numpy(crops).pin_memory())
    assert torch.equal(host, out)


# 2 snippets + overlapping tail in two batches; shorter than a snippet; shorter than one 13-frame window; a single frame
@pytest.mark.parametrize("n_frames,batch_size", [(150, 2), (40, 4), (9, 4), (1, 1)])
def test_video_path_matches_reference_semantics(cuda, n_frames, batch_size):
    """Tester.predict_frames / test_frames: windows clamped to the video, 64-frame snippets + tail snippet, the snippets of
    a DataLoader batch forwarded together (GRU over the batch), tail overwrites -- api/tester.py:53-121."""
    from tester import Tester
    crops = _crops(n_frames, 17)
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=batch_size, resnet_model=net, head_state_dict=sd)
    got = t.predict_frames(torch.from_numpy(crops).to(cuda)).cpu()
    assert got.shape == (n_frames, 2)
    # the reference's procedure restated with the oracle pieces and the device entry points it is built from
    gray = torch.from_numpy(P.crops_to_gray(crops))
    rgb = torch.from_numpy(P.crops_to_rgb(crops))
    ranges = O.snippet_ranges(n_frames)
    exact, ref = [], []
    for b0 in range(0, len(ranges), batch_size):
        batch = ranges[b0:b0 + batch_size]
        windows = torch.stack([O.gather_windows(gray, s, e) for s, e in batch])
        frames_rgb = torch.cat([rgb[s:e] for s, e in batch])
        dev = t.infer_clips(windows.to(cuda), frames_rgb.to(cuda)).cpu()
        exact += [dev[k].numpy() for k in range(len(batch))]
        p0, p1 = O.phase_diff_output(windows)
        with torch.no_grad():
            cpu = O.head_forward(sd, p0, p1, O.resnet_pool5(net, frames_rgb).view(len(batch), -1, 2048))
        ref += [cpu[k].numpy() for k in range(len(batch))]
    assert np.array_equal(got.numpy(), O.stitch(ranges, exact).astype(np.float32))      # same kernels, same bits
    err = np.abs(got.numpy() - O.stitch(ranges, ref)).max()
    print("video path (%d frames, %d snippets): valence/arousal max|err| %.3e" % (n_frames, len(ranges), err))
    assert err < 1e-3                                         # north_star: valence/arousal within 1e-3 abs
    frame = t.test_frames(crops, "clip_a")
    assert list(frame) == ["clip_a"] and list(frame["clip_a"].columns) == ["valence", "arousal"]
    assert np.array_equal(frame["clip_a"].to_numpy().astype(np.float32), got.numpy())
    from multi_gpu import run_videos                       # world size 1: the local block is everything
    n2 = max(1, n_frames - 7)
    both = run_videos(t, [crops, crops[:n2], crops, crops])  # the two trailing videos share a head forward when one batch holds their snippets
    assert len(both) == 4 and both[1].shape == (n2, 2)
    assert torch.equal(both[0].cpu(), got) and torch.equal(both[2].cpu(), got) and torch.equal(both[3].cpu(), got)
    with pytest.raises(ValueError):
        t.predict_frames(torch.zeros(0, 112, 112, 3, dtype=torch.uint8, device=cuda))


def test_fp16_operand_route_is_bit_identical(cuda, monkeypatch):
    """Tester feeds PhaseNet the phase tail's fp16 NHWC output directly (mimamo_pyr_phase_indexed_nhwc16 ->
    mimamo_head_forward_nhwc16); forcing the reference-shaped fp32 phase tensors (MIMAMO_PHASE_FP32=1: fp32 NCHW maps +
    the head's own transposition) must give the same predictions bit for bit, on every entry point."""
    from tester import Tester
    B, Fr = 3, 20
    crops = torch.from_numpy(_crops(B * Fr, 23).reshape(B, Fr, 112, 112, 3))
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=2, resnet_model=net, head_state_dict=sd)
    video = crops.reshape(B * Fr, 112, 112, 3)[:50]
    monkeypatch.delenv("MIMAMO_PHASE_FP32", raising=False)
    fast = (t.infer_crops(crops.to(cuda)), t.infer_crops_host(crops.pin_memory(), to_host=False), t.predict_frames(video.to(cuda)),
            t.predict_videos([video.to(cuda), video[:30].to(cuda)])[1])
    assert t._operand_buffers(B * Fr, cuda) is not None            # the operand route is the one that ran
    monkeypatch.setenv("MIMAMO_PHASE_FP32", "1")
    slow = (t.infer_crops(crops.to(cuda)), t.infer_crops_host(crops.pin_memory(), to_host=False), t.predict_frames(video.to(cuda)),
            t.predict_videos([video.to(cuda), video[:30].to(cuda)])[1])
    for a, b in zip(fast, slow):
        assert a.shape == b.shape and torch.equal(a, b)


def test_fast_and_file_routes_agree(cuda, tmp_path):
    """Tester.test_aligned on a synthetic OpenFace output directory: fast=True (crops decoded once, every transform on the
    device) against fast=False, the reference's file route (Resnet50_Extractor.run -> %05d.npy -> Snippet_Sampler with
    PIL -> DataLoader -> test_on_dataloader, api/tester.py:60-74).  Same numbers bit for bit, and the file route leaves the
    reference's feature cache behind."""
    from PIL import Image
    from tester import Tester
    n = 70                                                    # one full snippet + the overlapping tail snippet
    frames = _crops(n, 31)
    aligned = tmp_path / "clip_opface" / "clip_aligned"
    aligned.mkdir(parents=True)
    for i in range(n):
        Image.fromarray(frames[i]).save(str(aligned / ("frame_det_00_%06d.bmp" % (i + 1))))
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    t = Tester(None, batch_size=4, resnet_model=net, head_state_dict=sd)
    fast = t.test_aligned(str(tmp_path / "clip_opface"), "clip", fast=True)
    feature_dir = tmp_path / "clip_pool5"
    slow = t.test_aligned(str(tmp_path / "clip_opface"), "clip", str(feature_dir), fast=False)
    assert list(fast) == list(slow) == ["clip"]
    a, b = fast["clip"].to_numpy(), slow["clip"].to_numpy()
    assert a.shape == b.shape == (n, 2) and list(slow["clip"].columns) == ["valence", "arousal"]
    assert np.array_equal(a, b)
    assert len(list(feature_dir.glob("*.npy"))) == n and np.load(str(feature_dir / "00001.npy")).shape == (2048,)
