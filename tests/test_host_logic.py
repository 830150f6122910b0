"""CPU: host-side logic of the drop-in package (tables, index rules, error behaviour, state_dict
keys, C-ABI exports).  No GPU compute is called here."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest
import torch

import mimamo_b200
from oracle import mimamo_oracle as O


def _plan_tables():
    from steerable import plan_tables
    return plan_tables


@pytest.mark.parametrize("H,height,nb,levels", [(48, 4, 2, [1, 2]), (44, 4, 6, [2]), (64, 5, 3, [1, 2, 3]),
                                                 (33, 4, 3, [1, 2]), (112, 4, 4, [1, 2])])
def test_folded_tables_reproduce_oracle(H, height, nb, levels):
    """The DCT-domain tables the kernels consume, evaluated in NumPy, equal the oracle's FFT pyramid."""
    pt = _plan_tables()
    tb = pt.build_tables(H, height, nb, levels)
    x = torch.rand(1, 2, H, H, dtype=torch.float64, generator=torch.Generator().manual_seed(H))
    ref = O.build_pyramid(x, height, nb, levels, dtype=torch.float64)
    got = pt.emulate(tb, x[0].numpy())
    for lv, r, e in zip(tb.levels, ref, got):
        r = r[0].permute(1, 0, 2, 3, 4).numpy()
        assert r.shape == e.shape
        assert np.abs(r - e).max() < 2e-7          # float32 table rounding only


def test_table_errors_match_reference():
    pt = _plan_tables()
    with pytest.raises(RuntimeError, match="Cannot build 7 levels, image too small."):
        pt.build_tables(224, 7, 8, [1])
    pt.build_tables(224, 6, 2, [4])                 # 4 oriented scales is the maximum at 224
    with pytest.raises(RecursionError):
        pt.build_tables(48, 4, 1, [1])


def test_product_index_rules_match_oracle():
    from sampler.snippet_sampler import snippet_ranges, window_frame_ids
    for n in (1, 13, 50, 63, 64, 65, 127, 128, 129, 300, 1000):
        assert snippet_ranges(n) == O.snippet_ranges(n)
        for s in (16, 32):
            assert snippet_ranges(n, 64, s) == O.snippet_ranges(n, 64, s)
        for f in (0, n // 2, n - 1):
            assert window_frame_ids(f, n) == O.window_frame_ids(f, n)


def test_stitch_matches_oracle():
    from tester import stitch_predictions
    ranges = np.array(O.snippet_ranges(150))
    preds = np.random.RandomState(0).rand(len(ranges), 64, 2)
    names = np.array(["v"] * len(ranges))
    got = stitch_predictions(names, ranges, preds)["v"]
    assert np.array_equal(got, O.stitch(ranges, list(preds)))


def test_two_stream_rnn_state_dict_keys():
    from mimamo_net import Two_Stream_RNN
    m = Two_Stream_RNN()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == O.head_state_dict_spec()
    m.load_state_dict(O.synthetic_state_dict(O.head_state_dict_spec(), seed=1))     # strict
    assert sum(p.numel() for p in m.parameters()) == 2636425
    with pytest.raises(RuntimeError):
        m.train()([torch.zeros(1), torch.zeros(1)], torch.zeros(1, 1, 2048))


def test_video_processor_errors():
    from video_processor import Video_Processor
    with pytest.raises(ValueError, match="OpenFace_exe has to be string object and needs to exist."):
        Video_Processor()
    vp = Video_Processor(OpenFace_exe="/bin/true")
    with pytest.raises(ValueError):
        vp.process("/nonexistent/video.mp4")
    argv = vp.command("/tmp/a.mp4", "/tmp/out")
    assert argv[1] == "-f" and "-simalign" in argv and "-nobadaligned" in argv and "-nomask" in argv and "-q" in argv


def test_extractor_argument_errors_without_gpu():
    from phase_difference_extractor import Phase_Difference_Extractor
    pde = Phase_Difference_Extractor(height=4, nbands=2, extract_level=[1, 2])
    assert (pde.height, pde.nbands, pde.scale_factor, pde.extract_level, pde.visualize) == (4, 2, 2, [1, 2], False)
    with pytest.raises(ValueError):
        pde.build_pyramid(torch.zeros(3, 48, 48))                       # not 4-D
    with pytest.raises(AssertionError):
        pde.build_pyramid(torch.zeros(1, 2, 48, 48, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="Cannot build 7 levels, image too small."):
        Phase_Difference_Extractor(height=7, nbands=2).build_pyramid(torch.zeros(1, 2, 48, 48, device=pde.pyramid.device))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pde.build_pyramid(torch.zeros(1, 2, 48, 48))


def test_load_aligned_crops_order_and_layout(tmp_path):
    """Tester.test's fast route decodes <video>_aligned/frame_det_00_%06d.bmp in OpenFace frame order (host I/O only)."""
    Image = pytest.importorskip("PIL.Image")
    from tester import load_aligned_crops
    d = tmp_path / "clip_opface" / "clip_aligned"
    d.mkdir(parents=True)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (12, 16, 16, 3), dtype=np.uint8)
    for i in (11, 3, 0, 7, 1, 2, 10, 4, 5, 9, 6, 8):                     # written out of order
        Image.fromarray(frames[i], "RGB").save(str(d / ("frame_det_00_%06d.bmp" % (i + 1))))
    got = load_aligned_crops(str(tmp_path / "clip_opface"), "clip")
    assert got.dtype == torch.uint8 and tuple(got.shape) == (12, 16, 16, 3)
    assert np.array_equal(got.numpy(), frames)
    with pytest.raises(ValueError):
        load_aligned_crops(str(tmp_path / "clip_opface"), "other")


def test_crop_path_has_no_cpu_fallback():
    """The uint8 face-crop entry points must fail loudly without a GPU (no host re-implementation hides behind them)."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from utils.crop_preprocessor import Crop_Preprocessor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Crop_Preprocessor()
    from resnet50_extractor import Resnet50_Extractor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Resnet50_Extractor(model={})


def test_bench_inputs_are_the_baseline_config():
    """bench.py's workload is BASELINE.json configs[1]: 32 clips x 64 frames of 112x112x3 uint8 per GPU."""
    sys.path.insert(0, os.path.dirname(mimamo_b200.PACKAGE_DIR))
    import bench
    from bench_inputs import make_crops
    assert (bench.CLIPS, bench.FRAMES, bench.WINDOWS) == (32, 64, 2048)
    crops = make_crops(seed=1, clips=2, frames=3)
    assert crops.shape == (2, 3, 112, 112, 3) and crops.dtype == torch.uint8
    assert torch.equal(crops, make_crops(seed=1, clips=2, frames=3))      # seeded: every run sees the same bits


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(os.path.dirname(mimamo_b200.PACKAGE_DIR), "include", "mimamo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mimamo_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 22
    assert os.path.exists(mimamo_b200.LIB_PATH), "run python __graft_entry__.py first"
    lib = ctypes.CDLL(mimamo_b200.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert _native.lib().mimamo_abi_version() == 1


def test_reference_arm_contract():
    """`bench.py --impl reference`: rank 0 prints ONE JSON line with the arm's keys (the reference's own classes from
    baseline/_ref when staged, the oracle port otherwise), every other rank exits 0 without work or output."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert other.returncode == 0 and other.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "windows/s" and d["higher_is_better"] is True and d["steps"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
