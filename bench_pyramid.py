"""BASELINE.json configs[2]: steerable pyramid + phase difference only, 224x224 frames, height 5 / 8 orientations /
extract levels [1,2,3], batch of 256 windows x 13 frames, 1 GPU -- the HBM-roofline run.

    python bench_pyramid.py [--windows 256] [--steps 3] [--warmup 1] [--size 224]

Prints one JSON line: windows/s with the frames resident in HBM, and the roofline object against the measured copy
bandwidth.  Algorithmic bytes per window = 4*T*H^2 + 4*nb*(T-1)*sum_l (H/2^(l-1))^2.
The main bench (bench.py) covers configs[1]; this secondary script only reports the pyramid stage at its own config.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    assert torch.cuda.is_available(), "needs a GPU (there is no CPU fallback)"
    import mimamo_b200
    mimamo_b200.install()
    from phase_difference_extractor import Phase_Difference_Extractor
    T, nb, H = 13, 8, args.size
    dev = torch.device("cuda", 0)
    pde = Phase_Difference_Extractor(height=5, nbands=nb, extract_level=[1, 2, 3])
    g = torch.Generator().manual_seed(0)
    frames = torch.rand(args.windows, T, H, H, generator=g).to(dev)          # distinct frames: nothing to de-duplicate
    for _ in range(args.warmup):
        outs = pde.phase_difference(frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        outs = pde.phase_difference(frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bytes_per_window = 4 * T * H * H + 4 * nb * (T - 1) * sum((H >> l) ** 2 for l in range(3))
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    gbs = bytes_per_window * args.windows / (ms / 1e3) / 1e9
    print(json.dumps({
        "metric": "face-windows/sec, steerable pyramid + phase difference only", "value": args.windows / (ms / 1e3),
        "unit": "windows/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: %d windows x %d frames of %dx%d, height 5, 8 orientations, levels [1,2,3]" % (args.windows, T, H, H),
                   "l2": "inputs + outputs (%.1f GB) exceed the 126 MB L2" % (bytes_per_window * args.windows / 1e9)},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                     "algorithmic_bytes_per_window": bytes_per_window,
                     "note": "fp32-ALU-bound stage (~113 FLOP per compulsory byte, SURVEY.md section 8(d)): frames this large run "
                             "the shared-memory pyramid kernel over a per-CTA global scratch"},
        "output_shapes": [list(o.shape) for o in outs]}))


if __name__ == "__main__":
    main()
