"""Benchmark of MIMAMO-Net's per-window valence/arousal inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the full hot path (face-crop preprocessing, steerable pyramid + phase
difference, ResNet50 pool5, two-stream GRU head) over BASELINE.json configs[1]: 32 synthetic
64-frame 112x112 clips per GPU = 2048 face-windows (uint8 crops (32,64,112,112,3), seeded weights).
`value` times Tester.infer_crops with the crops resident in HBM, `e2e` times Tester.infer_crops_host
from pinned host memory; `float_inputs` reports the same two legs through the fp32-tensor entry
points (Tester.infer_clips / infer_clips_host: gray windows (32,64,13,48,48) + RGB (2048,3,224,224)
prepared on the host, the tensors the reference's DataLoader ships to the GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS, FRAMES, T, SIZE = 32, 64, 13, 48
WINDOWS = CLIPS * FRAMES
RESNET_FLOP = 7.712e9            # per image (3856 MMAC)
PHASENET_FLOP = 0.393e9          # per window
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the tcgen05 kernels from one `ncu --set full` capture of a
# 512-image ResNet50 pass (profiles/gemm_r1_ncu_final.txt): 22.28 GB over its 53 launches, 0.92x the algorithmic 24.16 GB.
# Traffic scales with the images per pass; the 12 PhaseNet launches per step were not captured (they move < 3 % of the bytes).
NCU_DRAM_BYTES_PER_IMAGE = 22.28e9 / 512
ALGO_BYTES_PER_IMAGE = 24.16e9 / 512
METRIC = "face-windows/sec end-to-end V/A inference"
UNIT = "windows/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}




def cpu_reference_rate(sample_windows=64, sample_images=16, threads=None):
    """The reference's CPU path (oracle port) on a bounded sample: windows/s per stage and composed."""
    from oracle import mimamo_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    gray = torch.rand(1, sample_windows, T, SIZE, SIZE, generator=g)
    rgb = torch.randint(0, 256, (sample_images, 3, 224, 224), generator=g).float() - torch.tensor(O.RESNET_MEAN)[None, :, None, None]
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    feats = torch.rand(1, sample_windows, 2048, generator=g)

    def best(fn, reps=2):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    with torch.no_grad():
        t_p = best(lambda: O.phase_diff_output(gray))
        p0, p1 = O.phase_diff_output(gray)
        t_h = best(lambda: O.head_forward(sd, p0, p1, feats))
        t_r = best(lambda: O.resnet_pool5(net, rgb))
    per_window = t_p / sample_windows + t_h / sample_windows + t_r / sample_images
    return 1.0 / per_window, {"pyramid_phase_windows_per_s": sample_windows / t_p, "head_windows_per_s": sample_windows / t_h,
                              "resnet50_images_per_s": sample_images / t_r}, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, stages, cores = cpu_reference_rate()
        vals.append(v)
    value = sum(vals) / len(vals)
    sample = "64 windows (pyramid+phase, head) + 16 images (ResNet50 fp32) per step, composed per window"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": 1, "ms_per_step": 1e3 * (time.perf_counter() - t0) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: full MIMAMO inference, 64-frame clips, batch 32 (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "stages": stages},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="print per-launch GEMM times to stderr")
    ap.add_argument("--quick", action="store_true", help="profiling runs: 1 warm-up, device-timed leg only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)                 # before the drop-in modules pick their default device
    import mimamo_b200
    mimamo_b200.install()
    import _native
    from tester import Tester
    from bench_inputs import make_crops, make_inputs, synthetic_weights

    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, set on some boxes) goes nowhere
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    if not args.quick:
        args.warmup = max(args.warmup, 3)

    resnet_sd, head_sd = synthetic_weights()
    tester = Tester(None, batch_size=CLIPS, resnet_model=resnet_sd, head_state_dict=head_sd)
    crops_h = make_crops(seed=100 + rank, clips=CLIPS, frames=FRAMES)                         # pinned host buffer
    crops_d = crops_h.to(dev)
    gathered = torch.empty(world * CLIPS, FRAMES, 2, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                             # > 126 MB L2

    def finish(out, to_host):
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)          # per-video predictions to every rank
            return gathered.cpu() if to_host else gathered
        return out.cpu() if to_host else out

    def step_device():
        flush.zero_()                                           # the 77 MB of crops must not survive in L2 between steps
        return finish(tester.infer_crops(crops_d), False)

    def step_e2e():
        # the public host-facing call: pinned host crops in, host predictions out
        flush.zero_()
        return finish(tester.infer_crops_host(crops_h, to_host=False), True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)                               # rank 0 samples its own GPU (one nvidia-smi poller per job, not per rank)
    if rank == 0:
        sampler.start()
    lib = _native.lib()
    lib.mimamo_profile_gemm(1)
    launches0 = _native.launch_count()
    ms = timed(step_device, args.steps)
    launches = _native.launch_count() - launches0
    import ctypes
    gemm_ms, gemm_n, issued = ctypes.c_double(0), ctypes.c_uint64(0), ctypes.c_double(0)
    lib.mimamo_profile_gemm_read(ctypes.byref(gemm_ms), ctypes.byref(gemm_n), ctypes.byref(issued))
    if args.layers:                                             # per-launch CUDA-event times of the GEMM kernels (steady state, real clocks)
        buf = (ctypes.c_float * 4096)()
        n = lib.mimamo_profile_gemm_launches(buf, 4096)
        per = n // args.steps
        avg = [sum(buf[k * per + i] for k in range(args.steps)) / args.steps * 1e3 for i in range(per)]
        print("gemm launches per step: %d; us per launch (avg over %d steps):" % (per, args.steps), file=sys.stderr)
        print(" ".join("%.0f" % v for v in avg), file=sys.stderr)
    lib.mimamo_profile_gemm(0)
    # per-stage device time (same inputs, each stage alone), reported next to the whole-step number
    def stage_ms(fn, reps=3):
        fn()
        return timed(fn, reps) / reps

    with torch.no_grad():
        pde, rn, hd = tester.phase_difference_extractor, tester.resnet50_extractor, tester.model
        pre = tester.crop_preprocessor()
        flat = crops_d.view(WINDOWS, 112, 112, 3)
        widx = tester.clip_window_index(CLIPS, FRAMES, dev)

        def pyramid_stage():
            return pde.phase_difference_indexed(pre.gray(flat), widx)

        p0, p1 = [d.view(CLIPS, FRAMES, -1, d.shape[-2], d.shape[-1]) for d in pyramid_stage()]
        feats = rn.features_from_crops(flat, pre).view(CLIPS, FRAMES, 2048)
        stages = {"pyramid_phase_ms": stage_ms(pyramid_stage),
                  "resnet50_ms": stage_ms(lambda: rn.features_from_crops(flat, pre)),
                  "head_ms": stage_ms(lambda: hd([p0, p1], feats))}
        del p0, p1, feats
    float_inputs = None
    if args.quick:
        e2e_ms = float("nan")
    else:
        for _ in range(2):
            step_e2e()
        e2e_ms = timed(step_e2e, args.steps)
    if not args.quick and world == 1:
        # the fp32-tensor entry points (what the reference's DataLoader would hand over), fewer steps; single-GPU runs only
        # (1.5 GB of pinned host memory per rank buys no extra information at N > 1)
        gray_h, rgb_h = make_inputs(seed=100 + rank, clips=CLIPS, frames=FRAMES, t=T, size=SIZE)
        gray_d, rgb_d = gray_h.to(dev), rgb_h.to(dev)
        n_f = max(2, min(args.steps, 4))
        tester.infer_clips(gray_d, rgb_d)
        f_dev = timed(lambda: finish(tester.infer_clips(gray_d, rgb_d), False), n_f)
        tester.infer_clips_host(gray_h, rgb_h)
        f_e2e = timed(lambda: finish(tester.infer_clips_host(gray_h, rgb_h, to_host=False), True), n_f)
        float_inputs = {"value": world * WINDOWS * n_f / (f_dev / 1e3), "e2e": world * WINDOWS * n_f / (f_e2e / 1e3),
                        "unit": UNIT, "steps": n_f, "h2d_bytes_per_step": world * (gray_h.numel() + rgb_h.numel()) * 4}
        del gray_d, rgb_d
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    value = world * WINDOWS * args.steps / (ms / 1e3)
    e2e_value = world * WINDOWS * args.steps / (e2e_ms / 1e3)
    peak_tf, peak_hbm, peak_src = peaks()
    algo_flops = (RESNET_FLOP + PHASENET_FLOP) * WINDOWS * args.steps
    achieved_tf = algo_flops / (gemm_ms.value / 1e3) / 1e12 if gemm_ms.value > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "configs[1]: full MIMAMO inference (SCFpyr+phase, ResNet50 pool5, 2-stream GRU) on "
                               "32 synthetic 64-frame 112x112 clips per GPU = 2048 face-windows/step/GPU",
                   "inputs": "uint8 face crops (32,64,112,112,3) per GPU; PIL-exact preprocessing on the device",
                   "dtypes": "preprocessing u8/int32, pyramid+phase f32, ResNet50 f16 (f32 accumulate), PhaseNet f16, dense+GRU f32",
                   "l2": "a 256 MB buffer is rewritten before every timed step (L2 flush)", "videos_sharded_by": "rank"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": world * crops_h.numel(),
                "d2h_bytes_per_step": world * (world if world > 1 else 1) * CLIPS * FRAMES * 2 * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "tcgen05 convolution engine: conv_gemm_kernel / conv3x3_halo_kernel / conv1_line_kernel (ResNet50 + PhaseNet convs)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                     "traffic": NCU_DRAM_BYTES_PER_IMAGE * WINDOWS * args.steps / gemm_n.value if gemm_n.value else None,
                     "traffic_unit": "bytes per launch (ncu dram read+write of a 512-image pass, scaled to this run's images per launch)",
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE * WINDOWS * args.steps / gemm_n.value if gemm_n.value else None,
                     "algorithmic_flop_per_launch": algo_flops / gemm_n.value if gemm_n.value else None,
                     "peak_source": peak_src, "kernel_ms_per_step": gemm_ms.value / args.steps,
                     "kernel_share_of_step": gemm_ms.value / ms if ms else None,
                     "launches_per_step": gemm_n.value / args.steps,
                     "issued_tflops": issued.value / (gemm_ms.value / 1e3) / 1e12 if gemm_ms.value > 0 else None},
        "stage_ms": stages,
        "clocks": sampler.summary(),
    }
    if float_inputs is not None:
        line["float_inputs"] = float_inputs
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        v, stages, cores = cpu_reference_rate()
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "64 windows (pyramid+phase, head) + 16 images (ResNet50 fp32), composed per window",
                                "stages": stages}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
