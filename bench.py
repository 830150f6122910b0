"""Benchmark of MIMAMO-Net's per-window valence/arousal inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config e2e|pyramid224|resnet512|videos] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config e2e (default, BASELINE.json configs[1]): a step = one pass of the full hot path (face-crop preprocessing,
    steerable pyramid + phase difference, ResNet50 pool5, two-stream GRU head) over 32 synthetic 64-frame 112x112 clips
    per GPU = 2048 face-windows (uint8 crops (32,64,112,112,3), seeded weights).  `value` times Tester.infer_crops with
    the crops resident in HBM, `e2e` times Tester.infer_crops_host from pinned host memory; `float_inputs` reports the
    same two legs through the fp32-tensor entry points the reference's DataLoader feeds.
--config pyramid224 (configs[2]): steerable pyramid + phase difference only, 256 windows x 13 frames of 224x224,
    height 5 / 8 orientations / levels [1,2,3]; roofline against the measured HBM copy bandwidth.
--config resnet512 (configs[3]): ResNet50 pool5 of 512 images 224x224 (fp16 operands, fp32 accumulation); roofline
    against the measured sustained tensor throughput.
--config videos (configs[4]): 128 synthetic 300-frame videos per GPU (1024 over 8 GPUs) through
    multi_gpu.run_videos -- 13-frame windows clamped to each video, 4 snippets + the overlapping tail snippet per
    video, one GRU batch per video, one gather of the per-video predictions to rank 0.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import ctypes
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
ALGO_BYTES_PER_IMAGE = 24.16e9 / 512      # 16-bit activations in + out of every ResNet50 layer + weights
METRIC = "face-windows/sec end-to-end V/A inference"
UNIT = "windows/s"
VIDEOS_PER_GPU, VIDEO_FRAMES = 128, 300
PYR = {"windows": 256, "T": 13, "H": 224, "height": 5, "nbands": 8, "levels": [1, 2, 3]}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json: sustained bf16 / copy bandwidth)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per ResNet50 image of the tcgen05 kernels, from the ncu capture of one 2048-image step of THIS build
    (profiles/r2_dram_traffic.json, written by tools_ncu_summary.py --traffic); None when the capture is missing."""
    path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if os.path.exists(path):
        return json.load(open(path))
    return None


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


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own code (baseline/_ref through oracle/ref_shim.py) where it exists, the oracle port otherwise
# ------------------------------------------------------------------------------------------------------------------
def _best(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


def load_reference():
    """The UNMODIFIED reference api/ staged under baseline/_ref (git-ignored; __graft_entry__.build() copies it from
    /root/reference in the build container and it travels to the GPU box with the snapshot), imported through the
    compatibility shim.  None when it is not there."""
    root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "api", "steerable")):
        return None
    os.environ["MIMAMO_REFERENCE_ROOT"] = root
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = root
    try:
        return ref_shim.load()
    except Exception as exc:                                     # the port below still gives a CPU number
        print("bench: reference under baseline/_ref could not be imported (%s); timing the oracle port" % exc, file=sys.stderr)
        return None


def cpu_pyramid_rate(R, windows, frame, height, nbands, levels):
    """windows/s of build_pyramid + extract on (windows, 13, frame, frame) gray stacks: reference class or oracle port."""
    from oracle import mimamo_oracle as O
    g = torch.Generator().manual_seed(0)
    x = torch.rand(windows, T, frame, frame, generator=g)
    if R is not None:
        pde = R.Phase_Difference_Extractor(height=height, nbands=nbands, extract_level=list(levels))

        def run():
            with torch.no_grad():
                return [pde.extract(c) for c in pde.build_pyramid(x)]
    else:
        def run():
            with torch.no_grad():
                return [O.extract(c) for c in O.build_pyramid(x, height, nbands, list(levels))]
    return windows / _best(run)


def cpu_reference_rate(sample_windows=64, sample_images=16, threads=None):
    """The reference's CPU path on a bounded sample of configs[1]: windows/s per stage and composed per window.
    Pyramid + phase and the two-stream head run the unmodified reference classes when baseline/_ref is present
    (kind "reference"); ResNet50 is always the restated architecture (its definition is not part of the reference)."""
    from oracle import mimamo_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    R = load_reference()
    g = torch.Generator().manual_seed(0)
    rgb = torch.randint(0, 256, (sample_images, 3, 224, 224), generator=g).float() - torch.tensor(O.RESNET_MEAN)[None, :, None, None]
    net = O.resnet_synthetic(1)
    sd = O.synthetic_state_dict(O.head_state_dict_spec(), seed=1)
    feats = torch.rand(1, sample_windows, 2048, generator=g)
    gray = torch.rand(1, sample_windows, T, SIZE, SIZE, generator=g)
    p0, p1 = O.phase_diff_output(gray)
    if R is not None:
        model = R.Two_Stream_RNN().eval()
        model.load_state_dict(sd)
        head = lambda: model([p0, p1], feats)
    else:
        head = lambda: O.head_forward(sd, p0, p1, feats)
    with torch.no_grad():
        r_p = cpu_pyramid_rate(R, sample_windows, SIZE, 4, 2, (1, 2))
        t_h = _best(head)
        t_r = _best(lambda: O.resnet_pool5(net, rgb))
    per_window = 1.0 / r_p + t_h / sample_windows + t_r / sample_images
    kind = "reference" if R is not None else "port"
    what = ("unmodified reference classes from baseline/_ref (Phase_Difference_Extractor, Two_Stream_RNN) + restated ResNet50"
            if R is not None else "oracle port")
    return 1.0 / per_window, {"pyramid_phase_windows_per_s": r_p, "head_windows_per_s": sample_windows / t_h,
                              "resnet50_images_per_s": sample_images / t_r}, threads, kind, what


def cpu_baseline_line(config):
    """The `cpu_baseline` object of a GPU line: the reference arm run in a SEPARATE process (the reference's module names
    -- steerable, phase_difference_extractor, mimamo_net -- collide with the drop-in modules this process has imported)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", config, "--steps", "1", "--warmup", "0"],
                             capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="1")).stdout
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
    except Exception as exc:
        return {"error": str(exc)}
    return {"error": "the reference arm printed no JSON line"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the configured workload on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mimamo_oracle as O
    steps = max(1, min(args.steps, 20))          # K steps as asked (each a bounded sample of a few seconds); more than 20 are capped and reported
    warm = max(0, min(args.warmup, 2))           # untimed steps on top of the warm-up call every step's own best-of-2 makes
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    vals, stages, metric, unit, workload, sample = [], None, METRIC, UNIT, None, None
    kind = "port"
    t0 = time.perf_counter()
    for it in range(warm + steps):
        if it == warm:
            vals, t0 = [], time.perf_counter()
        if args.config == "pyramid224":
            R = load_reference()
            kind = "reference" if R is not None else "port"
            vals.append(cpu_pyramid_rate(R, 2, PYR["H"], PYR["height"], PYR["nbands"], PYR["levels"]))
            metric, workload = "face-windows/sec, steerable pyramid + phase difference only", "configs[2] (bounded CPU sample)"
            sample = "2 windows x 13 frames of 224x224, height 5, 8 orientations, levels [1,2,3] per step (%s)" % (
                "unmodified Phase_Difference_Extractor from baseline/_ref" if R is not None else "oracle port")
        elif args.config == "resnet512":
            g = torch.Generator().manual_seed(0)
            rgb = torch.randint(0, 256, (16, 3, 224, 224), generator=g).float() - torch.tensor(O.RESNET_MEAN)[None, :, None, None]
            net = O.resnet_synthetic(1)
            vals.append(16 / _best(lambda: O.resnet_pool5(net, rgb)))
            metric, unit, workload = "images/sec, ResNet50 pool5_7x7_s1 features", "images/s", "configs[3] (bounded CPU sample)"
            sample = "16 images 224x224 fp32 per step, restated resnet50_ferplus_dag (the definition is not part of the reference), torch CPU"
        else:
            v, stages, cores, kind, what = cpu_reference_rate()
            vals.append(v)
            workload = ("configs[4]: per-video inference, 300-frame videos (bounded CPU sample)" if args.config == "videos"
                        else "configs[1]: full MIMAMO inference, 64-frame clips, batch 32 (bounded CPU sample)")
            sample = "64 windows (pyramid+phase, head) + 16 images (ResNet50 fp32) per step, composed per window; " + what
    value = sum(vals) / len(vals)
    base = {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample}
    if stages:
        base["stages"] = stages
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm, "ms_per_step": 1e3 * (time.perf_counter() - t0) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload}, "cpu_baseline": base,
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------------------------
class Ctx(object):
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.args = args
        # stdout carries exactly ONE JSON line: anything the libraries print on fd 1 meanwhile (NCCL's version banner, ...)
        # is sent to stderr, and the line itself is written to the saved descriptor by finish()
        sys.stdout.flush()
        self._stdout_fd = os.dup(1)
        os.dup2(2, 1)
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
        torch.cuda.set_device(self.local)            # before the drop-in modules pick their default device
        import mimamo_b200
        mimamo_b200.install()
        import _native
        self.native = _native
        self.lib = _native.lib()
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, set on some boxes) goes nowhere
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        if not args.quick:
            args.warmup = max(args.warmup, 3)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.sampler = ClockSampler(self.local)       # rank 0 samples its own GPU (one nvidia-smi poller per job)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """ms for `steps` calls: CUDA events on the current stream, barrier + synchronize on both sides, MAX over ranks."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        self.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def gemm_profile(self, fn, steps):
        """Times fn like `timed` with the per-launch CUDA events of the tcgen05 kernels switched on."""
        self.lib.mimamo_profile_gemm(1)
        l0 = self.native.launch_count()
        ms = self.timed(fn, steps)
        launches = self.native.launch_count() - l0
        gemm_ms, gemm_n, issued = ctypes.c_double(0), ctypes.c_uint64(0), ctypes.c_double(0)
        self.lib.mimamo_profile_gemm_read(ctypes.byref(gemm_ms), ctypes.byref(gemm_n), ctypes.byref(issued))
        per = None
        if self.args.layers:
            buf = (ctypes.c_float * 8192)()
            n = self.lib.mimamo_profile_gemm_launches(buf, 8192)
            k = n // steps
            per = [sum(buf[s * k + i] for s in range(steps)) / steps * 1e3 for i in range(k)]
            print("gemm launches per step: %d; us per launch (avg over %d steps):" % (k, steps), file=sys.stderr)
            print(" ".join("%.0f" % v for v in per), file=sys.stderr)
        self.lib.mimamo_profile_gemm(0)
        return ms, launches, gemm_ms.value, gemm_n.value, issued.value

    def finish(self, line):
        self.sampler.stop_flag = True
        if self.rank == 0:
            if self.sampler.is_alive():
                self.sampler.join(timeout=2)
            line["clocks"] = self.sampler.summary()
            sys.stdout.flush()
            os.write(self._stdout_fd, (json.dumps(line) + "\n").encode())
        if self.world > 1:
            self.dist.destroy_process_group()


def tensor_roofline(gemm_ms, gemm_n, issued, algo_flops, images, steps, step_ms, kernel):
    peak_tf, _, src = peaks()
    achieved = algo_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    tr = measured_traffic()
    traffic = tr["dram_bytes_per_image"] * images * steps / gemm_n if (tr and gemm_n) else None
    return {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
            "traffic_unit": "bytes per launch" + ((": " + tr["source"]) if tr else " (no ncu capture of this build under profiles/)"),
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE * images * steps / gemm_n if gemm_n else None,
            "algorithmic_flop_per_launch": algo_flops / gemm_n if gemm_n else None, "peak_source": src,
            "kernel_ms_per_step": gemm_ms / steps, "kernel_share_of_step": gemm_ms / step_ms if step_ms else None,
            "launches_per_step": gemm_n / steps, "issued_tflops": issued / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None}


def bench_e2e(ctx):
    args, world, rank, dev, dist = ctx.args, ctx.world, ctx.rank, ctx.dev, ctx.dist
    from tester import Tester
    from bench_inputs import make_crops, make_inputs, synthetic_weights
    resnet_sd, head_sd = synthetic_weights()
    tester = Tester(None, batch_size=CLIPS, resnet_model=resnet_sd, head_state_dict=head_sd)
    crops_h = make_crops(seed=100 + rank, clips=CLIPS, frames=FRAMES)                         # pinned host buffer
    crops_d = crops_h.to(dev)
    gathered = [torch.empty(CLIPS, FRAMES, 2, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def finish(out, to_host):
        # per-video predictions: one gather to rank 0, one device->host copy there
        if world > 1:
            dist.gather(out, gathered, dst=0)
            if rank == 0 and to_host:
                return torch.stack(gathered).cpu()
            return out
        return out.cpu() if to_host else out

    def step_device():
        ctx.flush.zero_()                                       # the 77 MB of crops must not survive in L2 between steps
        return finish(tester.infer_crops(crops_d), False)

    def step_e2e():
        # the public host-facing call: pinned host crops in, host predictions out
        ctx.flush.zero_()
        return finish(tester.infer_crops_host(crops_h, to_host=False), True)

    for _ in range(args.warmup):
        step_device()
    if rank == 0:
        ctx.sampler.start()
    ms, launches, gemm_ms, gemm_n, issued = ctx.gemm_profile(step_device, args.steps)

    def stage_ms(fn, reps=10):
        fn()
        return ctx.timed(fn, reps) / reps

    with torch.no_grad():
        pde, rn, hd = tester.phase_difference_extractor, tester.resnet50_extractor, tester.model
        pre = tester.crop_preprocessor()
        flat = crops_d.view(WINDOWS, 112, 112, 3)
        widx = tester.clip_window_index(CLIPS, FRAMES, dev)

        def pyramid_stage():                                    # gray + pyramid + phase tail, as infer_crops runs it
            return tester._phase_streams(pre.gray(flat), widx)

        streams = pyramid_stage()
        feats = rn.features_from_crops(flat, pre)
        stages = {"pyramid_phase_ms": stage_ms(pyramid_stage),
                  "resnet50_ms": stage_ms(lambda: rn.features_from_crops(flat, pre)),
                  "head_ms": stage_ms(lambda: tester._head(streams, slice(0, WINDOWS), feats, CLIPS, FRAMES)),
                  "phase_route": streams[0]}
        del streams, feats
    float_inputs = None
    if args.quick:
        e2e_ms = float("nan")
    else:
        for _ in range(2):
            step_e2e()
        e2e_ms = ctx.timed(step_e2e, args.steps)
    if not args.quick and world == 1:
        # the fp32-tensor entry points (what the reference's DataLoader would hand over), fewer steps; single-GPU runs only
        gray_h, rgb_h = make_inputs(seed=100 + rank, clips=CLIPS, frames=FRAMES, t=T, size=SIZE)
        gray_d, rgb_d = gray_h.to(dev), rgb_h.to(dev)
        n_f = max(2, min(args.steps, 4))
        tester.infer_clips(gray_d, rgb_d)
        f_dev = ctx.timed(lambda: finish(tester.infer_clips(gray_d, rgb_d), False), n_f)
        tester.infer_clips_host(gray_h, rgb_h)
        f_e2e = ctx.timed(lambda: finish(tester.infer_clips_host(gray_h, rgb_h, to_host=False), True), n_f)
        float_inputs = {"value": world * WINDOWS * n_f / (f_dev / 1e3), "e2e": world * WINDOWS * n_f / (f_e2e / 1e3),
                        "unit": UNIT, "steps": n_f, "h2d_bytes_per_step": world * (gray_h.numel() + rgb_h.numel()) * 4}
        del gray_d, rgb_d
    algo_flops = (RESNET_FLOP + PHASENET_FLOP) * WINDOWS * args.steps
    _, peak_hbm, _ = peaks()
    pyr_bytes = 4 * T * SIZE * SIZE + 4 * 2 * (T - 1) * (SIZE * SIZE + (SIZE // 2) ** 2)
    line = {
        "metric": METRIC, "value": world * WINDOWS * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "configs[1]: full MIMAMO inference (SCFpyr+phase, ResNet50 pool5, 2-stream GRU) on "
                               "32 synthetic 64-frame 112x112 clips per GPU = 2048 face-windows/step/GPU",
                   "inputs": "uint8 face crops (32,64,112,112,3) per GPU; PIL-exact preprocessing on the device",
                   "dtypes": "preprocessing u8/int32, pyramid+phase f32, ResNet50 f16 (f32 accumulate), PhaseNet f16, dense+GRU f32",
                   "l2": "a 256 MB buffer is rewritten before every timed step (L2 flush)", "videos_sharded_by": "rank"},
        "e2e": {"value": world * WINDOWS * args.steps / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": world * crops_h.numel(),
                "d2h_bytes_per_step": world * CLIPS * FRAMES * 2 * 4},          # one copy of the gathered predictions, on rank 0
        "gpu_launches": launches,
        "roofline": tensor_roofline(gemm_ms, gemm_n, issued, algo_flops, WINDOWS, args.steps, ms,
                                    "tcgen05 convolution engine: conv_gemm_kernel / conv_gemm2_kernel / conv3x3_halo_kernel / "
                                    "conv1_line_kernel (ResNet50 + PhaseNet convs)"),
        "stage_ms": stages,
        "pyramid_stage_hbm": {"achieved_gbs": pyr_bytes * WINDOWS / (stages["pyramid_phase_ms"] / 1e3) / 1e9, "peak_gbs": peak_hbm,
                              "frac": pyr_bytes * WINDOWS / (stages["pyramid_phase_ms"] / 1e3) / 1e9 / peak_hbm,
                              "algorithmic_bytes_per_window": pyr_bytes},
    }
    if float_inputs is not None:
        line["float_inputs"] = float_inputs
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline_line("e2e")
    ctx.finish(line)


def bench_pyramid224(ctx):
    """configs[2]: the pyramid + phase stage alone at 224x224 -- the HBM-roofline run."""
    args, dev = ctx.args, ctx.dev
    from phase_difference_extractor import Phase_Difference_Extractor
    W, H, nb = PYR["windows"], PYR["H"], PYR["nbands"]
    height = args.pyr_height or PYR["height"]
    levels = [int(v) for v in args.pyr_levels.split(",")] if args.pyr_levels else list(PYR["levels"])
    pde = Phase_Difference_Extractor(height=height, nbands=nb, extract_level=levels)
    g = torch.Generator().manual_seed(200 + ctx.rank)
    frames_h = torch.rand(W, T, H, H, generator=g).pin_memory()                # distinct frames: nothing to de-duplicate
    frames = frames_h.to(dev)
    # the same 256 windows as they occur in Tester: sliding 13-frame windows over 4 clips of 64 frames
    clip = torch.rand(4 * 64, H, H, generator=g).to(dev)
    half = (T - 1) // 2
    cidx = (torch.arange(64)[:, None] + torch.arange(-half, half + 1)[None, :]).clamp_(0, 63)
    cidx = (cidx[None] + (torch.arange(4) * 64)[:, None, None]).reshape(W, T).to(device=dev, dtype=torch.int32)
    steps = args.steps

    def step_device():
        return pde.phase_difference(frames)

    def step_clip():
        return pde.phase_difference_indexed(clip, cidx)

    def step_e2e():
        outs = pde.phase_difference(frames_h.to(dev, non_blocking=True))
        return [o.sum(dim=(-1, -2)).cpu() for o in outs]          # per-map sums stand in for the consumer (PhaseNet stays on the device)

    for _ in range(max(1, min(args.warmup, 3))):
        outs = step_device()
    if ctx.rank == 0:
        ctx.sampler.start()
    l0 = ctx.native.launch_count()
    ms = ctx.timed(step_device, steps)
    launches = ctx.native.launch_count() - l0
    step_clip()
    ms_clip = ctx.timed(step_clip, steps)
    step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps)
    bytes_per_window = 4 * T * H * H + 4 * nb * (T - 1) * sum((H >> (l - 1)) ** 2 for l in levels)
    _, peak, src = peaks()
    world = ctx.world
    gbs = bytes_per_window * W * steps / (ms / 1e3) / 1e9
    line = {
        "metric": "face-windows/sec, steerable pyramid + phase difference only", "value": world * W * steps / (ms / 1e3),
        "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[2]: steerable pyramid + phase difference only, %d windows x %d DISTINCT frames of %dx%d per GPU, "
                               "height %d, 8 orientations, levels %s" % (W, T, H, H, height, str(levels).replace(" ", "")),
                   "l2": "inputs + outputs (%.1f GB per step) exceed the 126 MB L2" % (bytes_per_window * W / 1e9)},
        "e2e": {"value": world * W * steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": world * frames_h.numel() * 4,
                "d2h_bytes_per_step": world * sum(o.shape[0] * o.shape[1] * o.shape[2] for o in outs) * 4,
                "note": "frames from pinned host memory; the per-map sums of the phase maps are read back (their consumer, PhaseNet, is on the device)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "pyr_build_umma_kernel (tcgen05 kind::tf32, 3xTF32) + phase_tail_kernel (whole stage)",
                     "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                     "algorithmic_bytes_per_window": bytes_per_window, "peak_source": src,
                     "note": "algorithmic bytes = 4*T*H^2 in + 4*nb*(T-1)*sum_l (H/2^(l-1))^2 out per window (SURVEY.md section 8(d)); "
                             "the stage is bound by the dense transform (1.7 GFLOP per frame as split-TF32 tensor-core products, their "
                             "shared-memory staging) and the tail's fp32 instruction issue, not by HBM (DESIGN.md section 3.1)"},
        "sliding_windows": {"value": world * W * steps / (ms_clip / 1e3), "unit": UNIT, "ms_per_step": ms_clip / steps,
                            "frac": bytes_per_window * W * steps / (ms_clip / 1e3) / 1e9 / peak,
                            "note": "the same 256 windows taken as sliding windows over 4 clips of 64 frames (what Tester feeds): "
                                    "each distinct frame is transformed once (mimamo_pyr_phase_indexed)"},
        "output_shapes": [list(o.shape) for o in outs],
    }
    if ctx.rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline_line("pyramid224")
    ctx.finish(line)


def bench_resnet512(ctx):
    """configs[3]: ResNet50 pool5 features of 512 images, the tensor-core path alone."""
    args, dev = ctx.args, ctx.dev
    from resnet50_extractor import Resnet50_Extractor
    from bench_inputs import synthetic_weights, RESNET_MEAN
    B = 512
    resnet_sd, _ = synthetic_weights()
    rn = Resnet50_Extractor(model=resnet_sd)
    g = torch.Generator().manual_seed(300 + ctx.rank)
    x_h = torch.randint(0, 256, (B, 3, 224, 224), generator=g, dtype=torch.uint8).float()
    x_h -= torch.tensor(RESNET_MEAN)[None, :, None, None]
    x_h = x_h.pin_memory()
    x = x_h.to(dev)
    steps = args.steps

    def step_device():
        ctx.flush.zero_()
        return rn.features(x)

    def step_e2e():
        ctx.flush.zero_()
        return rn.features_host(x_h, chunk=128, to_host=True)

    for _ in range(args.warmup):
        step_device()
    if ctx.rank == 0:
        ctx.sampler.start()
    ms, launches, gemm_ms, gemm_n, issued = ctx.gemm_profile(step_device, steps)
    step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps)
    world = ctx.world
    line = {
        "metric": "images/sec, ResNet50 pool5_7x7_s1 features", "value": world * B * steps / (ms / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "configs[3]: ResNet50 pool5_7x7_s1 feature extraction, 512 images 224x224 per GPU, 16-bit operands "
                               "(fp16: bf16 misses the 1e-3 valence/arousal budget, DESIGN.md section 2), fp32 accumulation",
                   "inputs": "fp32 NCHW (512,3,224,224), 0-255 minus mean -- what Resnet50_Extractor.get_vec is fed",
                   "l2": "a 256 MB buffer is rewritten before every timed step (L2 flush)"},
        "e2e": {"value": world * B * steps / (ms_e2e / 1e3), "unit": "images/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": world * x_h.numel() * 4, "d2h_bytes_per_step": world * B * 2048 * 4},
        "gpu_launches": launches,
        "roofline": tensor_roofline(gemm_ms, gemm_n, issued, RESNET_FLOP * B * steps, B, steps, ms,
                                    "tcgen05 convolution engine (53 ResNet50 layers)"),
    }
    if ctx.rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline_line("resnet512")
    ctx.finish(line)


def bench_videos(ctx):
    """configs[4]: per-video inference sharded by video, 128 videos of 300 frames per GPU, one gather to rank 0."""
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    from tester import Tester
    from multi_gpu import run_local_videos, gather_predictions
    from bench_inputs import synthetic_weights
    resnet_sd, head_sd = synthetic_weights()
    tester = Tester(None, batch_size=8, resnet_model=resnet_sd, head_state_dict=head_sd)    # a video's 5 snippets = one batch
    n_local = VIDEOS_PER_GPU if not args.quick else 16
    g = torch.Generator().manual_seed(400 + rank)
    block = torch.randint(0, 256, (n_local, VIDEO_FRAMES, 112, 112, 3), generator=g, dtype=torch.uint8).pin_memory()
    videos_h = [block[i] for i in range(n_local)]
    videos_d = [v.to(dev) for v in videos_h]
    windows = n_local * VIDEO_FRAMES
    steps = max(1, min(args.steps, 5))

    def step(videos, to_host):
        local = run_local_videos(tester, videos)
        return gather_predictions(local, n_local * world, dst=0 if world > 1 else None, to_host=to_host)

    for _ in range(min(args.warmup, 2)):
        step(videos_d, False)
    if rank == 0:
        ctx.sampler.start()
    ms, launches, gemm_ms, gemm_n, issued = ctx.gemm_profile(lambda: step(videos_d, False), steps)
    step(videos_h, True)
    ms_e2e = ctx.timed(lambda: step(videos_h, True), steps)
    line = {
        "metric": METRIC, "value": world * windows * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "configs[4]: end-to-end valence/arousal over %d synthetic %d-frame 112x112 face videos per GPU (1024 over 8 GPUs), "
                               "sharded by video; windows clamped to the video, 4 snippets + tail snippet per video, one GRU batch per video, "
                               "one gather of the per-video predictions to rank 0" % (n_local, VIDEO_FRAMES),
                   "inputs": "uint8 aligned face crops (300,112,112,3) per video",
                   "l2": "%.2f GB of crops per step per GPU exceed the 126 MB L2" % (block.numel() / 1e9)},
        "e2e": {"value": world * windows * steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": world * block.numel(), "d2h_bytes_per_step": world * windows * 2 * 4},
        "gpu_launches": launches,
        "roofline": tensor_roofline(gemm_ms, gemm_n, issued, (RESNET_FLOP + PHASENET_FLOP) * windows * steps, windows, steps, ms,
                                    "tcgen05 convolution engine (ResNet50 over groups of whole videos, PhaseNet per video)"),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        line["cpu_baseline"] = cpu_baseline_line("videos")
    ctx.finish(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="e2e", choices=["e2e", "pyramid224", "resnet512", "videos"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pyr-height", type=int, default=None,
                    help="pyramid224 only: pyramid height (SURVEY 8(d) secondary number: 6 = four oriented scales, the maximum at 224x224)")
    ap.add_argument("--pyr-levels", default=None, help="pyramid224 only: comma-separated extract levels, e.g. 1,2,3,4")
    ap.add_argument("--layers", action="store_true", help="print per-launch GEMM times to stderr")
    ap.add_argument("--quick", action="store_true", help="profiling runs: 1 warm-up, device-timed leg only")
    args = ap.parse_args()
    if args.impl == "reference":
        # the reference's CPU path on the host cores: its own get_device() picks cuda:0 whenever a GPU is visible (and then asserts
        # on the CPU tensors it is fed), so the arm hides the GPUs from this process before CUDA is initialised
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        return run_reference(args)
    ctx = Ctx(args)
    {"e2e": bench_e2e, "pyramid224": bench_pyramid224, "resnet512": bench_resnet512, "videos": bench_videos}[args.config](ctx)


if __name__ == "__main__":
    main()
