"""CPU (gloo, world_size 2): the video-sharding and prediction-gather logic of the multi-GPU driver."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lengths, q):
    sys.path.insert(0, ROOT)
    import mimamo_b200
    mimamo_b200.install()
    from multi_gpu import gather_predictions, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(len(lengths), rank, world)
    # "predictions" are a deterministic function of (video, frame) so every rank can check all of them
    local = [torch.arange(lengths[v] * 2, dtype=torch.float32).view(lengths[v], 2) + 1000 * v for v in range(lo, hi)]
    everything = gather_predictions(local, len(lengths))
    ok = len(everything) == len(lengths)
    for v, p in enumerate(everything):
        want = torch.arange(lengths[v] * 2, dtype=torch.float32).view(lengths[v], 2) + 1000 * v
        ok = ok and p.shape == want.shape and torch.equal(p, want)
    # gather to one rank only (bench.py / run_videos(dst=0)): rank 0 gets everything as host tensors, rank 1 nothing
    on_zero = gather_predictions(local, len(lengths), dst=0, to_host=True)
    ok = ok and ((on_zero is None) if rank != 0 else all(torch.equal(a, b) for a, b in zip(on_zero, everything)))
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


@pytest.mark.parametrize("lengths", [[300, 64, 70, 129, 13], [64, 64], [10]])
def test_shard_and_gather_world2(lengths):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + len(lengths)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in results)
    bounds = sorted(b for _, _, b in results)
    assert bounds[0][0] == 0 and bounds[-1][1] == len(lengths) and bounds[0][1] == bounds[1][0]


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    import mimamo_b200
    mimamo_b200.install()
    from multi_gpu import shard_bounds
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
