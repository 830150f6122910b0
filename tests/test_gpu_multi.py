"""Multi-GPU determinism (SURVEY.md sections 4(v), 8(e)): the gathered per-video predictions of multi_gpu.run_videos at
world size 2 / 4 / 8 (NCCL, one process per GPU) are BIT-IDENTICAL to a single process running every video -- videos are
never split across ranks and no kernel depends on what else shares its batch.  Each case needs that many GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LENGTHS = [150, 64, 300, 300, 70, 13, 129, 300, 150, 150]   # ragged: tail snippets, a short video, uneven shards, runs of equal
                                                            # lengths (those share a head forward: more GRU batch rows)


def _videos():
    rng = np.random.default_rng(5)
    return [torch.from_numpy(rng.integers(0, 256, (n, 112, 112, 3), dtype=np.uint8)) for n in LENGTHS]


def _tester():
    sys.path.insert(0, ROOT)
    import mimamo_b200
    mimamo_b200.install()
    from bench_inputs import synthetic_weights
    from tester import Tester
    resnet_sd, head_sd = synthetic_weights()
    return Tester(None, batch_size=8, resnet_model=resnet_sd, head_state_dict=head_sd)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tester = _tester()
    from multi_gpu import run_videos
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    everything = run_videos(tester, _videos())                          # all-gather: every rank holds every video
    on_zero = run_videos(tester, _videos(), dst=0, to_host=True)        # gather to rank 0's host only
    ok = (on_zero is None) == (rank != 0)
    if rank == 0:
        ok = ok and all(torch.equal(a.cpu(), b) for a, b in zip(everything, on_zero))
    q.put((rank, ok, [p.cpu().numpy() for p in everything]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_bit_identical_to_one(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    tester = _tester()
    single = [tester.predict_frames(v.to(cuda)).cpu().numpy() for v in _videos()]      # one video at a time, like Tester.test
    from multi_gpu import run_videos
    grouped = [p.cpu().numpy() for p in run_videos(tester, _videos())]                  # world size 1, videos grouped per pass
    for a, b in zip(single, grouped):
        assert np.array_equal(a, b)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, ok, preds in results:
        assert ok and len(preds) == len(LENGTHS)
        for a, b in zip(single, preds):
            assert a.shape == b.shape and np.array_equal(a, b), "rank %d: gathered predictions differ from the single-process run" % rank
