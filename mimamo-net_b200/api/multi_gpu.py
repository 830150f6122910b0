"""Multi-GPU driver: videos shard embarrassingly across ranks.

The GRU couples the snippets of one batch and `Tester` batches are
per-video (api/tester.py:65-73), so the shardable unit is the VIDEO.  Ranks own contiguous blocks
of videos, replicate the weights, exchange nothing during compute and gather the per-video
predictions once at the end (one all-gather of a packed fixed-stride buffer + a length vector).
Works with backend "nccl" (one process per GPU, NVLink/NVSwitch) and "gloo" (CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of items owned by `rank`; earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_predictions(local_preds, n_videos_total, group=None, dst=None, to_host=False):
    """local_preds: list of (T_v, n_labels) float32 tensors for this rank's videos (in order).
    Returns the list for ALL videos, identical on every rank (dst=None: all-gather), or only on rank `dst`
    (gather; the other ranks get None).  Ragged lengths travel in the same exchange as a length vector; payloads are
    packed at a fixed stride (the global max length).  to_host=True hands the list back as CPU tensors through ONE
    device->host copy of the packed buffer (on rank `dst` only when a destination is given)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if not to_host or not local_preds:
            return list(local_preds)
        lens1 = [p.shape[0] for p in local_preds]
        flat = torch.cat(list(local_preds), 0).cpu()
        return list(flat.split(lens1))
    rank = dist.get_rank(group)
    device = local_preds[0].device if local_preds else torch.device('cuda', torch.cuda.current_device()) \
        if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    n_labels = local_preds[0].shape[-1] if local_preds else 2
    per_rank = max(shard_bounds(n_videos_total, r, world)[1] - shard_bounds(n_videos_total, r, world)[0]
                   for r in range(world))
    lens = torch.zeros(per_rank, dtype=torch.int64, device=device)
    for i, p in enumerate(local_preds):
        lens[i] = p.shape[0]
    all_lens = torch.empty(world * per_rank, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_lens, lens, group=group)
    t_max = int(all_lens.max().item()) if all_lens.numel() else 0
    packed = torch.zeros(per_rank, t_max, n_labels, dtype=torch.float32, device=device)
    for i, p in enumerate(local_preds):
        packed[i, :p.shape[0]] = p
    if dst is None:
        everything = torch.empty(world * per_rank, t_max, n_labels, dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(everything, packed, group=group)
    else:
        pieces = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
        dist.gather(packed, pieces, dst=dst, group=group)
        if rank != dst:
            return None
        everything = torch.cat(pieces, 0)
    if to_host:
        everything, all_lens = everything.cpu(), all_lens.cpu()
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_videos_total, r, world)
        for i in range(hi - lo):
            out.append(everything[r * per_rank + i, :int(all_lens[r * per_rank + i])])
    return out


def run_videos(tester, videos, group=None, dst=None, to_host=False, group_frames=4096):
    """BASELINE config 5: `videos` is the FULL list of videos (each uint8 (n_v, S, S, 3) aligned face crops, host or
    device); every rank runs Tester.predict_videos on its contiguous block and the per-video (n_v, 2) predictions are
    gathered once (to every rank, or to rank `dst` only).  Bit-identical for any world size, because a video is never
    split across ranks and the per-frame kernels do not depend on what else is in a batch."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(videos), rank, world)
    local = run_local_videos(tester, videos[lo:hi], group_frames)
    return gather_predictions(local, len(videos), group, dst=dst, to_host=to_host)


def run_local_videos(tester, videos, group_frames=4096):
    """This rank's block: host videos are copied to the device group by group on a copy stream, one group ahead of
    the compute (pinned host memory makes the copies asynchronous)."""
    device = torch.device('cuda', torch.cuda.current_device())
    main = torch.cuda.current_stream(device)
    if getattr(tester, '_video_copy_stream', None) is None:
        tester._video_copy_stream = torch.cuda.Stream(device)
    copy = tester._video_copy_stream
    groups, cur, total = [], [], 0
    for v in videos:
        v = torch.as_tensor(v)
        if cur and total + v.shape[0] > group_frames:
            groups.append(cur)
            cur, total = [], 0
        cur.append(v)
        total += v.shape[0]
    if cur:
        groups.append(cur)

    def stage(g):
        copy.wait_stream(main)
        with torch.cuda.stream(copy):
            dev = [v if v.is_cuda else v.to(device, non_blocking=True) for v in g]
            ev = torch.cuda.Event()
            ev.record(copy)
        return dev, ev

    out = []
    nxt = stage(groups[0]) if groups else None
    for gi in range(len(groups)):
        dev, ev = nxt
        nxt = stage(groups[gi + 1]) if gi + 1 < len(groups) else None
        main.wait_event(ev)
        for t in dev:
            t.record_stream(main)
        out += tester.predict_videos(dev, group_frames)
    return out
