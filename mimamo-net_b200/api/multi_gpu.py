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


def gather_predictions(local_preds, n_videos_total, group=None):
    """local_preds: list of (T_v, n_labels) float32 tensors for this rank's videos (in order).
    Returns the list for ALL videos, identical on every rank.  Ragged lengths travel in the same
    exchange as a length vector; payloads are packed at a fixed stride (the global max length)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return list(local_preds)
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
    everything = torch.empty(world * per_rank, t_max, n_labels, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(everything, packed, group=group)
    out = []
    for r in range(world):
        lo, hi = shard_bounds(n_videos_total, r, world)
        for i in range(hi - lo):
            out.append(everything[r * per_rank + i, :int(all_lens[r * per_rank + i])])
    return out


def run_videos(tester, videos, group=None):
    """BASELINE config 5: `videos` is the FULL list of videos (each uint8 (n_v, S, S, 3) aligned face crops, host or
    device); every rank runs Tester.predict_frames on its contiguous block and the per-video (n_v, 2) predictions are
    gathered once.  Returns the list for all videos on every rank; identical for any world size, because a video is
    never split across ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(videos), rank, world)
    device = torch.device('cuda', torch.cuda.current_device())
    local = [tester.predict_frames(torch.as_tensor(v).to(device, non_blocking=True)) for v in videos[lo:hi]]
    return gather_predictions(local, len(videos), group)
