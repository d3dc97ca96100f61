"""Channel sharding across the GPUs of one box.

Every reference object is one channel with private state, so the multi-GPU form is a
partition of the channel axis: rank r owns a contiguous range and the state of those channels stays
resident on its GPU.  There is no collective on the data path; `gather_audio` is the optional
collection of the (decimated, small) outputs on one rank.
"""


def channel_range(total_channels, rank, world_size):
    """Contiguous channel range [lo, hi) of `rank`; the first `total % world` ranks hold one more."""
    if world_size < 1 or not (0 <= rank < world_size) or total_channels < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(total_channels, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_audio(local, dst=0, group=None):
    """Gather per-rank output blocks `[channels_r x samples]` (torch tensors, same sample count) on
    rank `dst` in channel order.  Works with the NCCL backend (GPU tensors, NVLink) and with gloo."""
    import torch
    import torch.distributed as dist

    # `dst` is a GLOBAL rank, as torch.distributed.gather takes it; compare like with like (a sub-group's local ranks
    # need not be 0..N-1 of the job)
    world, rank = dist.get_world_size(group), dist.get_rank()
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
