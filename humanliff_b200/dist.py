"""Data-parallel plumbing of the sampling path: one process per GPU, batch sharded across ranks with
NO data-path collective during the T-step loop (samples are independent: GroupNorm is per-sample,
there is no cross-sample op), then ONE all-gather of the finished samples (+ labels) -- exactly the
collective the reference issues at human_diffusion/scripts/triplane_sample_layered.py:211-219.
Backend: NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def shard_batch(global_batch, rank=None, world_size=None):
    """Contiguous [start, stop) slice of the global batch owned by ``rank`` (remainder to low ranks)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_samples(sample, labels=None, group=None):
    """Gather ``sample [B_local, ...]`` (and int64 ``labels [B_local]``) from every rank, rank-major.

    Equal shards use a single ``all_gather_into_tensor`` into one pre-allocated
    ``[world * B_local, ...]`` buffer (NCCL: one ncclAllGather over NVLink; labels ride in the same
    stream right behind it).  Ragged shards fall back to the list form the reference uses."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sample, labels
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=sample.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([sample.shape[0]], dtype=torch.int64, device=sample.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    sample = sample.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(sample.shape[1:]), dtype=sample.dtype, device=sample.device)
        dist.all_gather_into_tensor(out, sample, group=group)
        lab = None
        if labels is not None:
            lab = torch.empty(world * sizes[0], dtype=labels.dtype, device=labels.device)
            dist.all_gather_into_tensor(lab, labels.contiguous(), group=group)
        return out, lab
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(sample.shape[1:]), dtype=sample.dtype, device=sample.device)
    pad[:sample.shape[0]] = sample
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)
    lab = None
    if labels is not None:
        lpad = torch.zeros(mx, dtype=labels.dtype, device=labels.device)
        lpad[:labels.shape[0]] = labels
        lb = [torch.empty_like(lpad) for _ in range(world)]
        dist.all_gather(lb, lpad, group=group)
        lab = torch.cat([b[:n] for b, n in zip(lb, sizes)], 0)
    return out, lab
