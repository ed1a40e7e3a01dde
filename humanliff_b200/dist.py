"""Data-parallel plumbing of the sampling path: one process per GPU, batch sharded across ranks with
NO data-path collective during the T-step loop (samples are independent: GroupNorm is per-sample,
there is no cross-sample op), then ONE all-gather of the finished samples (+ labels) -- exactly the
collective the reference issues at human_diffusion/scripts/triplane_sample_layered.py:211-219.
Backend: NCCL over NVLink 5 / NVSwitch on GPUs, gloo in the CPU tests.

Equal shards (the sampling scripts' case: every rank draws ``batch_size`` samples) take ONE collective: the int64
labels are bit-cast into the tail of the fp32 sample buffer, so samples + labels travel in a single
``all_gather_into_tensor`` into a receive buffer that is allocated once per (shape, world) and reused -- no size
exchange, no host synchronisation.  ``warm_up()`` runs that collective once on dummy data so that communicator /
NVLink channel set-up (tens of ms on the first NCCL call) never lands in a timed or latency-critical region."""
import torch
import torch.distributed as dist

_RECV = {}      # (device, dtype, send numel, world) -> (send staging buffer, receive buffer)


def shard_batch(global_batch, rank=None, world_size=None):
    """Contiguous [start, stop) slice of the global batch owned by ``rank`` (remainder to low ranks)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _buffers(device, dtype, n_send, world):
    key = (str(device), dtype, n_send, world)
    b = _RECV.get(key)
    if b is None:
        b = (torch.empty(n_send, dtype=dtype, device=device), torch.empty(world * n_send, dtype=dtype, device=device))
        _RECV[key] = b
    return b


def _gather_equal(sample, labels, group, world):
    """One all_gather_into_tensor: [sample words | labels bit-cast to the sample dtype] per rank."""
    B = sample.shape[0]
    flat = sample.reshape(-1)
    n_s = flat.numel()
    lab_words = 0
    if labels is not None:
        lab_raw = labels.contiguous().view(torch.uint8).view(sample.dtype)      # int64 -> 2 fp32 words each, bit-exact
        lab_words = lab_raw.numel()
    send, recv = _buffers(sample.device, sample.dtype, n_s + lab_words, world)
    send[:n_s].copy_(flat)
    if lab_words:
        send[n_s:].copy_(lab_raw)
    dist.all_gather_into_tensor(recv, send, group=group)
    per = recv.view(world, n_s + lab_words)
    out = per[:, :n_s].reshape((world * B,) + tuple(sample.shape[1:]))       # one strided copy out of the receive buffer
    lab = None
    if lab_words:
        lab = per[:, n_s:].contiguous().view(torch.uint8).view(labels.dtype).reshape(world * B)
    return out, lab


def all_gather_samples(sample, labels=None, group=None, equal_shards=None):
    """Gather ``sample [B_local, ...]`` (and int64 ``labels [B_local]``) from every rank, rank-major.

    ``equal_shards``: True -> every rank holds the same number of rows (the caller knows: ``shard_batch`` of a
    divisible batch, or the scripts' fixed per-rank batch): no size exchange; None -> one small all-gather of the
    row counts decides; ragged shards use the padded list form the reference uses."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sample, labels
    world = dist.get_world_size(group)
    sample = sample.contiguous()
    if equal_shards is None:
        mine = torch.tensor([sample.shape[0]], dtype=torch.int64, device=sample.device)
        sizes_t = torch.empty(world, dtype=torch.int64, device=sample.device)
        dist.all_gather_into_tensor(sizes_t, mine, group=group)
        sizes = sizes_t.tolist()                      # one host sync for all ranks' counts
        equal_shards = len(set(sizes)) == 1
    else:
        sizes = [sample.shape[0]] * world
    if equal_shards:
        return _gather_equal(sample, labels, group, world)
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


def warm_up(shape, device, dtype=torch.float32, with_labels=True, group=None):
    """Run the equal-shard gather once on zeros of the production shape: creates the communicator's channels and
    allocates the receive buffer the real call will reuse."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    s = torch.zeros(shape, dtype=dtype, device=device)
    lab = torch.zeros(shape[0], dtype=torch.int64, device=device) if with_labels else None
    all_gather_samples(s, lab, group=group, equal_shards=True)
