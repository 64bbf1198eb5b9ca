"""Batch sharding for multi-GPU sampling: one process per GPU, weights replicated,
each rank samples a contiguous slice of the global batch with per-sample seeds (so the gathered
result does not depend on the number of GPUs), and ONE all-gather of the decoded images at the end.
The reference has no multi-GPU code at all (no torch.distributed call sites)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world_size):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def per_sample_randn(indices, shape, seed_base, device="cpu"):
    """Sample i of the GLOBAL batch is drawn from its own generator (seed_base + i): invariant to sharding."""
    out = []
    for i in indices:
        rng = np.random.Generator(np.random.PCG64(seed_base + int(i)))
        out.append(torch.from_numpy(rng.standard_normal(size=tuple(shape), dtype=np.float64).astype(np.float32)))
    t = torch.stack(out, 0) if out else torch.empty((0,) + tuple(shape))
    return t.to(device)


_gather_bufs = {}


def gather_images(local, global_batch, group=None):
    """All-gather per-rank image slices [b_r, ...] into [global_batch, ...] on every rank (NCCL on GPUs, gloo on CPU).

    Equal slices (the usual case: global_batch divisible by the world size) go through ONE `all_gather_into_tensor` straight
    into a pre-allocated [global_batch, ...] buffer that is reused from call to call — no per-rank list, no padding copy, no
    concatenation pass over the 200 MB of images.  Slices that differ in length by one are padded to the maximum first.
    The returned tensor is the reused buffer: clone it if it must survive the next call."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    lens = [hi - lo for lo, hi in sizes]
    local = local.contiguous()
    if min(lens) == max(lens):
        key = (tuple(local.shape), local.dtype, str(local.device), global_batch, id(group))
        out = _gather_bufs.get(key)
        if out is None:
            out = _gather_bufs[key] = torch.empty((global_batch,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    mx = max(lens)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
