"""Batch-sharded multi-GPU sampling: images never interact, so each rank runs the
whole path on a contiguous slice of the batch with its own CUDA graph; the ONLY collective is
the final gather of the outputs.  Nothing crosses NVLink inside the step loop.

The reference has no multi-GPU path (its single `dist.all_reduce` lives in dead code,
srdiff/model.py:1315-1324); this is new plumbing over torch.distributed (NCCL on GPUs, gloo in
the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Contiguous slice [start, start+count) of `total` images owned by `rank`; the first
    `total % world` ranks get one extra image."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_outputs(local, total, group=None):
    """All-gather per-rank [count_r, ...] tensors into [total, ...] on every rank (ragged shards are
    padded to the largest shard for the collective, then trimmed)."""
    world = dist.get_world_size(group)
    counts = [shard_range(total, world, r)[1] for r in range(world)]
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    buf = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, pad.contiguous(), group=group)
    parts = [buf[r * mx: r * mx + counts[r]] for r in range(world)]
    return torch.cat(parts, 0)


def sample_sharded(diffusion, img_lr, img_lr_up, sample_fn=None, group=None):
    """Run `GaussianDiffusion.sample` on this rank's slice of the global batch and gather.

    img_lr / img_lr_up are the GLOBAL batch (same on every rank).  Returns (img, rrdb_out) for the
    global batch on every rank.  `runtime.image_offset` is set to the slice start so in-kernel
    Philox noise depends only on the global image index (shard-invariant results).
    """
    from . import model as M
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = img_lr.shape[0]
    start, count = shard_range(total, world, rank)
    lr, up = img_lr[start:start + count], img_lr_up[start:start + count]
    old = M.runtime.image_offset
    M.runtime.image_offset = start
    try:
        if count > 0:
            fn = sample_fn or (lambda a, b, off: diffusion.sample(a, b, tuple(b.shape)))
            img, rrdb_out = fn(lr, up, start)
        else:
            img = img_lr_up.new_zeros((0,) + tuple(img_lr_up.shape[1:]))
            rrdb_out = img.clone()
    finally:
        M.runtime.image_offset = old
    return gather_outputs(img, total, group), gather_outputs(rrdb_out, total, group)
