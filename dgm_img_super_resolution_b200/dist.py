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


def sample_sharded(diffusion, img_lr, img_lr_up, sample_fn=None, group=None, total=None):
    """Run `GaussianDiffusion.sample` on this rank's slice of the global batch and gather.

    img_lr / img_lr_up are the GLOBAL batch (same on every rank) -- or, when `total` is given, only THIS RANK'S
    slice `shard_range(total, world, rank)` of a global batch of `total` images (a host that feeds each GPU its own
    images need not ship the whole batch to every rank).  Returns (img, rrdb_out) for the global batch on every rank.  Noise is ALWAYS the in-kernel Philox stream here (`runtime.noise` is
    forced to "philox" for the call) keyed by the global image index (`runtime.image_offset` = slice
    start) and by ONE seed agreed on by all ranks (rank 0's, broadcast), so the result does not depend on
    how the batch is sharded: N ranks produce bit for bit what one GPU produces for the whole batch.
    (torch.randn-based noise would give every rank the same per-slice draws: correlated, shard-dependent.)
    """
    from . import model as M
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if total is None:
        total = img_lr.shape[0]
        start, count = shard_range(total, world, rank)
        lr, up = img_lr[start:start + count], img_lr_up[start:start + count]
    else:
        start, count = shard_range(total, world, rank)
        if img_lr.shape[0] != count:
            raise ValueError("rank %d owns %d of %d images but was given %d" % (rank, count, total, img_lr.shape[0]))
        lr, up = img_lr, img_lr_up
    old = (M.runtime.image_offset, M.runtime.noise, M.runtime.seed_counter, M.runtime.seed_override)
    # one seed for the whole job: rank 0 draws it, everybody uses it
    seed = torch.tensor([M.next_seed(img_lr.device) if rank == 0 else 0], dtype=torch.int64, device=img_lr.device)
    dist.broadcast(seed, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    M.runtime.image_offset, M.runtime.noise, M.runtime.seed_override = start, "philox", int(seed.item())
    try:
        if count > 0:
            fn = sample_fn or (lambda a, b, off: diffusion.sample(a, b, tuple(b.shape)))
            img, rrdb_out = fn(lr, up, start)
        else:
            img = img_lr_up.new_zeros((0,) + tuple(img_lr_up.shape[1:]))
            rrdb_out = img.clone()
    finally:
        M.runtime.image_offset, M.runtime.noise, _, M.runtime.seed_override = old
    return gather_outputs(img, total, group), gather_outputs(rrdb_out, total, group)
