"""CPU (gloo, world_size 2) test of the multi-GPU host logic: shard ranges and the output gather.
The per-rank compute is replaced by a deterministic function of the global image index so the
plumbing is checked without a GPU."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_sample(lr, up, offset):
    idx = torch.arange(offset, offset + lr.shape[0], dtype=torch.float32).view(-1, 1, 1, 1)
    return up * 2 + idx, up - idx


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    from dgm_img_super_resolution_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    lr = torch.randn((total, 3, 4, 4), generator=g)
    up = torch.randn((total, 3, 16, 16), generator=g)
    img, rr = D.sample_sharded(None, lr, up, sample_fn=_fake_sample)
    want_img, want_rr = _fake_sample(lr, up, 0)
    ok = torch.equal(img, want_img) and torch.equal(rr, want_rr)
    q.put((rank, bool(ok), tuple(img.shape)))
    dist.destroy_process_group()


def _run(total, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    return res


def test_shard_ranges_cover_batch():
    from dgm_img_super_resolution_b200.dist import shard_range
    for total in (0, 1, 5, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_gather_even_and_ragged_world2():
    for total in (4, 5, 1):
        res = _run(total)
        assert len(res) == 2
        for rank, ok, shape in res:
            assert ok, (total, rank)
            assert shape == (total, 3, 16, 16)


def _worker_slices(rank, world, port, total, q):
    import torch.distributed as dist
    from dgm_img_super_resolution_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    lr = torch.randn((total, 3, 4, 4), generator=g)
    up = torch.randn((total, 3, 16, 16), generator=g)
    start, count = D.shard_range(total, world, rank)
    img, rr = D.sample_sharded(None, lr[start:start + count], up[start:start + count], sample_fn=_fake_sample, total=total)
    want_img, want_rr = _fake_sample(lr, up, 0)
    ok = torch.equal(img, want_img) and torch.equal(rr, want_rr)
    try:
        D.sample_sharded(None, lr, up, sample_fn=_fake_sample, total=total + world)     # wrong slice size
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok), tuple(img.shape)))
    dist.destroy_process_group()


def test_sample_sharded_accepts_per_rank_slices_world2():
    """`total=`: every rank passes only ITS slice of the global batch (ragged too); the gathered result is the same."""
    ctx = mp.get_context("spawn")
    for total in (6, 5):
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker_slices, args=(r, 2, port, total, q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=120) for _ in ps]
        for p in ps:
            p.join(timeout=60)
        for rank, ok, shape in res:
            assert ok, (total, rank)
            assert shape == (total, 3, 16, 16)
