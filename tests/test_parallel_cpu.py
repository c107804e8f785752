"""Batch-sharding host logic (open_universe_b200/parallel.py) on CPU with world_size-2 gloo:
contiguous row split, padding of uneven batches, all_gather, trimming.  The per-rank enhance()
is replaced by a deterministic stand-in (the CUDA path has no CPU fallback by design)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from open_universe_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_enhance(mix, **kw):
    # row-wise deterministic function: lets the test detect misplaced / duplicated rows
    return mix * 2.0 + mix.mean(dim=1, keepdim=True)


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    mix = torch.randn(batch, 37, generator=g)
    out = parallel.enhance_sharded(None, mix, enhance_fn=_fake_enhance)
    want = _fake_enhance(mix)
    q.put((rank, bool(torch.equal(out, want)), tuple(out.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5, 1])
def test_sharded_enhance_gloo(batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (batch, 37) for _, _, shape in res), res


def test_shard_bounds():
    assert parallel.shard_bounds(32, 8, 3) == (12, 16, 4)
    assert parallel.shard_bounds(5, 2, 1) == (3, 5, 3)
    assert parallel.shard_bounds(1, 2, 1) == (1, 1, 1)       # rank with no rows
    covered = []
    for r in range(4):
        lo, hi, per = parallel.shard_bounds(10, 4, r)
        covered += list(range(lo, hi))
    assert covered == list(range(10))


def test_global_noise_uses_the_seeded_generator():
    """enhance_sharded(seed=..., global_noise=True): the per-rank enhance() calls ``randn(x, sigma,
    rng=None)`` -- the patched function must still draw from the SEEDED generator (closure, not a
    default argument the caller overrides) and hand every shard its own rows of the
    global draw."""
    from open_universe_b200.networks.universe import universe as U

    def fake_enhance(mix, **kw):
        sigma = torch.ones(mix.shape[0])
        return mix + U.randn(mix[:, None, :], sigma, rng=kw.get("rng"))[:, 0, :]

    mix = torch.zeros(6, 11)
    a = parallel.enhance_sharded(None, mix, seed=5, enhance_fn=fake_enhance)
    torch.manual_seed(123)                      # the global generator must not matter
    b = parallel.enhance_sharded(None, mix, seed=5, enhance_fn=fake_enhance)
    want = torch.randn(6, 1, 11, generator=torch.Generator().manual_seed(5))[:, 0, :]
    assert torch.equal(a, b) and torch.equal(a, want)
    c = parallel.enhance_sharded(None, mix, seed=6, enhance_fn=fake_enhance)
    assert not torch.equal(a, c)
    assert U.randn is U._default_randn          # the hook is restored
