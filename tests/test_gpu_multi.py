"""Batch-sharded enhance() on REAL GPUs over NCCL (needs >= 2 devices: skipped on a single-GPU box; run
with ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``): the sharded result equals
the unsharded run on the same globally drawn, seeded noise."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, q):
    import torch.distributed as dist
    from open_universe_b200 import parallel
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.networks.universe import universe as U
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    m.eval(no_ema=True)
    m = m.to(dev)
    g = torch.Generator().manual_seed(1)
    mix = (0.05 * torch.randn(batch, 8000, generator=g)).to(dev)
    out = parallel.enhance_sharded(m, mix, seed=77, global_noise=True, n_steps=4)
    ok, err = True, 0.0
    if rank == 0:
        gen = torch.Generator(device=dev).manual_seed(77)
        orig = U.randn
        U.randn = lambda x, sigma, rng=None: torch.randn(x.shape, dtype=x.dtype, device=x.device,
                                                         generator=gen) * sigma[:, None, None]
        try:
            want = m.enhance(mix, n_steps=4)
        finally:
            U.randn = orig
        err = float((out - want).abs().max())
        ok = out.shape == want.shape and err < 1e-6
    q.put((rank, ok, err, tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("batch", [4, 5])
def test_enhance_sharded_nccl_matches_single_gpu(batch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert all(shape == (batch, 8000) for *_, shape in res), res
