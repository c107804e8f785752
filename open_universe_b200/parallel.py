"""Batch-sharded ``enhance()`` across the GPUs of one box.

Every batch row is independent end to end (per-clip normalisation, per-sequence GRU, per-clip peak
limiter), so the path shards with NO data-path collective: a contiguous split of the batch over
the ranks, weights replicated, and one ``all_gather`` of the (B/R, T) fp32 result at the end
(16 MB per rank at BASELINE cfg-5) over NCCL / NVLink.  Uneven batches are padded to ceil(B/R)
rows per rank and trimmed after the gather.

Fixed-noise parity with a single-GPU run: each rank draws the GLOBAL (B, 1, T_pad) noise with the
same generator seed and keeps its own rows (``global_noise=True``); a production run can instead
let every rank draw only its rows from a rank-offset seed.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """Contiguous split; every rank gets ceil(batch / world) rows, the tail ranks fewer / none."""
    per = -(-batch // world_size)
    lo = min(rank * per, batch)
    hi = min(lo + per, batch)
    return lo, hi, per


def gather_rows(local, per, batch, group=None):
    """all_gather equally sized (per, T) blocks and trim the padding rows."""
    world = dist.get_world_size(group)
    if local.shape[0] < per:
        pad = local.new_zeros((per - local.shape[0],) + tuple(local.shape[1:]))
        local = torch.cat([local, pad], dim=0)
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:batch]


def enhance_sharded(model, mix, group=None, seed=None, global_noise=True, enhance_fn=None, **kwargs):
    """``model.enhance(mix)`` for a (B, T) batch, rows split across the ranks of ``group``.
    Every rank passes the same full ``mix`` and receives the full (B, T) result."""
    if mix.ndim != 2:
        raise ValueError("enhance_sharded expects a (B, T) batch")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    batch = mix.shape[0]
    lo, hi, per = shard_bounds(batch, world, rank)
    fn = enhance_fn if enhance_fn is not None else model.enhance
    if hi > lo:
        local_mix = mix[lo:hi]
        if seed is not None and global_noise:
            from .networks.universe import universe as U
            gen = torch.Generator(device=mix.device).manual_seed(seed)
            orig = U.randn

            def sliced_randn(x, sigma, rng=None):
                # enhance() always passes its own ``rng`` (None here): the seeded generator is
                # closed over, not a default argument the caller would override
                full = torch.randn((batch,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device,
                                   generator=gen)
                return full[lo:hi] * sigma[:, None, None]

            U.randn = sliced_randn
            try:
                local = fn(local_mix, **kwargs)
            finally:
                U.randn = orig
        else:
            if seed is not None:
                kwargs = dict(kwargs, rng=torch.Generator(device=mix.device).manual_seed(seed + rank))
            local = fn(local_mix, **kwargs)
    else:
        local = mix.new_zeros((0, mix.shape[1]))
    if world == 1:
        return local
    return gather_rows(local, per, batch, group)
