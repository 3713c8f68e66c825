"""Multi-GPU layer: the (delay | sky) x draw grid shards with no data-path collective.

One process per GPU (torchrun).  Every rank holds a replica of the bins/counts (a few MB) and evaluates a
contiguous slice of the posterior draws (config 3), of the grid points (config 2/4) or of the GRBs (config 5);
the only communication is one ``all_gather`` of the per-rank float64 log-likelihood blocks at the end
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Bandwidth is irrelevant at these sizes (16 MB per rank for
4096 x 500 draws); what matters for scaling is equal work per rank and not re-uploading data per call.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of ``n`` items for ``rank`` (first ``n % world`` ranks get one extra)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_shard_sizes(n: int, world: int):
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def gather_columns(local_block, n_total: int, group=None):
    """All-gather per-rank ``(G, S_local)`` torch blocks (ragged S_local allowed) into the full ``(G, n_total)``.

    Works for CUDA tensors (NCCL) and CPU tensors (gloo).  Ragged shards are padded to the largest shard so a
    single fixed-size ``all_gather_into_tensor`` suffices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = all_shard_sizes(n_total, world)
    G = local_block.shape[0]
    mx = max(sizes)
    send = local_block
    if local_block.shape[1] != mx:
        send = torch.zeros((G, mx), dtype=local_block.dtype, device=local_block.device)
        send[:, : local_block.shape[1]] = local_block
    send = send.contiguous()
    recv = torch.empty((world, G, mx), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return torch.cat([recv[r, :, : sizes[r]] for r in range(world)], dim=1)


def loglik_delay_grid_sharded(time, exposure, counts, omega, a, b, amp, bkg, delays, compute=None, device=None, group=None):
    """Draw-sharded delay grid: every rank returns the full ``(G, S)`` float64 numpy array.

    ``compute(time, exposure, counts, omega, a, b, amp, bkg, delays) -> (G, S_local)`` defaults to the CUDA library on
    this rank's GPU; the CPU tests inject their own callable to exercise the sharding/gather logic under gloo."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = np.atleast_2d(omega).shape[0]
    lo, hi = shard_bounds(S, rank, world)
    if compute is None:
        from ._lib import default_context

        compute = default_context().loglik_delay_grid
    sl = slice(lo, hi)
    local = compute(time, exposure, counts, np.atleast_2d(omega)[sl], np.atleast_2d(a)[sl], np.atleast_2d(b)[sl],
                    np.atleast_1d(amp)[sl], np.atleast_1d(bkg)[sl], delays)
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    return gather_columns(t, S, group=group).cpu().numpy()


def global_argmax(ll_local, lo: int, group=None):
    """Global (value, flat grid index) of the maximum over a grid-point-sharded vector via one all_gather."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    i = int(torch.argmax(ll_local))
    pair = torch.tensor([float(ll_local[i]), float(lo + i)], dtype=torch.float64, device=ll_local.device)
    out = torch.empty((world, 2), dtype=torch.float64, device=ll_local.device)
    dist.all_gather_into_tensor(out.view(-1), pair, group=group)
    r = int(torch.argmax(out[:, 0]))
    return float(out[r, 0]), int(out[r, 1])
