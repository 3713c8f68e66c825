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


def gather_rows(local_block, n_total: int, group=None):
    """All-gather per-rank ``(G_local, S)`` blocks (ragged G_local allowed) into the full ``(n_total, S)``: the
    grid-point-sharded twin of ``gather_columns`` (config 2: delays, config 4: pixels)."""
    return gather_columns(local_block.t().contiguous(), n_total, group=group).t().contiguous()


def loglik_delay_grid_sharded(time, exposure, counts, omega, a, b, amp, bkg, delays, compute=None, device=None, group=None,
                              shard="draws"):
    """Sharded delay grid: every rank returns the full ``(G, S)`` float64 numpy array.

    ``shard="draws"`` (config 3: many posterior draws) gives each rank a slice of the draws, ``shard="delays"`` (config 2:
    one draw, 4096 delays) a slice of the delays.  ``compute(time, exposure, counts, omega, a, b, amp, bkg, delays) ->
    (G_local, S_local)`` defaults to the CUDA library on this rank's GPU; the CPU tests inject their own callable to
    exercise the sharding/gather logic under gloo."""
    import torch
    import torch.distributed as dist

    if shard not in ("draws", "delays"):
        raise ValueError("shard must be 'draws' or 'delays'")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    omega, a, b = np.atleast_2d(omega), np.atleast_2d(a), np.atleast_2d(b)
    amp, bkg, delays = np.atleast_1d(amp), np.atleast_1d(bkg), np.atleast_1d(delays)
    S, G = omega.shape[0], delays.shape[0]
    if compute is None:
        from ._lib import default_context

        compute = default_context().loglik_delay_grid
    if shard == "draws":
        sl = slice(*shard_bounds(S, rank, world))
        local = compute(time, exposure, counts, omega[sl], a[sl], b[sl], amp[sl], bkg[sl], delays)
    else:
        sl = slice(*shard_bounds(G, rank, world))
        local = compute(time, exposure, counts, omega, a, b, amp, bkg, delays[sl])
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    full = gather_columns(t, S, group=group) if shard == "draws" else gather_rows(t, G, group=group)
    return full.cpu().numpy()


def loglik_sky_grid_sharded(nbins, time, exposure, counts, sc_pos, ghat, omega, a, b, amp, bkg, compute=None, device=None,
                            group=None):
    """Pixel-sharded sky grid (config 4): every rank evaluates all detectors for a contiguous slice of the directions, so
    the detector sum stays local; one all-gather returns the full ``(P, S)`` array on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ghat = np.atleast_2d(np.asarray(ghat, dtype=np.float64))
    P = ghat.shape[0]
    if compute is None:
        from ._lib import default_context

        compute = default_context().loglik_sky_grid
    lo, hi = shard_bounds(P, rank, world)
    local = compute(nbins, time, exposure, counts, sc_pos, ghat[lo:hi], omega, a, b, amp, bkg)
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    return gather_rows(t, P, group=group).cpu().numpy()


def draw_marginal_logsumexp(ll_local, n_draws_total: int, group=None):
    """``log (1/S) sum_s exp(ll[g, s])`` over draws that are sharded across ranks: ``ll_local`` is this rank's ``(G, S_local)``
    torch block.  One all-reduce(MAX) and one all-reduce(SUM) of G doubles."""
    import math

    import torch
    import torch.distributed as dist

    if ll_local.shape[1] > 0:
        m = ll_local.max(dim=1).values
    else:
        m = torch.full((ll_local.shape[0],), -math.inf, dtype=ll_local.dtype, device=ll_local.device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    safe = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    acc = torch.exp(ll_local - safe[:, None]).sum(dim=1)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return torch.where(torch.isfinite(m), safe + torch.log(acc) - math.log(n_draws_total), m)


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
