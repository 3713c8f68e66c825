"""CPU, gloo, world_size = 2: the draw-sharding + all-gather logic of pyipn_b200.distributed (the N > 1 path).

The per-rank compute is injected (the oracle) because there is no GPU here; on the GPU box the same functions run
with the CUDA library and NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyipn_b200.distributed import all_shard_sizes, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 500, 4000, 4097):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = all_shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ipn_oracle as orc
        from pyipn_b200.distributed import global_argmax, loglik_delay_grid_sharded
        from pyipn_b200.synth import delay_grid_workload

        w = delay_grid_workload(T=400, G=9, S=S, dt_bin=0.05)

        def compute(time, exposure, counts, omega, a, b, amp, bkg, delays):
            return orc.loglik_delay_grid(counts, time, exposure, omega, a, b, amp, bkg, delays)

        full = loglik_delay_grid_sharded(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"],
                                         w["bkg"], w["delays"], compute=compute)
        # grid-point sharded argmax of draw 0
        lo, hi = shard_bounds(9, rank, world)
        val, idx = global_argmax(torch.from_numpy(full[lo:hi, 0].copy()), lo)
        np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
        np.save(os.path.join(out_dir, f"argmax_{rank}.npy"), np.array([val, idx]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S", [5, 4])  # ragged and even shards
def test_sharded_delay_grid_gloo_world2(tmp_path, S):
    from oracle import ipn_oracle as orc
    from pyipn_b200.synth import delay_grid_workload

    port = _free_port()
    mp.spawn(_worker, args=(2, port, S, str(tmp_path)), nprocs=2, join=True)
    w = delay_grid_workload(T=400, G=9, S=S, dt_bin=0.05)
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    for r in range(2):
        np.testing.assert_allclose(np.load(tmp_path / f"full_{r}.npy"), exp, rtol=0, atol=1e-9)
        val, idx = np.load(tmp_path / f"argmax_{r}.npy")
        assert int(idx) == int(np.argmax(exp[:, 0])) and val == pytest.approx(exp[:, 0].max())


def _worker_grid(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle as corc
        from pyipn_b200.distributed import (draw_marginal_logsumexp, loglik_delay_grid_sharded, loglik_sky_grid_sharded)
        from pyipn_b200.synth import delay_grid_workload, sky_grid_workload

        corc.set_threads(1)
        # config 2 shape: one draw, delays sharded (7 delays over 2 ranks: ragged)
        w = delay_grid_workload(T=300, G=7, S=1, dt_bin=0.05)

        def compute(time, exposure, counts, omega, a, b, amp, bkg, delays):
            return corc.loglik_delay_grid(counts, time, exposure, omega, a, b, amp, bkg, delays)

        full = loglik_delay_grid_sharded(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                         w["delays"], compute=compute, shard="delays")
        np.save(os.path.join(out_dir, f"delays_{rank}.npy"), full)
        # config 4 shape: pixels sharded, all detectors on every rank
        s = sky_grid_workload(D=3, T=200, S=2, nside=1, dt_bin=0.05)

        def compute_sky(nbins, time, exposure, counts, sc_pos, ghat, omega, a, b, amp, bkg):
            return corc.loglik_sky_grid(counts, time, exposure, nbins, sc_pos, ghat, omega, a, b, amp, bkg)

        sky = loglik_sky_grid_sharded(s["nbins"], s["time"], s["exposure"], s["counts"], s["sc_pos"], s["ghat"], s["omega"],
                                      s["a"], s["b"], s["amp"], s["bkg"], compute=compute_sky)
        np.save(os.path.join(out_dir, f"sky_{rank}.npy"), sky)
        # draw-marginal over draw-sharded columns (5 draws: 3 + 2), including a pixel that is impossible for every draw
        ll = torch.from_numpy(np.random.default_rng(0).normal(size=(6, 5)) * 30.0)
        ll[2] = -np.inf
        from pyipn_b200.distributed import shard_bounds as sb

        lo, hi = sb(5, rank, world)
        np.save(os.path.join(out_dir, f"marg_{rank}.npy"), draw_marginal_logsumexp(ll[:, lo:hi].clone(), 5).numpy())
    finally:
        dist.destroy_process_group()


def test_grid_point_sharding_and_draw_marginal_gloo_world2(tmp_path):
    from oracle import ipn_oracle as orc
    from pyipn_b200.synth import delay_grid_workload, sky_grid_workload
    from scipy.special import logsumexp

    port = _free_port()
    mp.spawn(_worker_grid, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w = delay_grid_workload(T=300, G=7, S=1, dt_bin=0.05)
    exp = orc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    s = sky_grid_workload(D=3, T=200, S=2, nside=1, dt_bin=0.05)
    exp_sky = orc.loglik_sky_grid(s["counts"], s["time"], s["exposure"], s["nbins"], s["sc_pos"], s["ghat"], s["omega"], s["a"],
                                  s["b"], s["amp"], s["bkg"])
    ll = np.random.default_rng(0).normal(size=(6, 5)) * 30.0
    ll[2] = -np.inf
    exp_marg = logsumexp(ll, axis=1) - np.log(5)
    for r in range(2):
        np.testing.assert_allclose(np.load(tmp_path / f"delays_{r}.npy"), exp, rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.load(tmp_path / f"sky_{r}.npy"), exp_sky, rtol=0, atol=1e-9)
        got = np.load(tmp_path / f"marg_{r}.npy")
        assert got[2] == -np.inf
        np.testing.assert_allclose(np.delete(got, 2), np.delete(exp_marg, 2), rtol=1e-13)
