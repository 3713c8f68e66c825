"""CPU: the reference arm of bench.py prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--bins", "2000", "--delays", "64", "--draws-per-gpu", "4"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--bins", "2000", "--delays", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_parity_check_covers_every_batch_and_catches_an_error():
    """bench.py's oracle check of the timed result: one draw of every internal draw batch of every rank is picked, an exact
    result passes, and an error of 2e-3 nats at any picked pair fails the run (the function is pure host code)."""
    import types

    import numpy as np

    sys.path.insert(0, ROOT)
    import bench
    from oracle import c_oracle as corc
    from pyipn_b200.distributed import shard_bounds
    from pyipn_b200.synth import delay_grid_workload

    S, world, batch = 23, 2, 3
    w = delay_grid_workload(T=1500, G=48, S=S, dt_bin=0.01)
    full = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    a = types.SimpleNamespace(parity_draws=64)
    p = bench.parity_check(a, w, full.copy(), S, world, batch)
    assert p["ok"] and p["max_abs_nats"] < 1e-9 and p["argmax_equal"]
    n_batches = sum(-(-(shard_bounds(S, r, world)[1] - shard_bounds(S, r, world)[0]) // batch) for r in range(world))
    assert p["draw_batches_of_the_library"] == n_batches and p["draws_checked"] >= n_batches
    # every batch of every rank holds at least one checked draw: an error in ANY batch is seen
    gsel = p["delays_checked"]
    for r in range(world):
        lo, hi = shard_bounds(S, r, world)
        for b0 in range(lo, hi, batch):
            bad = full.copy()
            bad[gsel[0], b0:min(hi, b0 + batch)] += 2e-3
            q = bench.parity_check(a, w, bad, S, world, batch)
            assert not q["ok"] and q["max_abs_nats"] > 1.9e-3, (r, b0)
