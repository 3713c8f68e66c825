"""GPU parity of the product epilogue (IPN_OPT_HIFI = 4): opt-in, `IPN_TEST_PROD=1 pytest -m gpu tests/test_gpu_prod.py`.

The mode was written after the last GPU session of round 1: its arithmetic is pinned on the CPU (tests/test_emu.py compiles
the epilogue header and restates the count-sorted permutation around it), but the device code of the permutation kernels
and the per-tile dispatch has not run on a B200 yet.  Until it has, these tests stay out of the default `-m gpu` run (the
driver runs that with `-x`) and the library never selects the mode by itself.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle as corc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("IPN_TEST_PROD") != "1", reason="product epilogue: opt-in (IPN_TEST_PROD=1)")]
ATOL = 1e-3


def _prod(ctx, w, delays=None):
    from pyipn_b200._lib import OPT_HIFI

    ctx.set_option(OPT_HIFI, 4)
    try:
        return ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                     w["delays"] if delays is None else delays)
    finally:
        ctx.set_option(OPT_HIFI, 0)


@pytest.mark.parametrize("T,G,S", [(1, 1, 1), (31, 5, 2), (1000 + 17, 130, 3), (4096, 192, 3), (20_000, 257, 2)])
def test_product_epilogue_small_shapes(ctx, T, G, S):
    from pyipn_b200 import synth

    corc.build()
    w = synth.delay_grid_workload(T=T, G=G, S=S, dt_bin=0.01)
    ll = _prod(ctx, w)
    ref = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"], w["delays"])
    assert np.abs(ll - ref).max() < ATOL
    np.testing.assert_array_equal(np.argmax(ll, axis=0), np.argmax(ref, axis=0))


def test_product_epilogue_all_count_classes_and_zero_exposure(ctx):
    """0 ... dozens of counts per bin (five count classes, the per-bin loop), bins of zero exposure, amp = 0, tiny source."""
    from pyipn_b200 import synth

    corc.build()
    w = synth.delay_grid_workload(T=5000 + 3, G=64, S=1, bkg=5000.0, boost=4000.0, dt_bin=0.002)
    assert w["counts"].max() > 8 and (w["counts"] == 0).any()
    for amp in (None, 0.0, 4e-6):
        ww = dict(w) if amp is None else dict(w, amp=np.array([amp]))
        ll = _prod(ctx, ww)
        ref = corc.loglik_delay_grid(ww["counts"], ww["time"], ww["exposure"], ww["omega"], ww["a"], ww["b"], ww["amp"], ww["bkg"],
                                     ww["delays"])
        assert np.abs(ll - ref).max() < ATOL, amp
    ez = w["exposure"].copy()
    ez[::7] = 0.0
    cz = np.where(ez == 0.0, 0, w["counts"])
    wz = dict(w, exposure=ez, counts=cz)
    ll = _prod(ctx, wz)
    ref = corc.loglik_delay_grid(cz, wz["time"], ez, wz["omega"], wz["a"], wz["b"], wz["amp"], wz["bkg"], wz["delays"])
    assert np.abs(ll - ref).max() < ATOL


def test_product_epilogue_config2_full_size_against_emulation_and_oracle(ctx, tmp_path):
    """BASELINE config 2 (4096 delays x 1e5 bins): 48 delays against the oracle and against the host emulation of the same
    pipeline; exact arg-max over the whole grid against the default epilogue."""
    import test_emu as te
    from pyipn_b200 import synth

    corc.build()
    w = synth.delay_grid_workload(T=100_000, G=4096, S=1)
    ll = _prod(ctx, w)[:, 0]
    idx = np.sort(np.random.default_rng(0).choice(4096, 48, replace=False))
    ref = corc.loglik_delay_grid(w["counts"], w["time"], w["exposure"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                 w["delays"][idx])[:, 0]
    assert np.abs(ll[idx] - ref).max() < ATOL
    ll_default = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"], w["bkg"],
                                       w["delays"])[:, 0]
    assert np.argmax(ll) == np.argmax(ll_default)
    if os.path.exists(te.EXE):
        _, emu, _ = te._run_grid(te.EXE, str(tmp_path), w, 0, w["delays"][idx], prod=True)
        assert np.abs(ll[idx] - emu[:, 0]).max() < 3e-4
