import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def rff_golden():
    return dict(np.load(os.path.join(GOLDEN, "rff_golden.npz")))


@pytest.fixture(scope="session")
def loglik_golden():
    return dict(np.load(os.path.join(GOLDEN, "loglik_golden.npz")))


@pytest.fixture(scope="session")
def ctx():
    """libipn_b200 context on cuda:0.  No skip-on-failure: on the GPU box a missing library must FAIL the run."""
    from pyipn_b200 import Context

    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def hooks_ctx():
    """Context on the -DIPN_TC_HOOKS build of the library (libipn_b200_hooks.so): the only build in which the pacing /
    work-skipping environment switches of the hot kernel exist."""
    from pyipn_b200 import Context
    from pyipn_b200._lib import HOOKS_LIB_PATH

    c = Context(0, lib_path=HOOKS_LIB_PATH)
    yield c
    c.close()
