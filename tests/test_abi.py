"""CPU: the C-ABI library builds, loads, and exports every symbol include/ipn.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ipn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pyipn_b200._build import build_library

    return ctypes.CDLL(build_library())


def test_header_and_binding_agree():
    from pyipn_b200 import _lib

    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_string(lib):
    lib.ipn_version.restype = ctypes.c_char_p
    lib.ipn_last_error.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ipn_version()
    assert isinstance(lib.ipn_last_error(), bytes)


def test_library_contains_only_sm100a_code():
    import subprocess

    from pyipn_b200._lib import LIB_PATH

    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a GPU context creation must fail loudly with IPN_ERR_NO_DEVICE, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU-only box")
    h = ctypes.c_void_p()
    rc = lib.ipn_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and not h.value
    lib.ipn_last_error.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.ipn_last_error()
    from pyipn_b200 import IpnError, RFF
    import numpy as np

    with pytest.raises(IpnError):
        RFF(np.arange(4.0), np.ones(2), np.ones(2), np.ones(2), np.ones(2), 1.0, 1.0)
