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


def test_integration_guide_names_every_entry_point_and_option():
    """INTEGRATION.md is the maintainer's map from reference call sites to entry points: every function of the header and
    every IPN_OPT_* name appears there."""
    guide = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    header = open(os.path.join(ROOT, "include", "ipn.h")).read()
    names = declared_symbols() + sorted(set(re.findall(r"\b(IPN_OPT_[A-Z0-9_]+)\b", header)))
    assert [n for n in names if n not in guide] == []


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


def test_shipped_library_has_no_experiment_hooks(lib):
    """The pacing / work-skipping switches of the hot kernel (IPN_TC_*, IPN_RFF_FP64) exist only in the -DIPN_TC_HOOKS build
    that tests/ and tools/ load explicitly: the product library must not even contain their names."""
    from pyipn_b200._lib import HOOKS_LIB_PATH, LIB_PATH

    data = open(LIB_PATH, "rb").read()
    assert b"IPN_TC_" not in data and b"IPN_RFF_FP64" not in data
    hooks = open(HOOKS_LIB_PATH, "rb").read()
    assert b"IPN_TC_SKIP_MATH" in hooks and b"IPN_TC_GRID" in hooks
    hl = ctypes.CDLL(HOOKS_LIB_PATH)
    for name in declared_symbols():
        assert hasattr(hl, name), name


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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/ipn.h must be usable from C (the boundary is a C ABI, not C++): compile a C99 program against it, link it
    with the library and call the entry points that need no GPU."""
    import shutil
    import subprocess

    from pyipn_b200._build import build_library

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    LIB_PATH = build_library()
    root = ROOT
    src = tmp_path / "abi_c.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "ipn.h"
int main(void) {
    const char* v = ipn_version();
    if (!v || !*v) return 2;
    /* argument errors are reported before any CUDA call */
    int rc = ipn_rff_eval(NULL, 4, NULL, 1, 2, NULL, NULL, NULL, NULL, NULL, NULL);
    if (rc != IPN_ERR_INVALID_ARGUMENT) return 3;
    if (strlen(ipn_last_error()) == 0) return 4;
    ipn_ctx* ctx = NULL;
    rc = ipn_ctx_create(0, &ctx);
    if (rc == IPN_OK) { ipn_ctx_destroy(ctx); printf("gpu "); }
    else if (rc != IPN_ERR_NO_DEVICE && rc != IPN_ERR_CUDA) return 5;
    printf("ok %s\n", v);
    return 0;
}
''')
    exe = tmp_path / "abi_c"
    libdir = os.path.dirname(LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lipn_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
