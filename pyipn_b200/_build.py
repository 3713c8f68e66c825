"""In-tree build of libipn_b200.so (nvcc, sm_100a only).  Used by ``__graft_entry__.build()`` and ``make``.

The library is built next to this file (``pyipn_b200/libipn_b200.so``) so that it travels with the repo
snapshot to the GPU box; it is git-ignored.  No JIT, no torch extension machinery: plain ``nvcc -shared``.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# IPN_BUILD_TAG=<tag> (with IPN_NVCC_EXTRA=<flags>): an experimental variant next to the product library
# (libipn_b200_<tag>.so, objects in csrc/build_<tag>/), selected at run time with IPN_B200_LIB=<path>; kernel A/B timing only
TAG = os.environ.get("IPN_BUILD_TAG", "")
LIB = os.path.join(HERE, f"libipn_b200{'_' + TAG if TAG else ''}.so")
# the same library with the experiment / test hooks of the hot kernel compiled in (-DIPN_TC_HOOKS: environment switches that
# re-pace the pipeline or skip work); only tests/ and tools/ load it, the shipped library has none of it
HOOKS_LIB = os.path.join(HERE, f"libipn_b200{'_' + TAG if TAG else ''}_hooks.so")
HOOKS_SOURCES = ["loglik_tc.cu", "rff_eval.cu"]  # the translation units that contain hooks
OBJDIR = os.path.join(CSRC, "build" + ("_" + TAG if TAG else ""))

SOURCES = ["ipn_api.cu", "rff_eval.cu", "loglik_prep.cu", "loglik_simt.cu", "loglik_tc.cu", "poisson.cu", "correlate.cu", "binning.cu", "thinning.cu", "sky_post.cu", "umma_peak.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("IPN_NVCC_EXTRA", "").split()  # e.g. -DIPN_TC_EPI_WARPS=16 -DIPN_TC_GCOLS=8 for kernel experiments


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libipn_b200 cannot be built (there is no CPU fallback)")


def _stamp(sources) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    os.makedirs(OBJDIR, exist_ok=True)
    stamp_file = os.path.join(OBJDIR, "stamp")
    stamp = _stamp(sources)
    if (not force and os.path.exists(LIB) and os.path.exists(HOOKS_LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return LIB
    nvcc = _nvcc()
    log_path = os.path.join(OBJDIR, "ptxas.log")

    def compile_one(job):
        src, hooks = job
        obj = os.path.join(OBJDIR, src.replace(".cu", ".hooks.o" if hooks else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DIPN_TC_HOOKS"] if hooks else []), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, hooks, obj, r

    jobs = [(s, False) for s in sources] + [(s, True) for s in sources if s in HOOKS_SOURCES]
    objs, hook_objs, logs = {}, {}, []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        for src, hooks, obj, r in ex.map(compile_one, jobs):
            logs.append(f"==== {src}{' (-DIPN_TC_HOOKS)' if hooks else ''} ====\n{r.stdout}\n{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError(f"nvcc failed on {src}")
            (hook_objs if hooks else objs)[src] = obj
    with open(log_path, "w") as f:
        f.write("\n".join(logs))
    if verbose:
        sys.stderr.write("\n".join(logs))
    for lib, sel in ((LIB, objs), (HOOKS_LIB, {**objs, **hook_objs})):
        cmd = [nvcc, "-shared", "-o", lib, *[sel[s] for s in sources], "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc link failed")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
