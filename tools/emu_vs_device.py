"""Device run against the host emulation of its own arithmetic (tests/emu/tc_emu.cpp) at BASELINE config 2's full size.

For 48 of the 4096 delays: ll from libipn_b200 with each epilogue forced in turn (IPN_OPT_HIFI = 2 plain fp32,
3 compensated fp32, 1 fp64), the emulation of the same three epilogues, and the float64 C oracle.  What differs between
device and emulation by construction: the real MUFU.LG2 against its stand-in, FFMA contractions nvcc adds, and the order
of the float64 atomics -- all far below the distance of either from the oracle.  Writes one JSON line
(gpurun_out/emu_vs_device.json when that directory exists).  Needs a B200; takes a few seconds.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_emu as te  # the emulator's build recipe and its input format
    from pyipn_b200 import Context, synth
    from pyipn_b200._lib import OPT_HIFI

    os.makedirs(te.BUILD, exist_ok=True)
    if not os.path.exists(te.EXE):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", te.EXE, te.SRC], check=True)
    exe = te.EXE
    from oracle import c_oracle

    c_oracle.build()
    w = synth.delay_grid_workload(T=100_000, G=4096, S=1)
    idx = np.sort(np.random.default_rng(0).choice(4096, 48, replace=False))
    res, emu, ref = te._run_grid(exe, tempfile.mkdtemp(), w, 0, w["delays"][idx])
    dev = np.empty_like(emu)
    with Context(0) as ctx:
        for col, opt in ((0, 2), (1, 3), (2, 1)):
            ctx.set_option(OPT_HIFI, opt)
            dev[:, col] = ctx.loglik_delay_grid(w["time"], w["exposure"], w["counts"], w["omega"], w["a"], w["b"], w["amp"],
                                                w["bkg"], w["delays"])[idx, 0]
    names = ["fp32", "compensated fp32", "fp64"]
    out = {"workload": "config2: 4096 delays x 1e5 bins, K=50, 1 draw; 48 delays compared", "epilogues": names}
    for name, x, y in (("device_minus_emulation", dev, emu), ("device_minus_oracle", dev, ref[:, None]),
                       ("emulation_minus_oracle", emu, ref[:, None])):
        d = x - y
        out[name] = {"mean": d.mean(axis=0).tolist(), "rms": np.sqrt((d ** 2).mean(axis=0)).tolist(),
                     "max_abs": np.abs(d).max(axis=0).tolist()}
    line = json.dumps(out)
    print(line)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "emu_vs_device.json"), "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
