#!/bin/bash
mkdir -p gpurun_out/r2s7; O=gpurun_out/r2s7
timeout 600 python -m pytest tests/test_gpu_rff.py -m gpu -q -x > $O/pytest_rff.txt 2>&1; echo "pytest rff rc=$?"; tail -15 $O/pytest_rff.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
python - <<'PY' > $O/rff_rate.txt 2>&1
import time, numpy as np, torch
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_RFF_IMPL
from pyipn_b200.synth import delay_grid_workload
ctx = Context(0)
for S in (64, 1000, 4000):
    w = delay_grid_workload(T=100_000, G=8, S=S)
    T, J = 100_000, w["omega"].shape[1]
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(time=w["time"], omega=w["omega"], a=w["a"], b=w["b"], shift=np.full(S, 0.1234), amp=w["amp"]).items()}
    out = torch.empty((S, T), dtype=torch.float64, device="cuda")
    for impl in (2, 1):
        if impl == 1 and S > 1000: continue
        ctx.set_option(OPT_RFF_IMPL, impl)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.rff_eval_dev(T, dev["time"].data_ptr(), S, J, dev["omega"].data_ptr(), dev["a"].data_ptr(), dev["b"].data_ptr(), dev["shift"].data_ptr(), dev["amp"].data_ptr(), out.data_ptr())
            ctx.synchronize(); dt = time.perf_counter() - t0
        print("S", S, "impl", impl, "%.3f ms" % (dt * 1e3), "%.3g (draw*bin)/s" % (S * T / dt), "kernel ms", ctx.last_kernel_ms, ctx.last_main_kernel, flush=True)
PY
cat $O/rff_rate.txt
