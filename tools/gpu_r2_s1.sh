#!/bin/bash
# Round-2 session 1: GPU suite at HEAD (hooks build), first device run of the product epilogue, A/B bench.  Outputs -> gpurun_out/
mkdir -p gpurun_out/r2s1; O=gpurun_out/r2s1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
IPN_TEST_PROD=1 timeout 300 python -m pytest tests/test_gpu_prod.py -m gpu -q > $O/pytest_prod.txt 2>&1; echo "pytest prod rc=$?"
tail -8 $O/pytest_prod.txt
timeout 120 python tools/emu_vs_device.py --prod > $O/emu_vs_device_prod.log 2>&1; echo "emu_vs_device rc=$?"; tail -2 $O/emu_vs_device_prod.log
timeout 300 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --no-cpu-baseline --epilogue 4 > $O/bench_prod.json 2> $O/bench_prod.err; echo "bench prod rc=$?"
python - <<'PY'
import json
for name in ("default", "prod"):
    try:
        j = json.loads(open(f"gpurun_out/r2s1/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms" % j["roofline"]["kernel_ms_per_launch"], j["clocks"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
for m in 2 4; do IPN_TC_DEBUG=1 timeout 120 python tools/tc_period.py 64 $m > $O/tc_period_$m.log 2>&1; tail -2 $O/tc_period_$m.log; done
python - <<'PY'
from pyipn_b200 import Context
c = Context(0)
print("microbench fma TF/s", c.microbench(0), "mufu", c.microbench(1), "umma int8 TOP/s", c.microbench(2))
PY
