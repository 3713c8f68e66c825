#!/bin/bash
# First GPU contact of the product epilogue (IPN_OPT_HIFI = 4): parity, device vs emulation,
# then the A/B bench against the default epilogue and one ncu capture of the new instantiation.  Outputs -> gpurun_out/
# Every step runs under its own timeout: the mode's device code had not run on a B200 when this script was written.
mkdir -p gpurun_out
IPN_TEST_PROD=1 timeout 300 python -m pytest tests/test_gpu_prod.py -m gpu -x -q > gpurun_out/pytest_prod.txt 2>&1; echo "pytest prod rc=$?"
tail -5 gpurun_out/pytest_prod.txt
timeout 120 python tools/emu_vs_device.py --prod > gpurun_out/emu_vs_device_prod.log 2>&1; echo "emu_vs_device rc=$?"; tail -1 gpurun_out/emu_vs_device_prod.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --no-cpu-baseline --epilogue 4 > gpurun_out/bench_prod.json 2> gpurun_out/bench_prod.err; echo "bench prod rc=$?"
python - <<'PY'
import json
for name in ("default", "prod"):
    try:
        j = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms" % j["roofline"]["kernel_ms_per_launch"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
for m in 2 4; do IPN_TC_DEBUG=1 timeout 120 python tools/tc_period.py 64 $m > gpurun_out/tc_period_$m.log 2>&1; tail -2 gpurun_out/tc_period_$m.log; done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --epilogue 4"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:loglik_tc_kernelILi7ELi3 -s 2 -c 1 \
    -o gpurun_out/prof_prod $CMD > gpurun_out/ncu_prod.log 2>&1
echo "ncu prod rc=$?"
