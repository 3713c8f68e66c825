#!/bin/bash
# Round-2 session 3: fifth W digit plane + level-6 accumulator (12 products)
mkdir -p gpurun_out/r2s3; O=gpurun_out/r2s3
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_500.json 2> $O/bench_500.err; echo "bench 500 rc=$?"
python - <<'PY'
import json
for name in ("500",):
    try:
        j = json.loads(open(f"gpurun_out/r2s3/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches"]),
              "share %.3f" % j["roofline"]["kernel_share_of_step"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
tail -3 $O/bench_500.err
IPN_TC_DEBUG=1 timeout 120 python tools/tc_period.py 64 2 > $O/tc_period.log 2>&1; tail -2 $O/tc_period.log
timeout 120 python tools/emu_vs_device.py > $O/emu_vs_device.log 2>&1; echo "emu_vs_device rc=$?"; tail -1 $O/emu_vs_device.log
timeout 600 python bench.py --parity-draws 128 > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"; cat $O/bench_full.json; tail -2 $O/bench_full.err
