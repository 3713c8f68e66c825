#!/bin/bash
# Round-2 session 2: four Phi digit planes (fixed-point sincos), clamp-free exp2; A/B against -DIPN_TC_PP=3, -DIPN_EPI_CLAMP, 20 epilogue warps
mkdir -p gpurun_out/r2s2; O=gpurun_out/r2s2
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
for v in default pp3 clamp w20; do
  lib=pyipn_b200/libipn_b200_$v.so; [ $v = default ] && lib=pyipn_b200/libipn_b200.so
  IPN_B200_LIB=$PWD/$lib timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for name in ("default", "pp3", "clamp", "w20"):
    try:
        j = json.loads(open(f"gpurun_out/r2s2/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches"]),
              "share %.3f" % j["roofline"]["kernel_share_of_step"], j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
IPN_TC_DEBUG=1 timeout 120 python tools/tc_period.py 64 2 > $O/tc_period.log 2>&1; tail -2 $O/tc_period.log
timeout 120 python tools/emu_vs_device.py > $O/emu_vs_device.log 2>&1; echo "emu_vs_device rc=$?"; tail -1 $O/emu_vs_device.log
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"; cat $O/bench_full.json
