#!/bin/bash
mkdir -p gpurun_out/r2s14; O=gpurun_out/r2s14
for v in default w20g8 w24g8 w28g8; do
  lib=pyipn_b200/libipn_b200_$v.so; [ $v = default ] && lib=pyipn_b200/libipn_b200.so
  IPN_B200_LIB=$PWD/$lib timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"; tail -1 $O/bench_$v.err
done
python - <<'PY'
import json
for name in ("default", "w20g8", "w24g8", "w28g8"):
    try:
        j = json.loads(open(f"gpurun_out/r2s14/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches"]),
              j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
