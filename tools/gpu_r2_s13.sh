#!/bin/bash
# Round-2 session 13: tiles with and without counts interleaved (C Z Z C) against the natural order (-DIPN_TC_INTERLEAVE=0)
mkdir -p gpurun_out/r2s13; O=gpurun_out/r2s13
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
for v in default nointer; do
  lib=pyipn_b200/libipn_b200_$v.so; [ $v = default ] && lib=pyipn_b200/libipn_b200.so
  for e in 0 2; do
    IPN_B200_LIB=$PWD/$lib timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline --epilogue $e > $O/bench_${v}_e$e.json 2> $O/bench_${v}_e$e.err; echo "bench $v epilogue $e rc=$?"
  done
done
python - <<'PY'
import json
for name in ("default_e0", "nointer_e0", "default_e2", "nointer_e2"):
    try:
        j = json.loads(open(f"gpurun_out/r2s13/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches"]),
              j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
