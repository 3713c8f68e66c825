#!/bin/bash
# A/B on one box: the shipped library against variants built with IPN_BUILD_TAG=<tag> IPN_NVCC_EXTRA=... python -m pyipn_b200._build
# (libipn_b200_<tag>.so): il2 / il3 = mixed tiles (-DIPN_TC_INTERLEAVE=2 | 3).  A missing variant prints "no line".
mkdir -p gpurun_out/mixed; O=gpurun_out/mixed
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.txt
for v in default il2 il3 default il2 il3; do  # add further tags (e.g. a build of an older commit) as needed
  if [ $v = default ]; then unset IPN_B200_LIB; else export IPN_B200_LIB=$PWD/pyipn_b200/libipn_b200_$v.so; fi
  timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err; rc=$?
  python - $v $rc <<'PY'
import json, sys
v, rc = sys.argv[1:3]
try:
    j = json.loads(open(f"gpurun_out/mixed/bench_{v}.json").read().strip().splitlines()[-1])
    print(v, "rc", rc, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"].get("kernel_launches_per_step", 0)),
          j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
except Exception as e:
    print(v, "rc", rc, "no line:", e)
PY
done 2>&1 | tee $O/summary.txt
unset IPN_B200_LIB
IPN_TC_DEBUG=1 timeout 120 python tools/tc_period.py 24 0 > $O/period_default.txt 2>&1; tail -3 $O/period_default.txt
