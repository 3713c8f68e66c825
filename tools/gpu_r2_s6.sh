#!/bin/bash
# Round-2 session 6: epilogue-warp variants A/B, then the ncu evidence of the shipped build (launch list, hot kernel, FP32-pipe kernel)
mkdir -p gpurun_out/r2s6; O=gpurun_out/r2s6
for v in default w12 g8; do
  lib=pyipn_b200/libipn_b200_$v.so; [ $v = default ] && lib=pyipn_b200/libipn_b200.so
  IPN_B200_LIB=$PWD/$lib timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for name in ("default", "w12", "g8"):
    try:
        j = json.loads(open(f"gpurun_out/r2s6/bench_{name}.json").read().strip().splitlines()[-1])
        print(name, "%.4g units/s" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms x %d" % (j["roofline"]["kernel_ms_per_launch"], j["roofline"]["kernel_launches"]),
              j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "parity", j["parity"]["max_abs_nats"], j["parity"]["ok"])
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY
CMD="python bench.py --draws 500 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain.json 2> $O/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > $O/plain2.json 2> $O/plain2.err && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:loglik_tc_kernelILi7ELi0 -s 2 -c 1 -o $O/prof_tc $CMD > $O/ncu_full.log 2>&1
echo "ncu tc rc=$?"
python tools/ncu_target.py 1 16 > $O/simt_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:loglik_simt_kernel -s 1 -c 1 -o $O/prof_simt python tools/ncu_target.py 1 16 > $O/ncu_simt.log 2>&1
echo "ncu simt rc=$?"
ls -la $O
