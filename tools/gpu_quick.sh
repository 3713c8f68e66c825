#!/bin/bash
mkdir -p gpurun_out/r2s16; O=gpurun_out/r2s16
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.txt
timeout 300 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_500.json 2> $O/bench_500.err; echo "bench rc=$?"
timeout 300 python bench.py --workload config4 --steps 3 --warmup 2 > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"; python -c "import json; j=json.loads(open(\"$O/bench_c4.json\").read().strip().splitlines()[-1]); print(\"config4\", j[\"ms_per_step\"], j[\"parity\"][\"max_abs_nats\"])"
timeout 300 python bench.py --workload rff_eval --steps 5 --warmup 3 > $O/bench_rff.json 2> $O/bench_rff.err; echo "rff rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2s16/bench_500.json").read().strip().splitlines()[-1])
print("bench", "%.4g" % j["value"], "%.1f ms/step" % j["ms_per_step"], "kernel %.2f ms" % j["roofline"]["kernel_ms_per_launch"], "share %.4f" % j["roofline"]["kernel_share_of_step"], j["parity"]["max_abs_nats"])
j = json.loads(open("gpurun_out/r2s16/bench_rff.json").read().strip().splitlines()[-1])
print("rff", "%.4g" % j["value"], "%.3f ms/step" % j["ms_per_step"], j["device_ms_per_call_unflushed"], j["parity"])
PY
