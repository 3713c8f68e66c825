#!/bin/bash
mkdir -p gpurun_out/r2s9; O=gpurun_out/r2s9
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
timeout 600 python bench.py --workload config4 --steps 3 --warmup 2 > $O/bench_config4.json 2> $O/bench_config4.err; echo "bench config4 rc=$?"; python -c "
import json; j=json.loads(open('$O/bench_config4.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['parity'])"
timeout 600 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline --epilogue 3 > $O/bench_comp.json 2> $O/bench_comp.err; echo "bench comp rc=$?"; python -c "
import json; j=json.loads(open('$O/bench_comp.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['parity']['max_abs_nats'])"
