#!/bin/bash
mkdir -p gpurun_out/r2s11; O=gpurun_out/r2s11
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
python tools/correlate_bench.py > $O/correlate.json 2> $O/correlate.err; cat $O/correlate.json; tail -2 $O/correlate.err
timeout 600 python bench.py --draws 500 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_500.json 2> $O/bench_500.err; echo "bench rc=$?"; python -c "
import json; j=json.loads(open('$O/bench_500.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['parity']['max_abs_nats'], j['roofline']['kernel_ms_per_launch'])"
