#!/bin/bash
# headline bench at N = 8 and N = 1 on the same box (strong scaling), shipped build
mkdir -p gpurun_out/session_n8b; O=gpurun_out/session_n8b
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n8.json 2> $O/bench_n8.err; echo "n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for n in (8, 4, 1):
    j = json.loads(open(f"gpurun_out/session_n8b/bench_n{n}.json").read().strip().splitlines()[-1])
    print(n, "%.4g" % j["value"], "%.1f ms" % j["ms_per_step"], "e2e %.4g" % j["e2e"]["value"], j["clocks"]["sm_mhz"], j["parity"]["max_abs_nats"], j["parity"]["ok"])
PY
