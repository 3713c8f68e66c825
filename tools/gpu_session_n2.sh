#!/bin/bash
mkdir -p gpurun_out/r2n2; O=gpurun_out/r2n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; cut -c1-400 $O/bench_n2.json
python -c "
import json; j=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['parity'])"
timeout 600 $TR bench.py --gpus 2 --workload config4 --steps 3 --warmup 2 > $O/bench_n2_config4.json 2> $O/bench_n2_config4.err; echo "config4 rc=$?"
python -c "
import json; j=json.loads(open('$O/bench_n2_config4.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['parity'])"
for f in $O/*.err; do tail -2 $f; done
