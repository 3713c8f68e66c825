#!/bin/bash
mkdir -p gpurun_out/workloads; O=gpurun_out/workloads
for wl in rff_eval config2 config4; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 $O/bench_$wl.err; cat $O/bench_$wl.json
done
timeout 900 python bench.py --workload config5 --grbs 40 --steps 1 --warmup 1 > $O/bench_config5.json 2> $O/bench_config5.err; echo "bench config5 rc=$?"; tail -2 $O/bench_config5.err; cat $O/bench_config5.json
