#!/bin/bash
# Round-2 multi-GPU session (gpurun --gpus 8): strong and weak scaling of the headline bench, pixel-sharded config 4, GRB-sharded config 5
mkdir -p gpurun_out/session_n8; O=gpurun_out/session_n8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8_strong.json 2> $O/bench_n8_strong.err; echo "strong rc=$?"; cat $O/bench_n8_strong.json | cut -c1-600
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 --draws-per-gpu 500 --no-cpu-baseline > $O/bench_n8_weak.json 2> $O/bench_n8_weak.err; echo "weak rc=$?"; cat $O/bench_n8_weak.json | cut -c1-300
timeout 600 $TR bench.py --gpus 8 --workload config4 --steps 5 --warmup 3 > $O/bench_n8_config4.json 2> $O/bench_n8_config4.err; echo "config4 rc=$?"; cat $O/bench_n8_config4.json | cut -c1-300
timeout 900 $TR bench.py --gpus 8 --workload config5 --grbs 1000 --steps 1 --warmup 1 > $O/bench_n8_config5.json 2> $O/bench_n8_config5.err; echo "config5 rc=$?"; cat $O/bench_n8_config5.json | cut -c1-300
timeout 600 $TR bench.py --gpus 8 --workload rff_eval --draws 8000 --steps 4 --warmup 3 > $O/bench_n8_rff.json 2> $O/bench_n8_rff.err; echo "rff rc=$?"; cat $O/bench_n8_rff.json | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_n1_same_box.json 2> $O/bench_n1.err; echo "n1 rc=$?"; cat $O/bench_n1_same_box.json | cut -c1-300
for f in $O/*.err; do tail -2 $f; done
