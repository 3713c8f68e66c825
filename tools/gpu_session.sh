#!/bin/bash
# Evidence of the shipped build: GPU suite, smoke, bench (N = 1, full grid), reference arm, ncu launch list + captures
mkdir -p gpurun_out/session; O=gpurun_out/session
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt; tail -4 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cat $O/bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; cat $O/bench_ref.json
CMD="python bench.py --draws 500 --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > $O/plain.json 2> $O/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > $O/plain2.json 2> $O/plain2.err && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:loglik_tc_kernelILi7ELi1 -s 2 -c 1 -o $O/prof_tc $CMD > $O/ncu_full.log 2>&1
echo "ncu tc rc=$?"
python tools/ncu_target.py 1 16 > $O/simt_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:loglik_simt_kernel -s 1 -c 1 -o $O/prof_simt python tools/ncu_target.py 1 16 > $O/ncu_simt.log 2>&1
echo "ncu simt rc=$?"
cat > /tmp/rff_t.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from pyipn_b200 import Context
from pyipn_b200._lib import OPT_RFF_IMPL
from pyipn_b200.synth import delay_grid_workload
S, T = 1000, 100_000
w = delay_grid_workload(T=T, G=8, S=S)
J = w["omega"].shape[1]
ctx = Context(0)
dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dict(time=w["time"], omega=w["omega"], a=w["a"], b=w["b"], shift=np.full(S, 0.1234), amp=w["amp"]).items()}
out = torch.empty((S, T), dtype=torch.float64, device="cuda")
ctx.set_option(OPT_RFF_IMPL, 2)
for _ in range(3):
    ctx.rff_eval_dev(T, dev["time"].data_ptr(), S, J, dev["omega"].data_ptr(), dev["a"].data_ptr(), dev["b"].data_ptr(), dev["shift"].data_ptr(), dev["amp"].data_ptr(), out.data_ptr())
    ctx.synchronize()
print("ok", ctx.last_kernel_ms)
PY
python /tmp/rff_t.py > $O/rff_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_rff.csv python /tmp/rff_t.py > $O/ncu_rff_list.log 2>&1
echo "ncu rff list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:loglik_tc_kernelILi7ELi3 -s 2 -c 1 -o $O/prof_rff python /tmp/rff_t.py > $O/ncu_rff.log 2>&1
echo "ncu rff rc=$?"
python tools/correlate_bench.py > $O/correlate.json 2>/dev/null; cat $O/correlate.json
