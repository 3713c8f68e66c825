"""Turn the raw outputs of a GPU session (gpurun_out/[<subdir>/]) into the tracked summaries under profiles/.

    python tools/refresh_profiles.py r02 r2s6      # tag, sub-directory of gpurun_out/

writes profiles/<tag>_bench_n1.json, <tag>_bench_reference_n1.json, <tag>_launches.csv, <tag>_launches_summary.md,
<tag>_ncu_loglik_tc_kernel.txt and profiles/roofline_latest.json (the DRAM traffic per launch that bench.py reports).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__inst_executed.sum",
]


def launches(tag):
    src = os.path.join(OUT, "launches.csv")
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    k_name, k_val = head.index("Kernel Name"), head.index("Metric Value")
    k_unit = head.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows:
        v = float(r[k_val].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[k_unit], 1e-3)
        n, t = tot.get(r[k_name], (0, 0.0))
        tot[r[k_name]] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    shutil.copy(src, os.path.join(PROF, f"{tag}_launches.csv"))
    with open(os.path.join(PROF, f"{tag}_launches_summary.md"), "w") as f:
        f.write("# ncu launch list of `python bench.py --draws 500 --steps 2 --warmup 1 --no-cpu-baseline` (B200, --clock-control none)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            short = name.split("(")[0].replace("void ", "")
            f.write(f"| `{short}` | {n} | {t:.1f} | {100 * t / total:.2f} % |\n")
    return tot


def ncu(tag, rep_name="prof_bench.ncu-rep", label="loglik_tc_kernel", what=None, roofline=True):
    rep = os.path.join(OUT, rep_name)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units, vals = rows[0], rows[1], rows[2]
    rec = {h: (u, v) for h, u, v in zip(head, units, vals)}
    with open(os.path.join(PROF, f"{tag}_ncu_{label}.txt"), "w") as f:
        f.write(what or ("# ncu --set full --clock-control none --import-source on capture of the hot kernel inside `python bench.py --draws 500 --steps 2 "
                         "--warmup 1 --no-cpu-baseline`\n# (third launch of the compensated-epilogue instantiation, the one the bench data get; one launch = one draw batch of the bench step)\n\n"))
        for k in ("Kernel Name", "Grid Size", "Block Size"):
            f.write(f"{k} [] = {rec[k][1]}\n")
        for k in KEEP:
            if k in rec and rec[k][1] not in ("", "nan", "n/a"):
                f.write(f"{k} [{rec[k][0]}] = {rec[k][1]}\n")

    def to_bytes(k):
        u, v = rec[k]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    if not roofline:
        return traffic
    json.dump({"kernel": "loglik_tc_kernel", "traffic_bytes_per_launch": traffic,
               "source": f"profiles/{tag}_ncu_{label}.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch = one draw "
                         "batch of the bench step)"}, open(os.path.join(PROF, "roofline_latest.json"), "w"), indent=1)
    return traffic


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    for src, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("bench_ref.json", f"{tag}_bench_reference_n1.json")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    tot = launches(tag)
    print("kernels:", len(tot))
    rep = "prof_tc.ncu-rep" if os.path.exists(os.path.join(OUT, "prof_tc.ncu-rep")) else "prof_bench.ncu-rep"
    print("traffic bytes/launch:", ncu(tag, rep))
    if os.path.exists(os.path.join(OUT, "prof_simt.ncu-rep")):
        ncu(tag, "prof_simt.ncu-rep", "loglik_simt_kernel",
            "# ncu --set full --clock-control none capture of the FP32-pipe kernel (IPN_OPT_LOGLIK_IMPL = 1) inside `python tools/ncu_target.py 1 16`\n"
            "# (4096 delays x 16 draws x 1e5 bins, hi / lo W table): the CUDA-core baseline the tcgen05 kernel is measured against\n\n", roofline=False)
