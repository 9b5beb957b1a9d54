#!/usr/bin/env python3
"""Summarise an ncu report into a small markdown file for profiles/ (key metrics of each captured
kernel, stall mix, hottest source lines when the cubin is available).
usage: ncu_summary.py <report.ncu-rep> <out.md> [nwarps]"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "gcc__cache_requests_type_instruction.sum", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active"]
with open(out, "w") as f:
    f.write("# ncu summary of `%s`\n\n(captured with `ncu --set full --clock-control none --import-source on`; times under the profiler are not bench values)\n\n" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name[:120])
        for k in KEYS:
            if k in idx:
                f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
        rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
        f.write("| traffic (dram read+write) | %.4f | %s |\n" % (rd + wr, units[idx["dram__bytes_read.sum"]]))
        stalls = sorted(((float(r[idx[h]]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
        f.write("\nstall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (h.split("stalled_")[1].split("_per")[0], v) for v, h in stalls[:7]) + "\n\n")
print("wrote", out)
