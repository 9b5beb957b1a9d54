#!/usr/bin/env python3
"""One line per captured kernel from `ncu --page raw --csv` exports: time, DRAM traffic, issue / ALU / instruction-cache load.
usage: ncu_table.py a.csv [b.csv ...]"""
import csv, sys
COLS = [("ms", "gpu__time_duration.sum", 1.0), ("GB", None, 1.0), ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0), ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
        ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0), ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
        ("icache%", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", 1.0),
        ("warps", "smsp__warps_active.avg.per_cycle_active", 1.0), ("regs", "launch__registers_per_thread", 1.0),
        ("Minst", "smsp__inst_executed.sum", 1e-6)]
print("%-46s " % "kernel" + " ".join("%7s" % c[0] for c in COLS) + "  top stalls")
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# " + f)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("zb::", "")[:46]
        vals = []
        for label, key, scale in COLS:
            if key is None:
                rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
                u = units[ix["dram__bytes_read.sum"]]
                m = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]
                u2 = units[ix["dram__bytes_write.sum"]]
                m2 = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u2]
                vals.append(rd * m + wr * m2)
            else:
                v = float(r[ix[key]]) * scale
                if label == "ms":
                    v *= {"msecond": 1.0, "ms": 1.0, "usecond": 1e-3, "us": 1e-3, "second": 1e3, "s": 1e3, "nsecond": 1e-6, "ns": 1e-6}[units[ix[key]]]
                vals.append(v)
        stalls = sorted(((float(r[ix[h]]), h.split("stalled_")[1].split("_per")[0]) for h in hdr
                         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
        print("%-46s " % name + " ".join("%7.2f" % v for v in vals) + "  " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4] if n != "selected"))
