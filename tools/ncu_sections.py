#!/usr/bin/env python3
"""Per-instruction executed counts and stall samples from an ncu report, bucketed by instruction-index ranges.
usage: ncu_sections.py report.ncu-rep [bucket_size]   (prints warp-level instructions per warp-block)"""
import csv, subprocess, sys
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 250
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
first = float(body[0][ix["Instructions Executed"]])
tot = 0
acc = []
for i, r in enumerate(body):
    n = float(r[ix["Instructions Executed"]]) / first
    s = float(r[ix["# Samples"]])
    acc.append((i, r[ix["Source"]].strip(), n, s))
tots = sum(a[2] for a in acc); samp = sum(a[3] for a in acc)
print("total warp-instr per warp: %.0f; samples %d" % (tots, samp))
for b in range(0, len(acc), bucket):
    seg = acc[b:b + bucket]
    print("%5d-%5d  instr %7.0f (%4.1f%%)  samples %5.1f%%   %s" % (b, b + len(seg) - 1, sum(a[2] for a in seg), 100 * sum(a[2] for a in seg) / tots,
          100 * sum(a[3] for a in seg) / samp, seg[0][1][:50]))
if len(sys.argv) > 3:
    lo, hi = map(int, sys.argv[3].split(":"))
    for a in acc[lo:hi]:
        print("%5d %8.2f %6d  %s" % (a[0], a[2], a[3], a[1]))
