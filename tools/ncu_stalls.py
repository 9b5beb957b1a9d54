#!/usr/bin/env python3
"""Stall-reason samples by code region from an ncu report. usage: ncu_stalls.py rep [bucket]"""
import csv, subprocess, sys
rep=sys.argv[1]; bucket=int(sys.argv[2]) if len(sys.argv)>2 else 500
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
cols=['stall_no_inst','stall_wait','stall_math','stall_short_sb','stall_long_sb','stall_branch_resolving','stall_not_selected','stall_selected','stall_mio','stall_dispatch']
body=rows[2:]
first=float(body[0][ix['Instructions Executed']])
print("%-12s %7s %7s | " % ("range","instr","samples") + " ".join("%9s" % c[6:15] for c in cols))
for b in range(0,len(body),bucket):
    seg=body[b:b+bucket]
    ins=sum(float(r[ix['Instructions Executed']]) for r in seg)/first
    smp=sum(float(r[ix['# Samples']]) for r in seg)
    if ins==0: continue
    print("%5d-%5d %7.0f %7.0f | " % (b,b+len(seg)-1,ins,smp) + " ".join("%9.0f" % sum(float(r[ix[c]]) for r in seg) for c in cols))
