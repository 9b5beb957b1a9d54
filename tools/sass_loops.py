#!/usr/bin/env python3
"""List the loops (backward branches) of a kernel in a cubin with their static size and pipe mix.
   tools/sass_loops.py <cubin> <mangled-kernel-substring>"""
import re, subprocess, sys, collections
sys.path.insert(0, __import__("os").path.dirname(__file__))
cubin, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
start = end = None
for i, l in enumerate(txt):
    if l.startswith("//---") and ".text." in l:
        if start is not None:
            end = i; break
        if pat in l:
            start = i
end = end or len(txt)
ALU = set("LOP3 IADD3 SHF PRMT ISETP SEL LEA FSEL PLOP3 BMSK VIMNMX VIMNMX3 VIADDMNMX P2R R2P MOV SGXT IABS FMNMX CS2R".split())
FMA = set("IMAD FFMA FMUL FADD VIADD HFMA2".split())
ri = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)(.*?);")
labels, insts = {}, []
for l in txt[start:end]:
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        labels[m.group(1)] = len(insts)
    m = ri.match(l)
    if m:
        insts.append((int(m.group(1), 16), m.group(2), m.group(3)))
for idx, (addr, op, rest) in enumerate(insts):
    if op == "BRA":
        m = re.search(r"`\((\.L_x_\d+)\)", rest)
        if m and m.group(1) in labels and labels[m.group(1)] <= idx:
            lo = labels[m.group(1)]
            c = collections.Counter()
            for a, o, r in insts[lo:idx + 1]:
                c["alu" if o in ALU else "fma" if o in FMA else "lsu" if o in ("LDS", "STS", "LDG", "STG", "LDC", "LDCU") else "xu" if o in ("POPC", "FLO", "BREV", "F2I", "I2F", "MUFU") else "other"] += 1
            print("loop @%05x..%05x  %4d instr  %s" % (insts[lo][0], addr, idx + 1 - lo, dict(c)))
