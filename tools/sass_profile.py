#!/usr/bin/env python3
"""Join an `ncu --page source --csv` export (per-SASS executed counts) with `nvdisasm -gi` line info
of the same cubin and print dynamic instruction counts per source function and per issue pipe.

  tools/sass_profile.py <cubin> <mangled-kernel-substring> [<ncu-source.csv>]

Without the csv the counts are static (one per SASS instruction)."""
import csv, re, subprocess, sys, collections, os

PIPE = {}
for op in "LOP3 IADD3 SHF PRMT ISETP SEL LEA FSEL PLOP3 BMSK VIMNMX VIMNMX3 VIADDMNMX P2R R2P MOV SGXT IABS FMNMX CS2R".split():
    PIPE[op] = "alu"
for op in "IMAD FFMA FMUL FADD VIADD HFMA2 IDP".split():
    PIPE[op] = "fma"
for op in "DMUL DADD DFMA DSETP".split():
    PIPE[op] = "fp64"
for op in "F2I I2F MUFU POPC FLO BREV F2F I2I".split():
    PIPE[op] = "xu"
for op in "LDS STS LDG STG LDL STL LDC LDCU ATOMS ATOMG RED SHFL LDSM".split():
    PIPE[op] = "lsu"
for op in "BRA BSSY BSYNC EXIT CALL RET WARPSYNC VOTE BAR NOP S2R S2UR".split():
    PIPE[op] = "ctl"

def func_table(path):
    """line -> enclosing function name (crude: last line that looks like a function header)"""
    tab, cur = [], "?"
    rx = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__host__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    rx2 = re.compile(r"^\s*(?:__device__\s+)?(?:__forceinline__\s+)?(?:static\s+)?[\w:<>,\s\*&]+?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$")
    with open(path) as f:
        for i, line in enumerate(f, 1):
            m = rx.match(line)
            if m and m.group(1) not in ("if", "for", "while", "switch", "asm"):
                cur = m.group(1)
            elif line.startswith("  __device__") or line.startswith("  void") :
                m2 = rx2.match(line)
                if m2 and m2.group(1) not in ("if", "for", "while", "switch", "asm"):
                    cur = m2.group(1)
            tab.append(cur)
    return tab

def main():
    cubin, pat = sys.argv[1], sys.argv[2]
    csvp = sys.argv[3] if len(sys.argv) > 3 else None
    depth = int(os.environ.get("DEPTH", "0"))  # 0: innermost frame, 1: its caller, ...
    txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, l in enumerate(txt):
        if l.startswith("//---") and ".text." in l:
            if start is not None:
                end = i; break
            if pat in l:
                start = i
    else:
        end = len(txt)
    tables = {}
    insts = []  # (opcode, [frames innermost..outermost as (file,line)])
    frames = []
    rl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    ri = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)")
    pending = []
    for l in txt[start:end]:
        m = rl.search(l)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = ri.match(l)
        if m:
            if pending:
                frames = pending
                pending = []
            insts.append((m.group(2), frames))
    counts = None
    if csvp:
        rows = list(csv.reader(open(csvp)))
        hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
        col = rows[hdr].index("Instructions Executed")
        counts = [int(r[col]) for r in rows[hdr + 1:] if len(r) > col]
        assert len(counts) == len(insts), (len(counts), len(insts))
    byfn = collections.defaultdict(lambda: collections.Counter())
    for idx, (op, fr) in enumerate(insts):
        n = counts[idx] if counts else 1
        f = fr[min(depth, len(fr) - 1)] if fr else ("?", 0)
        if f[0] not in tables and os.path.exists(f[0]):
            tables[f[0]] = func_table(f[0])
        fn = tables[f[0]][f[1] - 1] if f[0] in tables and f[1] - 1 < len(tables[f[0]]) else "?"
        byfn[fn][PIPE.get(op, "other:" + op)] += n
    scale = 1.0
    if counts:
        scale = 1.0 / counts[0]  # per warp that ran the kernel prologue
    pipes = ["alu", "fma", "fp64", "xu", "lsu", "ctl"]
    tot = collections.Counter()
    print("%-28s %8s | " % ("function", "total") + " ".join("%7s" % p for p in pipes) + " | other")
    for fn, c in sorted(byfn.items(), key=lambda kv: -sum(kv[1].values())):
        t = sum(c.values())
        other = {k: v for k, v in c.items() if k not in pipes}
        print("%-28s %8.1f | " % (fn, t * scale) + " ".join("%7.1f" % (c[p] * scale) for p in pipes) + " | " +
              " ".join("%s=%.1f" % (k[6:], v * scale) for k, v in other.items()))
        tot.update(c)
    t = sum(tot.values())
    print("%-28s %8.1f | " % ("TOTAL", t * scale) + " ".join("%7.1f" % (tot[p] * scale) for p in pipes))

main()
