#!/usr/bin/env python3
"""Extract the reference's golden checksum tables into tests/golden/ref_checksums.json.

Source: /root/reference/tests/constants/checksums/{1,2,3,4}d{Int32,Int64,Float,Double}.h - the
tables the reference's own end-to-end tests (tests/src/endtoend/zfpEndtoendBase.c:386-464, used
unchanged by its serial, OpenMP *and* CUDA tests) compare against.  Each entry is
(key1, key2, checksum); key layout is documented in tests/utils/zfpChecksums.c:74-136:
  key1 = (((test_type << 2 | subject) << 3 | mode) << 4) | param,   key2 = packed dims.
Only numbers are extracted.  Run in the build container: python tests/golden/extract_ref_checksums.py
"""
import json
import os
import re

REF = os.environ.get("ZFP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_checksums.json")

tables = {}
d = os.path.join(REF, "tests", "constants", "checksums")
for name in sorted(os.listdir(d)):
    txt = open(os.path.join(d, name)).read()
    rows = re.findall(r"\{UINT64C\((0x[0-9a-f]+)\),\s*UINT64C\((0x[0-9a-f]+)\),\s*UINT64C\((0x[0-9a-f]+)\)\}", txt)
    tables[name[:-2]] = [[k1, k2, c] for k1, k2, c in rows]
json.dump(tables, open(OUT, "w"), indent=0, sort_keys=True)
print("wrote", OUT, {k: len(v) for k, v in tables.items()})
