#!/usr/bin/env python3
"""Generate tests/golden/kat.json: known-answer vectors from the UNMODIFIED reference library.

Inputs come from tests/helpers.make_field (pure-integer recipes, reproducible anywhere).  For
each case the reference serial path (oracle/_ref/libzfp_ref.so, built by oracle/Makefile from
/root/reference) compresses and decompresses; we record the byte count, a sha256 of the stream
and a sha256 of the decompressed array.  Run in the build container:
    make -C oracle && python tests/golden/make_kat.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import make_field, sha  # noqa: E402
from oracle.oracle import Reference  # noqa: E402

SHAPES = {1: [(4,), (13,), (259,)], 2: [(4, 4), (10, 7), (33, 18)], 3: [(4, 4, 4), (9, 11, 6), (20, 20, 20)],
          4: [(4, 4, 4, 4), (5, 6, 7, 8), (9, 8, 8, 9)]}
DTYPES = ["float32", "float64", "int32", "int64"]


def cases():
    seed = 0
    for dims, shapes in SHAPES.items():
        for shape in shapes:
            for dt in DTYPES:
                isf = dt.startswith("float")
                kinds = ["smooth", "noise", "sparse"] + (["tiny", "huge"] if isf else [])
                for kind in kinds:
                    modes = [{"rate": 3.5}, {"rate": 8}, {"rate": 20}, {"precision": 7}, {"precision": 40},
                             {"reversible": True}, {"expert": [100, 300, 20, -1074]}]
                    if isf:
                        modes += [{"accuracy": 2.0 ** -6}, {"accuracy": 1e-3}, {"accuracy": 0}]
                    for mode in modes:
                        seed += 1
                        yield dict(shape=list(shape), dtype=dt, kind=kind, seed=seed, mode=mode)
                if isf:
                    seed += 1
                    yield dict(shape=list(shape), dtype=dt, kind="special", seed=seed, mode={"reversible": True})


def main():
    R = Reference()
    out = []
    for c in cases():
        a = make_field(tuple(c["shape"]), c["dtype"], c["seed"], c["kind"])
        mode = dict(c["mode"])
        if "expert" in mode:
            mode["expert"] = tuple(mode["expert"])
        words = R.compress(a, **mode)
        back = R.decompress(words, a.shape, a.dtype, **mode)
        c.update(input=sha(a), nbytes=int(words.nbytes), stream=sha(words), decoded=sha(back))
        out.append(c)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
        f.write("\n")
    print(len(out), "cases")


if __name__ == "__main__":
    main()
