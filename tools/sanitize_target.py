"""Small set of calls covering every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zfp_b200 as zb
from helpers import analytic_field, make_field
cases = [(np.float64, (20, 24, 28)), (np.float32, (24, 20, 28)), (np.int32, (33, 18)), (np.float64, (70,)), (np.int64, (12, 16, 20)),
         (np.float64, (6, 5, 7, 6))]
n = 0
for dt, shape in cases:
    for kind in ("analytic", "noise"):
        a = analytic_field(shape, dt) if kind == "analytic" else make_field(shape, dt, 5, kind)
        x = torch.from_numpy(a).cuda()
        for mode in ({"rate": 8}, {"rate": 5.3}, {"rate": 48}, {"precision": 44}, {"accuracy": 1e-9}, {"reversible": True}):
            if np.dtype(dt).kind != "f" and "accuracy" in mode:
                continue
            c = zb.compress(x, **mode)
            y = zb.decompress(c)
            nb = int(np.prod([(s + 3) // 4 for s in shape]))
            zb.decompress_blocks(c, nb // 3, nb // 2 + 1, y)
            n += 1
torch.cuda.synchronize()
print("calls", n, "launches", zb.launch_count())
