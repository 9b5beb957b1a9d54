import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, torch
import zfp_b200 as zb
from test_gpu_fuzz import _cases
from helpers import make_field
from oracle.oracle import Port
P=Port(); bad=[]; n=0
import os
lo = int(os.environ.get('FUZZ_LO', '1000')); hi = int(os.environ.get('FUZZ_HI', '1040'))
for seed in range(lo, hi):
    for shape, dtype, kind, mode, fseed in _cases(seed, 70):
        a = make_field(shape, dtype, fseed, kind); x = torch.from_numpy(a).cuda()
        try:
            c = zb.compress(x, **mode); want = P.compress(a, **mode); got = c.to_numpy()
            ok = got.tobytes() == want.tobytes()
            if ok:
                ok = zb.decompress(c).cpu().numpy().tobytes() == P.decompress(want, a.shape, a.dtype, **mode).tobytes()
        except Exception as e:
            ok = False; print("EXC", e)
        n += 1
        if not ok: bad.append((shape, np.dtype(dtype).name, kind, mode, fseed)); print("BAD", bad[-1], flush=True)
print("cases", n, "bad", len(bad))
