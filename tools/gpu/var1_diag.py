"""Which part of the single-pass variable-rate path differs from the oracle (developer diagnostic)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import zfp_b200 as zb
from oracle.oracle import Port
from helpers import analytic_field, make_field
P = Port()
shape = (260, 264, 272)
for kind in ("smooth", "noise"):
    a = analytic_field(shape, np.float64) if kind == "smooth" else make_field(shape, np.float64, seed=8, kind="noise")
    x = torch.from_numpy(a).cuda()
    mode = {"reversible": True}
    want = P.compress(a, **mode)
    for rep in range(3):
        c = zb.compress(x, header=False, **mode)
        got = c.to_numpy()
        same = got.tobytes() == want.tobytes()
        first = -1
        if not same:
            n = min(len(got), len(want))
            d = np.nonzero(got[:n] != want[:n])[0]
            first = int(d[0]) if len(d) else n
        back = zb.decompress(c).cpu().numpy()
        okd = back.tobytes() == a.tobytes()
        buf = torch.zeros(zb.max_stream_words(x.shape, x.dtype, mode, 77), dtype=torch.int64, device="cuda")
        c2 = zb.compress(x, out=buf, start_bit=77, **mode)
        w2 = P.compress_raw(a.reshape(-1), 0, a.dtype, list(reversed(a.shape)) + [0], None, mode, start_bit=77)[0]
        g2 = c2.to_numpy()
        same2 = g2.tobytes() == w2[:len(g2)].tobytes()
        f2 = -1
        if not same2:
            d = np.nonzero(g2 != w2[:len(g2)])[0]
            f2 = (int(d[0]), int(d[-1]), len(d), len(g2), len(w2))
        print(kind, rep, "stream", same, "len", len(got), len(want), "first diff word", first, "| decode lossless", okd, "| offset-77 stream", same2, f2, flush=True)
