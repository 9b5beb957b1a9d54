"""Quick GPU sanity + timing (developer tool): parity vs the oracle on small fields, then timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zfp_b200 as zb
from oracle.oracle import Port
from helpers import analytic_field, make_field
P = Port()
bad = 0
cases = [(np.float64, (64, 64, 64)), (np.float32, (64, 64, 64)), (np.float64, (30, 33, 35)), (np.int32, (40, 40)),
         (np.float64, (100,)), (np.float64, (8, 8, 8, 8)), (np.int64, (16, 20, 24)), (np.float32, (50, 60))]
for dt, shape in ([] if os.environ.get('QUICK_NO_PARITY') else cases):
    for kind in ("analytic", "noise", "sparse"):
        a = analytic_field(shape, dt) if kind == "analytic" else make_field(shape, dt, 5, kind)
        x = torch.from_numpy(a).cuda()
        for mode in ({"rate": 8}, {"rate": 2}, {"rate": 16}, {"rate": 32}, {"rate": 5.3}, {"precision": 20}, {"accuracy": 1e-4}, {"reversible": True}):
            if np.dtype(dt).kind != 'f' and 'accuracy' in mode: continue
            c = zb.compress(x, **mode); got = c.to_numpy(); want = P.compress(a, **mode)
            ok = got.tobytes() == want.tobytes()
            back = zb.decompress(c).cpu().numpy(); ok2 = back.tobytes() == P.decompress(want, a.shape, a.dtype, **mode).tobytes()
            if not (ok and ok2):
                bad += 1; print("MISMATCH", np.dtype(dt).name, shape, kind, mode, got.nbytes, want.nbytes, ok, ok2)
print("parity mismatches:", bad)
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for dt in (torch.float64, torch.float32):
    g = torch.linspace(0, 1, side, device="cuda", dtype=torch.float64)
    z, y, xx = g[:, None, None], g[None, :, None], g[None, None, :]
    x = (torch.sin(2 * np.pi * (3 * xx + 0.5 * y)) * torch.cos(4 * np.pi * z) + 0.25 * torch.sin(14 * np.pi * xx * y * z)).to(dt)
    for rate in (4, 8, 16):
        c = zb.compress(x, rate=rate); y2 = zb.decompress(c)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        tc = td = 0.0
        for _ in range(5):
            e0.record(); c = zb.compress(x, out=c.words, rate=rate, async_fixed_rate=True); e1.record(); zb.decompress(c, out=y2); e2.record()
            torch.cuda.synchronize(); tc += e0.elapsed_time(e1) / 5; td += e1.elapsed_time(e2) / 5
        nb = x.numel() * x.element_size()
        print("%d^3 %s rate %2d: compress %.3f ms %7.1f GB/s (%.1f%% HBM) | decompress %.3f ms %7.1f GB/s (%.1f%% HBM)" % (
            side, str(dt).split('.')[-1], rate, tc, nb / tc / 1e6, 100 * (nb + c.nbytes) / tc / 1e6 / 6540.8, td, nb / td / 1e6, 100 * (nb + c.nbytes) / td / 1e6 / 6540.8))
