"""Timing on rough data (uniform noise): the lockstep coders' per-item fallback paths (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zfp_b200 as zb
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
torch.manual_seed(1)
for dtype in (torch.float64, torch.float32):
    x = (torch.rand((512, 512, 512), device="cuda", dtype=torch.float64) - 0.5).to(dtype)
    raw = x.numel() * x.element_size()
    for mode in ({"rate": 8}, {"rate": 16}, {"precision": 16}, {"accuracy": 1e-3}):
        c = zb.compress(x, **mode); y = torch.empty_like(x)
        tc = timeit(lambda: zb.compress(x, reuse=c, **mode)); td = timeit(lambda: zb.decompress(c, out=y))
        print("noise 512^3 %s %s: ratio %.2f compress %.3f ms %.0f GB/s | decompress %.3f ms %.0f GB/s" % (
            str(dtype).split(".")[-1], mode, raw / c.nbytes, tc, raw / tc / 1e6, td, raw / td / 1e6), flush=True)
