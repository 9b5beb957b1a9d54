"""Time every BASELINE.json configuration that fits one GPU (developer/reporting tool; the driver's
contract line comes from bench.py).  Prints one JSON object per config."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zfp_b200 as zb
from test_gpu_fullsize import device_field

PEAK = 6540.8
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

configs = [
    ("C2 3D f64 1024^3", (1024, 1024, 1024), torch.float64, [{"rate": 4}, {"rate": 8}, {"rate": 16}]),
    ("C2 3D f32 1024^3", (1024, 1024, 1024), torch.float32, [{"rate": 4}, {"rate": 8}, {"rate": 16}]),
    ("C3 2D f32 16384^2", (16384, 16384), torch.float32, [{"rate": 4}, {"rate": 8}, {"rate": 16}]),
    ("C3 4D f64 64^4", (64, 64, 64, 64), torch.float64, [{"rate": 8}]),
    ("C3 3D i32 1024^3", (1024, 1024, 1024), torch.int32, [{"reversible": True}]),
    ("C4 3D f64 1024^3", (1024, 1024, 1024), torch.float64, [{"accuracy": 1e-6}, {"precision": 32}]),
    ("1D f64 2^28", (1 << 28,), torch.float64, [{"rate": 8}]),
]
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, shape, dtype, modes in configs:
    if only and only not in name: continue
    x = device_field(shape, dtype)
    raw = x.numel() * x.element_size()
    for mode in modes:
        c = zb.compress(x, **mode)
        words = c.words
        y = torch.empty_like(x)
        tc = timeit(lambda: zb.compress(x, reuse=c, **mode))
        td = timeit(lambda: zb.decompress(c, out=y))
        print(json.dumps({"config": name, "mode": mode, "raw_GiB": round(raw / 2**30, 3), "ratio": round(raw / c.nbytes, 2),
                          "compress_ms": round(tc, 3), "compress_GBs": round(raw / tc / 1e6, 1),
                          "compress_hbm_frac": round((raw + c.nbytes) / tc / 1e6 / PEAK, 3),
                          "decompress_ms": round(td, 3), "decompress_GBs": round(raw / td / 1e6, 1),
                          "decompress_hbm_frac": round((raw + c.nbytes) / td / 1e6 / PEAK, 3)}), flush=True)
        del c, y, words
    del x
    torch.cuda.empty_cache()
