import os, sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import zfp_b200 as zb
from test_gpu_fullsize import device_field
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for shape, dtype in (((1024,1024,1024), torch.int32), ((512,512,512), torch.float64), ((768,768,768), torch.float32), ((512,512,512), torch.int64)):
    x = device_field(shape, dtype)
    raw = x.numel() * x.element_size()
    mode = {"reversible": True}
    c = zb.compress(x, **mode); y = torch.empty_like(x)
    tc = timeit(lambda: zb.compress(x, reuse=c, **mode)); td = timeit(lambda: zb.decompress(c, out=y))
    ok = torch.equal(x, y) if dtype in (torch.int32, torch.int64) else (x.view(torch.int64 if dtype==torch.float64 else torch.int32) == y.view(torch.int64 if dtype==torch.float64 else torch.int32)).all().item()
    print(shape, dtype, "ratio %.2f compress %.3f ms %.0f GB/s decompress %.3f ms %.0f GB/s lossless %s" % (raw / c.nbytes, tc, raw/tc/1e6, td, raw/td/1e6, ok), flush=True)
    del x, y, c
    torch.cuda.empty_cache()
