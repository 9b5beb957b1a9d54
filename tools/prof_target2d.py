"""Profiling target: 2-D fp32 field, fixed rate (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import zfp_b200 as zb
from test_gpu_fullsize import device_field
side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rate = float(sys.argv[2]) if len(sys.argv) > 2 else 8
x = device_field((side, side), torch.float32)
c = zb.compress(x, rate=rate)
out = torch.empty_like(x)
c = zb.compress(x, out=c.words, rate=rate)
zb.decompress(c, out=out)
torch.cuda.synchronize()
print("done", c.nbytes)
