"""Small profiling target: a few compress/decompress calls of a 3-D field (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zfp_b200 as zb
side = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dtype = {"f64": torch.float64, "f32": torch.float32, "i32": torch.int32, "i64": torch.int64}[sys.argv[2] if len(sys.argv) > 2 else "f64"]
arg = sys.argv[3] if len(sys.argv) > 3 else "8"
mode = {"reversible": True} if arg == "rev" else {"accuracy": float(arg[1:])} if arg.startswith("a") else {"precision": int(arg[1:])} if arg.startswith("p") else {"rate": float(arg)}
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.linspace(0, 1, side, device="cuda", dtype=torch.float64)
z, y, x = g[:, None, None], g[None, :, None], g[None, None, :]
f = (torch.sin(2 * np.pi * (3 * x + 0.5 * y)) * torch.cos(4 * np.pi * z) + 0.25 * torch.sin(14 * np.pi * x * y * z))
f = (torch.round(f * 2 ** 20) if dtype in (torch.int32, torch.int64) else f).to(dtype)
c = zb.compress(f, **mode)
out = torch.empty_like(f)
for _ in range(reps):
    c = zb.compress(f, out=c.words, **mode)
    zb.decompress(c, out=out)
torch.cuda.synchronize()
print("done", c.nbytes)
