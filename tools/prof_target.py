"""Small profiling target (for ncu): a few compress / decompress calls of an analytic field.
usage: prof_target.py <side | AxBxC...> [f64|f32|i32|i64] [<rate> | a<tol> | p<prec> | rev] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zfp_b200 as zb
arg0 = sys.argv[1] if len(sys.argv) > 1 else "512"
shape = tuple(int(v) for v in arg0.split("x") if v) if "x" in arg0 else (int(arg0),) * 3
dtype = {"f64": torch.float64, "f32": torch.float32, "i32": torch.int32, "i64": torch.int64}[sys.argv[2] if len(sys.argv) > 2 else "f64"]
arg = sys.argv[3] if len(sys.argv) > 3 else "8"
mode = {"reversible": True} if arg == "rev" else {"accuracy": float(arg[1:])} if arg.startswith("a") else {"precision": int(arg[1:])} if arg.startswith("p") else {"rate": float(arg)}
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = [torch.linspace(0, 1, n, device="cuda", dtype=torch.float64) for n in shape]
if len(shape) == 3:
    z, y, x = g[0][:, None, None], g[1][None, :, None], g[2][None, None, :]
    f = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) * torch.cos(4 * np.pi * z) + 0.25 * torch.sin(14 * np.pi * x * y * z)
elif len(shape) == 2:
    y, x = g[0][:, None], g[1][None, :]
    f = torch.sin(2 * np.pi * (3 * x + 0.5 * y)) + 0.25 * torch.sin(14 * np.pi * x * y)
elif len(shape) == 1:
    f = torch.sin(40 * np.pi * g[0]) + 0.25 * torch.sin(300 * np.pi * g[0] ** 2)
else:
    w, z, y, x = g[0][:, None, None, None], g[1][None, :, None, None], g[2][None, None, :, None], g[3][None, None, None, :]
    f = torch.sin(2 * np.pi * (x + 0.5 * y)) * torch.cos(3 * np.pi * z) + 0.25 * torch.sin(5 * np.pi * w * x)
f = (torch.round(f * 2 ** 20) if dtype in (torch.int32, torch.int64) else f).to(dtype)
c = zb.compress(f, **mode)
out = torch.empty_like(f)
for _ in range(reps):
    c = zb.compress(f, reuse=c, **mode)
    zb.decompress(c, out=out)
torch.cuda.synchronize()
print("done", c.nbytes)
