"""Variable-rate multi-GPU timing (BASELINE.json configs[4], second half): every rank compresses its
1024^3 fp64 slab at fixed accuracy, the ranks all_gather ONE int64 (slab bit length) over NCCL to learn
their base bit in the global stream, and decompress their slab again.  The stream stays distributed
(each GPU holds its slab + base offset, SURVEY 8e variant i).  Run under torchrun; rank 0 prints a line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import zfp_b200 as zb
from zfp_b200 import distributed as zd
from bench import field_slab, SIDE

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
zb.load_library(build_if_missing=False)
x = field_slab(torch, rank, SIDE, SIDE, SIDE, dev)
plan = zd.plan_slabs((SIDE * world, SIDE, SIDE), world)[rank]
mode = {"accuracy": 1e-6}
y = torch.empty_like(x)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5

def step():
    c, nbits, base, lengths = zd.compress_slab_cuda(x, plan, mode)
    zb.decompress(c, out=y)
    return c, nbits, base, lengths

for _ in range(3):
    c, nbits, base, lengths = step()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    c, nbits, base, lengths = step()
e1.record(); dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
err = float((x - y).abs().max())
if rank == 0:
    raw = x.numel() * 8
    print(json.dumps({"config": "3D fp64 %dx%dx%d over %d GPUs, accuracy 1e-6, compress (+ all_gather of slab bit lengths) + decompress" % (SIDE * world, SIDE, SIDE, world),
                      "ms_per_step": float(t.item()), "GBs": 2 * raw * world / (float(t.item()) * 1e-3) / 1e9,
                      "slab_bits": lengths, "base_bit_rank0": base, "ratio": raw * world * 8 / float(sum(lengths)), "max_abs_err_rank0": err}))
dist.destroy_process_group()
