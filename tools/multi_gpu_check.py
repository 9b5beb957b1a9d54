"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank compresses its slab,
fixed-rate slabs land at deterministic offsets, variable-rate slabs exchange their bit lengths with
an NCCL all_gather; the assembled stream must equal the stream one GPU produces for the whole array."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import zfp_b200 as zb
from zfp_b200 import distributed as zd
from test_gpu_fullsize import device_field

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for shape, dtype in (((250, 300, 260), torch.float64), ((2000, 3000), torch.float32), ((40, 36, 32, 44), torch.float64)):
    whole = device_field(shape, dtype)   # every rank generates the same field; uses only its slab
    plan = zd.plan_slabs(shape, world)[rank]
    slab = whole[plan.z0:plan.z1].contiguous()
    for mode in ({"rate": 8}, {"rate": 5.3}, {"accuracy": 1e-5}, {"precision": 20}, {"reversible": True}):
        c, nbits, base, lengths = zd.compress_slab_cuda(slab, plan, mode)
        total = torch.tensor([base + nbits], dtype=torch.int64, device="cuda")
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
        stream = zd.gather_stream_cuda(c, nbits, base, int(total.item()))
        ref = zb.compress(whole, **mode)
        nwords = (int(total.item()) + 63) // 64
        same = ref.nbytes == nwords * 8 and torch.equal(stream[:nwords], ref.words[:nwords])
        ok &= bool(same)
        if rank == 0:
            print("shape %s %s %s: slabs %d, total %d bits, identical to single-GPU stream: %s" % (shape, str(dtype).split(".")[-1], mode, world, int(total.item()), same), flush=True)
dist.barrier()
if rank == 0:
    print("MULTI-GPU PARITY", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
