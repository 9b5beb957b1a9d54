"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): slabs compressed on different GPUs,
placed with the NCCL all_gather of bit lengths + device bit copy, equal the single-GPU stream."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_slab_streams_equal_single_gpu_stream():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_one_process_two_devices():
    """The library keeps per-device state (scratch, shared-memory opt-in of the kernels): a process
    that compresses on cuda:0 and then on cuda:1 gets identical streams and arrays from both."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import zfp_b200 as zb
    from helpers import analytic_field
    zb.load_library(build_if_missing=False)
    a = analytic_field((96, 100, 104), np.float64)
    outs = []
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            x = torch.from_numpy(a).to("cuda:%d" % dev)
            res = []
            for mode in ({"rate": 8}, {"rate": 40}, {"accuracy": 1e-6}, {"reversible": True}):
                c = zb.compress(x, **mode)
                res.append((c.to_numpy().tobytes(), zb.decompress(c).cpu().numpy().tobytes()))
            outs.append(res)
    assert outs[0] == outs[1] == outs[2]
