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
