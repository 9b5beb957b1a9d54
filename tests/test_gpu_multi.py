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


def test_single_process_multi_gpu_c_entry():
    """zfp_b200_multi_*: one process, one communicator per device from ncclCommInitAll, slabs encoded
    concurrently on their devices, the slab bit lengths exchanged by ncclAllGather on the compute streams and
    the bases derived on the device.  Lengths / bases against the single-GPU stream of the whole field, slab
    streams against the oracle-checked single-GPU slabs, decode back."""
    import ctypes as C
    import numpy as np
    import torch
    ndev = min(torch.cuda.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs two GPUs")
    import zfp_b200 as zb
    from zfp_b200 import api
    from zfp_b200.distributed import plan_slabs
    from helpers import analytic_field
    L = zb.load_library(build_if_missing=False)
    devs = (C.c_int * ndev)(*range(ndev))
    m = L.zfp_b200_multi_create(ndev, devs)
    assert m, zb.last_error()
    shape = (8 * ndev * 3 + 4, 44, 52)
    a = analytic_field(shape, np.float64)
    plans = plan_slabs(shape, ndev)
    for mode in ({"accuracy": 1e-5}, {"rate": 8}, {"reversible": True}):
        mn, mx, mp, me = api.mode_params(mode, "float64", 3)
        descs = (api.Desc * ndev)()
        slabs, words, outs, idx = [], [], [], []
        for i, p in enumerate(plans):
            d = descs[i]
            d.type, d.dims = 4, 3
            for k, n in enumerate(reversed(p.slab_shape)):
                d.n[k], d.s[k] = n, 0
            d.minbits, d.maxbits, d.maxprec, d.minexp = mn, mx, mp, me
            with torch.cuda.device(i):
                slabs.append(torch.from_numpy(a[p.z0:p.z1].copy()).to("cuda:%d" % i))
                words.append(torch.zeros(L.zfp_b200_capacity(C.byref(d), 0) // 8 + 2, dtype=torch.int64, device="cuda:%d" % i))
                outs.append(torch.empty_like(slabs[-1]))
                idx.append(L.zfp_b200_index_create())
        for i in range(ndev):
            torch.cuda.synchronize(i)
        ptr = lambda ts: (C.c_void_p * ndev)(*[t.data_ptr() for t in ts])
        bits, base = (C.c_uint64 * ndev)(), (C.c_uint64 * ndev)()
        rc = L.zfp_b200_multi_compress(m, descs, ptr(slabs), ptr(words), (C.c_void_p * ndev)(*idx), bits, base)
        assert rc == 0, zb.last_error()
        whole = zb.compress(torch.from_numpy(a).to("cuda:0"), **mode)
        assert 0 <= whole.nbytes * 8 - sum(bits) < 64, (mode, list(bits), whole.nbytes)
        assert list(base) == [sum(list(bits)[:i]) for i in range(ndev)], mode
        stream = np.zeros(whole.nbytes // 8 + 1, dtype=np.uint64)
        from zfp_b200.distributed import place_bits
        for i in range(ndev):
            place_bits(stream, base[i], words[i].cpu().numpy().view(np.uint64), bits[i])
        assert stream[: whole.nbytes // 8].tobytes() == whole.to_numpy().tobytes(), mode
        rc = L.zfp_b200_multi_decompress(m, descs, ptr(outs), ptr(words), (C.c_void_p * ndev)(*idx))
        assert rc == 0, zb.last_error()
        full = zb.decompress(whole).cpu().numpy()
        for i, p in enumerate(plans):
            assert outs[i].cpu().numpy().tobytes() == full[p.z0:p.z1].tobytes(), (mode, i)
        for ix in idx:
            L.zfp_b200_index_destroy(ix)
    L.zfp_b200_multi_destroy(m)
