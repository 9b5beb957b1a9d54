"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Every comparison is bit-exact (streams byte for byte, decompressed arrays bit for bit); the
only floating-point tolerance that appears is the user's own fixed-accuracy bound.
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, MODE_ID, analytic_field, make_field, ref_key, ref_mode_cases, ref_table, sha

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64, np.int32, np.int64]


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import zfp_b200
    zfp_b200.load_library(build_if_missing=False)
    return zfp_b200


def _mode(c):
    mode = dict(c["mode"])
    if "expert" in mode:
        mode["expert"] = tuple(mode["expert"])
    return mode


def test_kat_vectors_host_pointers(zb):
    """1728 known-answer vectors made by the unmodified reference (tests/golden/make_kat.py):
    1-4 D, four types, partial blocks, all modes; host arrays staged by the backend."""
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        kat = json.load(f)
    bad = []
    for c in kat:
        a = make_field(tuple(c["shape"]), c["dtype"], c["seed"], c["kind"])
        mode = _mode(c)
        variable = "rate" not in mode
        res = zb.compress_numpy(a, want_index=variable, **mode)
        words, nbytes = res[0], res[1]
        ok = nbytes == c["nbytes"] and sha(words) == c["stream"]
        if ok:
            back, used = zb.decompress_numpy(words, a.shape, a.dtype, index=res[2] if variable else None, **mode)
            ok = used == nbytes and sha(back) == c["decoded"]
        if not ok:
            bad.append({k: c[k] for k in ("shape", "dtype", "kind", "mode")})
    assert not bad, "%d of %d vectors differ, first: %r" % (len(bad), len(kat), bad[:5])


@pytest.mark.parametrize("dims", [1, 2, 3, 4])
@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_golden_tables_device(zb, port, reftest, dtype, dims):
    """The reference's own end-to-end checksum tables (the ones its CUDA tests use,
    tests/src/endtoend/cudaExecBase.c:107-111), device-resident input and stream."""
    import torch
    a = reftest.smooth_field(dtype, dims)
    side = a.shape[0]
    table = ref_table(dtype, dims)
    x = torch.from_numpy(a).cuda()
    for name, p, mode in ref_mode_cases(dtype):
        c = zb.compress(x, **mode)
        words = c.to_numpy()
        assert port.hash_stream(words) == table[ref_key(1, MODE_ID[name], p, side, dims)], (name, p)
        back = zb.decompress(c).cpu().numpy()
        if name == "reversible":
            assert back.tobytes() == a.tobytes()
        else:
            assert port.hash_array(back) == table[ref_key(2, MODE_ID[name], p, side, dims)], (name, p)
            if name == "accuracy":
                assert np.max(np.abs(back.astype(np.float64) - a.astype(np.float64))) <= mode["accuracy"]


@pytest.mark.parametrize("dtype,shape", [(np.float64, (96, 100, 128)), (np.float32, (128, 96, 100)),
                                         (np.float32, (1000, 1028)), (np.float64, (515, 300)),
                                         (np.float64, (20, 24, 28, 32)), (np.int32, (64, 64, 68)),
                                         (np.float64, (100003,))])
def test_device_vs_oracle_analytic(zb, port, dtype, shape):
    """Smooth analytic fields, device resident, against the oracle on the same bytes."""
    import torch
    a = analytic_field(shape, dtype)
    x = torch.from_numpy(a).cuda()
    modes = [{"rate": 4}, {"rate": 8}, {"rate": 16}, {"precision": 32}, {"reversible": True}]
    if np.dtype(dtype).kind == "f":
        modes += [{"accuracy": 1e-6}]
    for mode in modes:
        c = zb.compress(x, **mode)
        want = port.compress(a, **mode)
        got = c.to_numpy()
        assert got.nbytes == want.nbytes, (mode, got.nbytes, want.nbytes)
        assert got.tobytes() == want.tobytes(), mode
        back = zb.decompress(c).cpu().numpy()
        assert back.tobytes() == port.decompress(want, a.shape, a.dtype, **mode).tobytes(), mode
        if "accuracy" in mode:
            assert np.max(np.abs(back.astype(np.float64) - a.astype(np.float64))) <= mode["accuracy"]


@pytest.mark.parametrize("dtype,shape", [(np.float64, (36, 40, 44)), (np.float32, (40, 36, 44)), (np.int64, (33, 32, 36)),
                                         (np.float64, (70, 66)), (np.int32, (1000,))])
def test_long_blocks_and_noisy_planes(zb, port, dtype, shape):
    """Noise and mixed noise/smooth/zero fields at high precision: blocks far longer than the
    shared-memory staging window of the variable-rate kernels (drain / restage), bit planes with
    many runs (the lockstep coders' per-item fallback), warps whose lanes finish at very
    different planes, and fixed rates where the budget runs out inside a group test."""
    import torch
    noise = make_field(shape, dtype, 11, "noise")
    smooth = make_field(shape, dtype, 12, "smooth")
    mixed = noise.copy()
    flat = mixed.reshape(-1)
    flat[: flat.size // 3] = smooth.reshape(-1)[: flat.size // 3]   # smooth part
    flat[flat.size // 3: flat.size // 2] = 0                          # zero blocks
    modes = [{"precision": 48}, {"precision": 30}, {"rate": 3}, {"rate": 24}, {"rate": 48}, {"expert": (64, 1900, 40, -1074)}]
    if np.dtype(dtype).kind == "f":
        modes += [{"accuracy": 1e-12}, {"accuracy": 1e-3}, {"expert": (1, 16000, 64, -1074)}]
    for a in (noise, mixed):
        x = torch.from_numpy(a).cuda()
        for mode in modes:
            c = zb.compress(x, **mode)
            want = port.compress(a, **mode)
            got = c.to_numpy()
            assert got.nbytes == want.nbytes, (mode, got.nbytes, want.nbytes)
            assert got.tobytes() == want.tobytes(), mode
            back = zb.decompress(c).cpu().numpy()
            assert back.tobytes() == port.decompress(want, a.shape, a.dtype, **mode).tobytes(), mode


def _arbitrary_stream(nblocks, maxbits, ebits, ebias, seed):
    """Fixed-rate stream (maxbits a multiple of 32) whose blocks are random bit soup of varying density
    behind a sane header: '1' + an exponent near the bias.  Any bit pattern is a legal block."""
    from helpers import splitmix64
    wpb = maxbits // 32
    n = nblocks * wpb
    r = [(splitmix64(np.arange(n, dtype=np.uint64), seed + 7919 * j) >> np.uint64(17)).astype(np.uint32) for j in range(4)]
    dens = (splitmix64(np.arange(nblocks, dtype=np.uint64), seed + 5) % np.uint64(5)).repeat(wpb)
    w = r[0].copy()
    w = np.where(dens >= 1, w & r[1], w)
    w = np.where(dens >= 2, w & r[2], w)
    w = np.where(dens >= 3, w & r[3], w)
    w = np.where(dens >= 4, w & (r[1] >> np.uint32(7)) & (r[2] << np.uint32(9)), w)
    head = w.reshape(nblocks, wpb)
    if ebits:
        e = (splitmix64(np.arange(nblocks, dtype=np.uint64), seed + 3) % np.uint64(40)).astype(np.int64) - 20 + ebias
        hdr_mask = np.uint32((1 << (1 + ebits)) - 1)
        head[:, 0] = (head[:, 0] & ~hdr_mask) | np.uint32(1) | (e.astype(np.uint32) << np.uint32(1))
    flat = head.reshape(-1)
    if flat.size % 2:
        flat = np.concatenate([flat, np.zeros(1, np.uint32)])
    return np.ascontiguousarray(flat).view(np.uint64)


@pytest.mark.parametrize("dtype,shape,rates", [(np.float64, (36, 40, 44), (1, 2, 4, 8, 16, 32)), (np.float64, (70, 66), (4, 8, 16, 32)),
                                               (np.float64, (1000,), (8, 16, 24, 32, 64)), (np.float32, (1001,), (8, 16, 24)),
                                               (np.float32, (36, 40, 44), (1, 2, 4, 8, 16)),
                                               (np.int64, (33, 32, 36), (1, 2, 4, 8, 16)), (np.int32, (70, 66), (4, 8, 16))])
def test_decode_of_arbitrary_bit_streams(zb, port, dtype, shape, rates):
    """The decoder against the oracle on streams no encoder produced: random bits of several densities
    decode to coefficients anywhere in the 64-bit range, so the inverse transform wraps around in
    int64 in some blocks and not in others - the FP64-pipe tail of the fp64 kernels must tell the two
    apart (its range bound) and agree with the integer arithmetic of the reference bit for bit."""
    dt = np.dtype(dtype)
    ebits, ebias = {"float64": (11, 1023), "float32": (8, 127)}.get(dt.name, (0, 0))
    nblocks = int(np.prod([(n + 3) // 4 for n in shape]))
    for rate in rates:
        maxbits = rate * 4 ** len(shape)
        assert maxbits % 32 == 0
        for seed in (1, 2):
            words = _arbitrary_stream(nblocks, maxbits, ebits, ebias, 1000 * rate + seed)
            got, _ = zb.decompress_numpy(words, shape, dt, rate=rate)
            want = port.decompress(words, shape, dt, rate=rate)
            assert got.tobytes() == want.tobytes(), (dt.name, shape, rate, seed)


@pytest.mark.parametrize("dtype,shape,rate", [(np.float32, (262, 256, 256), 8), (np.float64, (70, 512, 500), 4),
                                              (np.float32, (5000, 4100), 16), (np.float64, (9000001,), 8)])
def test_host_buffers_slab_pipeline(zb, dtype, shape, rate):
    """Host field + host stream at a fixed rate above the pipelining threshold: the backend cuts the
    array into slabs along the slowest dimension and overlaps H2D / kernels / D2H.  The stream and
    the decoded array must be identical to what the device-resident (single launch) path gives."""
    import torch
    a = analytic_field(shape, dtype)
    x = torch.from_numpy(a).cuda()
    c = zb.compress(x, rate=rate)
    want = c.to_numpy()
    words, nbytes = zb.compress_numpy(a, rate=rate)[:2]
    assert nbytes == want.nbytes
    assert words[: nbytes // 8].tobytes() == want.tobytes()
    back, used = zb.decompress_numpy(words, a.shape, a.dtype, rate=rate)
    assert used == nbytes
    assert back.tobytes() == zb.decompress(c).cpu().numpy().tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_strides_and_stream_offsets(zb, port, dtype):
    """Negative, gapped and permuted strides; payload starting mid-word after a header."""
    import torch
    rng = np.random.default_rng(3)
    for dims, n in [(1, [37, 0, 0, 0]), (2, [13, 10, 0, 0]), (3, [9, 6, 7, 0]), (4, [5, 6, 4, 7])]:
        total = int(np.prod([v for v in n if v]))
        base = make_field((3 * total + 11,), dtype, seed=dims, kind="smooth")
        shape = tuple(reversed([v for v in n if v]))
        layouts = []
        s, acc = [0] * 4, 1
        for d in range(dims):
            s[d] = -acc
            acc *= n[d]
        layouts.append((s, total - 1 + 5))
        s, acc = [0] * 4, 2
        for d in range(dims):
            s[d] = acc
            acc *= n[d]
        layouts.append((s, 3))
        if dims > 1:
            s, acc = [0] * 4, 1
            for d in reversed(range(dims)):
                s[d] = acc
                acc *= n[d]
            layouts.append((s, 0))
        xb = torch.from_numpy(base).cuda()
        for s, off in layouts:
            tstrides = tuple(reversed(s[:dims]))
            x = torch.as_strided(xb, shape, tstrides, off) if min(tstrides) > 0 else None
            for mode in ({"rate": 6}, {"precision": 11}, {"reversible": True}):
                for start in (0, 96, 37):
                    prefix = rng.integers(0, 2 ** 63, size=2, dtype=np.uint64)
                    prefix[start // 64:] = 0
                    if start % 64:
                        prefix[start // 64] = rng.integers(0, 2 ** 63, dtype=np.uint64) & np.uint64((1 << (start % 64)) - 1)
                    want, end = port.compress_raw(base, off, dtype, n, s, mode, start_bit=start, prefix_words=prefix)
                    # host view with the same strides (numpy handles negative strides)
                    view = np.lib.stride_tricks.as_strided(base[off:], shape, tuple(v * base.itemsize for v in tstrides))
                    got, nbytes = zb.compress_numpy(view, start_bit=start, prefix_words=prefix, **mode)
                    assert nbytes == 8 * ((end + 63) // 64)
                    assert got.tobytes() == want.tobytes(), (dims, s, mode, start)
                    if x is not None:
                        out = torch.zeros(zb.max_stream_words(shape, x.dtype, mode, start), dtype=torch.int64, device="cuda")
                        out[:2] = torch.from_numpy(prefix.view(np.int64)).cuda()
                        c = zb.compress(x, out=out, start_bit=start, **mode)
                        assert c.to_numpy().tobytes() == want.tobytes(), (dims, s, mode, start, "device")
                        y = torch.zeros_like(xb)
                        yv = torch.as_strided(y, shape, tstrides, off)
                        zb.decompress(c, out=yv)
                        ref_out = np.zeros_like(base)
                        port.decompress_raw(want, ref_out, off, dtype, n, s, mode, start_bit=start)
                        assert y.cpu().numpy().tobytes() == ref_out.tobytes()


def test_header_roundtrip_device_stream(zb, port):
    """zfp_write_header on a device-resident stream followed by compress: the payload starts at
    bit 96 and the header survives (the reference CUDA path overwrites it, execution.rst:203-204)."""
    import torch
    a = analytic_field((33, 40, 36), np.float64)
    x = torch.from_numpy(a).cuda()
    for mode in ({"rate": 8}, {"accuracy": 1e-4}):
        c = zb.compress(x, header=True, **mode)
        words = c.to_numpy()
        payload, end = port.compress_raw(a.reshape(-1), 0, a.dtype, [36, 40, 33, 0], None, mode, start_bit=96,
                                         prefix_words=words[:2] & np.array([2 ** 64 - 1, 2 ** 32 - 1], dtype=np.uint64))
        assert words.tobytes() == payload.tobytes()
        assert bytes(words[:1].tobytes()[:4]) == b"zfp\x05"
        back = zb.decompress(c, header=True)
        want = np.zeros_like(a)
        port.decompress_raw(payload, want.reshape(-1), 0, a.dtype, [36, 40, 33, 0], None, mode, start_bit=96)
        assert back.cpu().numpy().tobytes() == want.tobytes()


def test_foreign_variable_rate_stream_without_index(zb, port, ref):
    """Streams produced elsewhere (here: by the reference CPU library) carry no block index; the
    backend rebuilds it on the device and decodes bit-identically."""
    for dtype, shape in ((np.float64, (20, 22, 26)), (np.float32, (30, 41)), (np.int32, (9, 8, 7, 6)), (np.float64, (301,))):
        a = make_field(shape, dtype, seed=21, kind="smooth")
        for mode in ({"accuracy": 1e-3}, {"precision": 18}, {"reversible": True}, {"expert": (40, 700, 30, -30)}):
            if np.dtype(dtype).kind != "f" and "accuracy" in mode:
                continue
            words = ref.compress(a, **mode)
            want = ref.decompress(words, a.shape, a.dtype, **mode)
            got, used = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode)
            assert used == words.nbytes, (shape, mode)
            assert got.tobytes() == want.tobytes(), (shape, mode)


def test_foreign_stream_index_rebuilt_in_parallel(zb, port, monkeypatch):
    """From 16384 blocks on, zfp_decompress rebuilds the index of a stream that came without one segment-parallel
    (speculative walks stitched at block boundaries, zfp_b200_index_rebuild) instead of with one thread; the decode
    verifies the candidate.  Same result as the oracle and as the sequential walk, in a fraction of its time."""
    import time
    cases = [(np.float64, (160, 160, 164), {"accuracy": 1e-7}), (np.float64, (160, 160, 164), {"precision": 28}),
             (np.float32, (2048, 2052), {"reversible": True}), (np.int64, (1200000,), {"precision": 40}),
             (np.float64, (160, 156, 164), {"expert": (40, 900, 40, -40)})]
    for dtype, shape, mode in cases:
        a = make_field(shape, dtype, seed=31, kind="smooth")
        a.reshape(-1)[: a.size // 5] = 0                     # a run of empty blocks: one bit each
        words = port.compress(a, **mode)
        want = port.decompress(words, a.shape, a.dtype, **mode)
        n0 = zb.launch_count()
        t0 = time.perf_counter()
        got, used = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode)
        t_par = time.perf_counter() - t0
        launches_par = zb.launch_count() - n0
        assert used == words.nbytes and got.tobytes() == want.tobytes(), (shape, mode)
        monkeypatch.setenv("ZFP_B200_SERIAL_INDEX", "1")
        n0 = zb.launch_count()
        t0 = time.perf_counter()
        got2, _ = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode)
        t_ser = time.perf_counter() - t0
        launches_ser = zb.launch_count() - n0
        monkeypatch.delenv("ZFP_B200_SERIAL_INDEX")
        assert got2.tobytes() == want.tobytes()
        assert words.nbytes * 8 >= 1 << 23, "test case too small for the parallel rebuild"
        assert launches_par > launches_ser, "the parallel rebuild did not run"
        assert t_par < 0.6 * t_ser, (shape, mode, t_par, t_ser)
    # a payload that starts in the middle of a word (something was written before it)
    dtype, shape, mode = cases[1]
    a = make_field(shape, dtype, seed=32, kind="smooth")
    n4 = list(reversed(a.shape)) + [0] * (4 - a.ndim)
    shifted = port.compress_raw(a.reshape(-1), 0, a.dtype, n4, None, mode, start_bit=77)[0]
    want = port.decompress(port.compress(a, **mode), a.shape, a.dtype, **mode)
    n0 = zb.launch_count()
    got, _ = zb.decompress_numpy(shifted, a.shape, a.dtype, start_bit=77, index=None, **mode)
    assert got.tobytes() == want.tobytes()
    assert zb.launch_count() - n0 > 6, "the parallel rebuild did not run"



def test_untrusted_block_index_cannot_mislead_the_decoder(zb, port):
    """A block index that comes from outside (zfp_b200_index_import, the zfpy trailer) with lengths that
    are wrong - every block at the 16-bit maximum, all zero, or those of another field - must neither
    send reads past the stream (the scan and the staging cap every length at the worst case of a block)
    nor change the result: the decode notices the mismatch and rebuilds the index from the stream."""
    for dtype, shape in ((np.float64, (36, 40, 44)), (np.float32, (70, 66)), (np.int32, (20, 12, 9, 5))):
        a = make_field(shape, dtype, seed=5, kind="smooth")
        other = make_field(shape, dtype, seed=6, kind="noise")
        nblocks = int(np.prod([(n + 3) // 4 for n in shape]))
        for mode in ({"precision": 20}, {"reversible": True}):
            words = port.compress(a, **mode)
            want = port.decompress(words, a.shape, a.dtype, **mode)
            _, _, foreign = zb.compress_numpy(other, want_index=True, **mode)
            for lengths in (np.full(nblocks, 65535, np.uint16), np.zeros(nblocks, np.uint16), np.asarray(foreign, np.uint16)):
                got, used = zb.decompress_numpy(words, a.shape, a.dtype, index=lengths, **mode)
                assert used == words.nbytes, (shape, mode)
                assert got.tobytes() == want.tobytes(), (shape, mode)


def test_openmp_reference_stream_is_identical(zb, ref):
    """The reference guarantees policy-independent streams (docs/source/execution.rst:56-57): the
    GPU stream equals the serial AND the OpenMP reference streams."""
    a = analytic_field((40, 52, 48), np.float64)
    for mode in ({"rate": 8}, {"accuracy": 1e-5}, {"reversible": True}):
        got, _ = zb.compress_numpy(a, **mode)
        assert got.tobytes() == ref.compress(a, policy=0, **mode).tobytes()
        assert got.tobytes() == ref.compress(a, policy=1, threads=4, **mode).tobytes()


def test_reference_library_with_our_backend_as_its_cuda_policy(port):
    """INTEGRATION.md section A, executed: the unmodified reference libzfp built with -DZFP_WITH_CUDA
    and linked against libzfp_b200.so instead of src/cuda_zfp.  Its own zfp_compress / zfp_decompress
    under zfp_exec_cuda now run our kernels (host arrays, as in the reference's CUDA tests,
    tests/src/endtoend/cudaExecBase.c) and must reproduce the serial streams."""
    from oracle.oracle import REF_CUDA_SO, Reference
    if not os.path.exists(REF_CUDA_SO):
        pytest.skip("oracle/_ref/libzfp_ref_cuda.so not built")
    R = Reference(REF_CUDA_SO)
    for dtype, shape in ((np.float64, (33, 36, 40)), (np.float32, (65, 70)), (np.int32, (1000,)), (np.int64, (12, 16, 20))):
        a = make_field(shape, dtype, seed=31, kind="smooth")
        for rate in (4, 8, 19):
            mode = {"rate": rate}
            serial = R.compress(a, policy=0, **mode)
            cuda = R.compress(a, policy=2, **mode)
            assert cuda.tobytes() == serial.tobytes(), (shape, rate)
            assert cuda.tobytes() == port.compress(a, **mode).tobytes()
            # decompress through the reference API with the CUDA policy
            out = np.empty_like(a)
            n = list(reversed(a.shape)) + [0] * (4 - a.ndim)
            L = R.L
            f, dims = R._field(out.ctypes.data, dtype, n, None)
            words = np.concatenate([cuda, np.zeros(4, dtype=np.uint64)])
            bs = L.stream_open(words.ctypes.data, words.nbytes)
            z = L.zfp_stream_open(bs)
            R._set_mode(z, mode, dtype, dims)
            assert L.zfp_stream_set_execution(z, 2)
            assert L.zfp_decompress(z, f) == cuda.nbytes
            L.zfp_field_free(f); L.zfp_stream_close(z); L.stream_close(bs)
            assert out.tobytes() == R.decompress(serial, a.shape, a.dtype, **mode).tobytes()


@pytest.mark.parametrize("dtype,shape", [(np.float64, (40, 44, 52)), (np.float32, (70, 90)), (np.int32, (300,)),
                                         (np.float64, (8, 12, 8, 9))])
def test_random_access_block_range_decode(zb, dtype, shape):
    """zfp_b200_decode_blocks: any range of blocks decodes in parallel into its place and leaves the
    rest of the array untouched - fixed rate straight from the block number, variable rate through
    the block-offset index."""
    import torch
    a = analytic_field(shape, dtype)
    x = torch.from_numpy(a).cuda()
    nblocks = int(np.prod([(n + 3) // 4 for n in shape]))
    sentinel = 77
    modes = [{"rate": 8}, {"rate": 5.3}, {"precision": 20}, {"reversible": True}]
    if np.dtype(dtype).kind == "f":
        modes.append({"accuracy": 1e-4})
    for mode in modes:
        c = zb.compress(x, **mode)
        full = zb.decompress(c)
        # which block every element belongs to (stream order: x fastest)
        idx = np.zeros(shape, dtype=np.int64)
        mul = 1
        for ax in range(len(shape) - 1, -1, -1):
            coord = np.arange(shape[ax]) // 4
            view = [1] * len(shape)
            view[ax] = shape[ax]
            idx += coord.reshape(view) * mul
            mul *= (shape[ax] + 3) // 4
        owner = torch.from_numpy(idx).cuda()
        for b0, b1 in ((0, 1), (nblocks // 3, nblocks // 3 + 37 if nblocks > 80 else nblocks // 2 + 1), (nblocks - 1, nblocks),
                       (5, 5), (0, nblocks)):
            out = torch.full_like(full, sentinel)
            zb.decompress_blocks(c, b0, b1, out)
            inside = (owner >= b0) & (owner < b1)
            assert torch.equal(out[inside], full[inside]), (mode, b0, b1)
            assert bool((out[~inside] == sentinel).all()), (mode, b0, b1)


def test_random_access_box_decode(zb):
    """decompress_box: the blocks intersecting a coordinate box, nothing else."""
    import torch
    for dtype, shape, lo, hi in ((np.float64, (40, 44, 52), (5, 8, 17), (22, 9, 40)), (np.float32, (70, 90), (0, 33), (70, 34)),
                                 (np.float64, (9, 12, 8, 10), (4, 0, 3, 2), (9, 5, 8, 7))):
        a = analytic_field(shape, dtype)
        x = torch.from_numpy(a).cuda()
        for mode in ({"rate": 6}, {"rate": 8}, {"precision": 24}, {"accuracy": 1e-4} if np.dtype(dtype).kind == "f" else {"rate": 16}):
            c = zb.compress(x, **mode)
            full = zb.decompress(c)
            out = torch.full_like(full, 55)
            zb.decompress_box(c, lo, hi, out)  # one launch over the block list (zfp_b200_decode_box)
            inside = torch.ones(shape, dtype=torch.bool, device="cuda")
            for ax, (l, h) in enumerate(zip(lo, hi)):
                coord = torch.arange(shape[ax], device="cuda")
                keep = (coord >= (l // 4) * 4) & (coord < ((h + 3) // 4) * 4)
                view = [1] * len(shape)
                view[ax] = shape[ax]
                inside &= keep.reshape(view)
            assert torch.equal(out[inside], full[inside]), (shape, mode)
            assert bool((out[~inside] == 55).all()), (shape, mode)


def test_reference_cli_on_our_backend_c1(tmp_path):
    """BASELINE.json configs[0]: 3-D double 256^3, fixed rate 8, through the reference's own `zfp`
    command-line tool (utils/zfp.c, unmodified).  The tool linked against the drop-in library and run
    with `-x cuda -h` must write the same file, header included, as the stock CPU tool with
    `-x serial` (the reference's CUDA backend overwrites headers, docs/source/execution.rst:203-204;
    ours honours the stream offset), and decompress it to the same bytes."""
    import subprocess
    from oracle.oracle import HERE as ORACLE_DIR
    cpu, gpu = os.path.join(ORACLE_DIR, "_ref", "zfp_ref"), os.path.join(ORACLE_DIR, "_ref", "zfp_ref_cuda")
    if not (os.path.exists(cpu) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref/zfp_ref[_cuda] not built")
    a = analytic_field((256, 256, 256), np.float64)
    raw = tmp_path / "field.raw"
    a.tofile(raw)
    outs = {}
    for name, exe, policy in (("cpu", cpu, "serial"), ("gpu", gpu, "cuda")):
        z, o = tmp_path / (name + ".zfp"), tmp_path / (name + ".out")
        r = subprocess.run([exe, "-d", "-3", "256", "256", "256", "-r", "8", "-h", "-x", policy, "-i", str(raw), "-z", str(z), "-o", str(o)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs[name] = (z.read_bytes(), o.read_bytes())
    assert len(outs["gpu"][0]) == 16 + 256 ** 3  # 96-bit header rounded up to words + 8 bits/value
    assert outs["gpu"][0] == outs["cpu"][0], "compressed files differ"
    assert outs["gpu"][1] == outs["cpu"][1], "decompressed files differ"
    # and the GPU tool decompresses the CPU tool's file (header parsed, payload mid-word)
    o2 = tmp_path / "cross.out"
    r = subprocess.run([gpu, "-h", "-x", "cuda", "-z", str(tmp_path / "cpu.zfp"), "-o", str(o2)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert o2.read_bytes() == outs["cpu"][1]


def test_zfpy_compatible_module_interoperates_with_reference(zb, ref):
    """zfp_b200.zfpy: same signatures and byte format as the reference's zfpy (python/zfpy.pyx);
    streams with full headers travel both ways, with and without our index trailer."""
    from zfp_b200 import zfpy
    for dtype, shape in ((np.float64, (30, 40, 50)), (np.float32, (100, 90)), (np.int32, (24, 24, 24)), (np.float64, (12, 10, 9, 8))):
        a = make_field(shape, dtype, seed=41, kind="smooth")
        n = list(reversed(shape)) + [0] * (4 - len(shape))
        for kw, mode in (({"rate": 8}, {"rate": 8}), ({"tolerance": 1e-3}, {"accuracy": 1e-3}), ({"precision": 16}, {"precision": 16}), ({}, {"reversible": True})):
            if np.dtype(dtype).kind != "f" and "tolerance" in kw:
                continue
            ours = zfpy.compress_numpy(a, **kw)
            theirs, _ = ref.compress_raw(a.reshape(-1), 0, dtype, n, None, mode, header_mask=7)
            assert ours == theirs.tobytes(), (shape, kw)                 # identical bytes incl. header
            back_ref, used = ref.decompress_with_header(ours)           # the reference reads ours
            assert used == len(ours)
            back_ours = zfpy.decompress_numpy(theirs.tobytes())          # we read the reference's
            assert back_ours.shape == a.shape and back_ours.dtype == a.dtype
            assert back_ours.tobytes() == back_ref.tobytes()
            with_index = zfpy.compress_numpy(a, index_trailer=True, **kw)
            assert with_index[: len(ours)] == ours
            assert zfpy.decompress_numpy(with_index).tobytes() == back_ref.tobytes()
            back_ref2, _ = ref.decompress_with_header(with_index)       # trailer is invisible to zfp readers
            assert back_ref2.tobytes() == back_ref.tobytes()
            if not kw:
                assert back_ours.tobytes() == a.tobytes()
    import torch
    x = torch.from_numpy(make_field((20, 24, 28), np.float64, seed=5, kind="smooth")).cuda()
    buf, c = zfpy.compress_tensor(x, tolerance=1e-4)
    assert bytes(buf.cpu().numpy()[:4]) == b"zfp\x05"
    y = zfpy.decompress_tensor(c)
    assert float((x - y).abs().max()) <= 1e-4


def test_one_zfp_stream_many_fields_then_decompress_an_earlier_one(zb, port):
    """The block index a variable-rate compress leaves on its zfp_stream describes the LAST field
    compressed.  A legal pattern - one zfp_stream compresses u, v, w of the same shape into three
    buffers, then decompresses u - must not decode u with w's block lengths: the backend checks every
    block's parsed length against the index and falls back to rebuilding it from the stream."""
    import torch
    from zfp_b200.api import Stream, _make_field, load_library
    L = load_library()
    shape = (24, 28, 36)
    fields = [make_field(shape, np.float64, seed=s, kind=k) for s, k in ((3, "smooth"), (4, "noise"), (5, "sparse"))]
    for mode in ({"accuracy": 1e-4}, {"precision": 21}, {"reversible": True}):
        cap = zb.max_stream_words(shape, torch.float64, mode)
        bufs = [torch.zeros(cap, dtype=torch.int64, device="cuda") for _ in fields]
        s = Stream(bufs[0].data_ptr(), cap * 8, mode, 4, 3)
        sizes = []
        for a, buf in zip(fields, bufs):
            bs = L.stream_open(buf.data_ptr(), cap * 8)
            L.zfp_stream_set_bit_stream(s.z, bs)
            x = torch.from_numpy(a).cuda()
            f = _make_field(L, x.data_ptr(), 4, shape, None)
            sizes.append(L.zfp_compress(s.z, f))
            assert sizes[-1], zb.last_error()
            L.zfp_field_free(f)
            L.stream_close(bs)
        for a, buf, nbytes in zip(fields, bufs, sizes):  # the first two decodes meet a stale index
            bs = L.stream_open(buf.data_ptr(), cap * 8)
            L.zfp_stream_set_bit_stream(s.z, bs)
            out = torch.empty(shape, dtype=torch.float64, device="cuda")
            f = _make_field(L, out.data_ptr(), 4, shape, None)
            assert L.zfp_decompress(s.z, f) == nbytes, (mode, zb.last_error())
            L.zfp_field_free(f)
            L.stream_close(bs)
            want = port.decompress(port.compress(a, **mode), shape, a.dtype, **mode)
            assert out.cpu().numpy().tobytes() == want.tobytes(), mode
        L.zfp_stream_set_bit_stream(s.z, s.bs)
        s.close()


def test_two_host_threads_two_zfp_streams_one_gpu(zb, port):
    """Reference contract: thread-safe as long as threads do not share a zfp_stream
    (docs/source/faq.rst:1096-1098).  Two threads compress and decompress different fields at the same
    time on one device, host buffers and device buffers, all modes that use device scratch."""
    import threading
    import torch
    errors = []

    def worker(seed, kind):
        try:
            stream = torch.cuda.Stream()
            for rep in range(6):
                a = make_field((36, 40, 44), np.float64, seed=seed + rep, kind=kind)
                for mode in ({"accuracy": 1e-3}, {"precision": 20}, {"rate": 7.5}):
                    want = port.compress(a, **mode)
                    got = zb.compress_numpy(a, **mode)[0]                       # host staging buffers
                    assert got.tobytes() == want.tobytes(), ("host", seed, rep, mode)
                    x = torch.from_numpy(a).cuda()
                    c = zb.compress(x, cuda_stream=stream.cuda_stream, **mode)  # scan / slot scratch on a side stream
                    assert c.to_numpy().tobytes() == want.tobytes(), ("device", seed, rep, mode)
                    back = zb.decompress(c)
                    stream.synchronize()
                    assert back.cpu().numpy().tobytes() == port.decompress(want, a.shape, a.dtype, **mode).tobytes(), (seed, rep, mode)
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(100, "smooth")), threading.Thread(target=worker, args=(200, "noise"))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def _ref_cuda_roundtrip(R, a, mode):
    """zfp_compress + zfp_decompress of the reference library under zfp_exec_cuda on host arrays."""
    cuda = R.compress(a, policy=2, **mode)
    out = np.empty_like(a)
    n = list(reversed(a.shape)) + [0] * (4 - a.ndim)
    L = R.L
    f, dims = R._field(out.ctypes.data, a.dtype, n, None)
    words = np.concatenate([cuda, np.zeros(4, dtype=np.uint64)])
    bs = L.stream_open(words.ctypes.data, words.nbytes)
    z = L.zfp_stream_open(bs)
    R._set_mode(z, mode, a.dtype, dims)
    assert L.zfp_stream_set_execution(z, 2)
    used = L.zfp_decompress(z, f)
    L.zfp_field_free(f); L.zfp_stream_close(z); L.stream_close(bs)
    return cuda, out, used


@pytest.mark.parametrize("dtype", DTYPES)
def test_patched_reference_dispatch_all_modes_all_dims(dtype):
    """North-star subsystem (1), executed: the reference library with integration/zfp_cuda_dispatch.patch
    applied (fixed-rate gate of src/template/cuda{,de}compress.c lifted, 4-D rows of the function tables in
    src/zfp.c:1085,1089,1145,1149 filled) on our backend.  Through the reference's OWN zfp_compress /
    zfp_decompress with zfp_exec_cuda, all four modes + expert, 1-4 D, every scalar type give the streams
    and arrays of its serial policy."""
    from oracle.oracle import REF_CUDA_ALL_SO, Reference
    if not os.path.exists(REF_CUDA_ALL_SO):
        pytest.skip("oracle/_ref/libzfp_ref_cuda_all.so not built (make -C oracle ref_cuda_all)")
    R = Reference(REF_CUDA_ALL_SO)
    fp = np.dtype(dtype).kind == "f"
    modes = [{"rate": 8}, {"rate": 11.5}, {"precision": 17}, {"reversible": True}, {"expert": (30, 900, 40, -40)}]
    if fp:
        modes.append({"accuracy": 1e-3})
    bad = []
    for shape in ((131,), (37, 42), (18, 21, 25), (6, 9, 7, 10)):
        for kind in ("smooth", "noise"):
            a = make_field(shape, dtype, seed=41 + len(shape), kind=kind)
            for mode in modes:
                serial = R.compress(a, policy=0, **mode)
                cuda, out, used = _ref_cuda_roundtrip(R, a, mode)
                ok = cuda.tobytes() == serial.tobytes() and used == serial.nbytes and \
                    out.tobytes() == R.decompress(serial, a.shape, a.dtype, **mode).tobytes()
                if not ok:
                    bad.append((shape, kind, mode, cuda.nbytes, serial.nbytes, used))
    assert not bad, bad[:6]


def test_patched_reference_cli_variable_rate_reversible_and_4d(tmp_path):
    """The reference's command-line tool (utils/zfp.c, unmodified) on the patched library: `-x cuda` with
    -a (fixed accuracy, with a header), -R (reversible) and a 4-D array writes the files `-x serial` writes
    and decompresses them to the same bytes."""
    import subprocess
    from oracle.oracle import HERE as ORACLE_DIR
    cpu, gpu = os.path.join(ORACLE_DIR, "_ref", "zfp_ref"), os.path.join(ORACLE_DIR, "_ref", "zfp_ref_cuda_all")
    if not (os.path.exists(cpu) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref/zfp_ref[_cuda_all] not built")
    cases = [("acc", (96, 100, 104), ["-d", "-3", "104", "100", "96", "-a", "1e-6", "-h"]),
             ("rev", (64, 72, 80), ["-d", "-3", "80", "72", "64", "-R", "-h"]),
             ("prec4d", (12, 16, 20, 24), ["-d", "-4", "24", "20", "16", "12", "-p", "24"]),
             ("rate4d", (12, 16, 20, 24), ["-d", "-4", "24", "20", "16", "12", "-r", "12", "-h"])]
    for name, shape, args in cases:
        a = analytic_field(shape, np.float64)
        raw = tmp_path / (name + ".raw")
        a.tofile(raw)
        got = {}
        for who, exe, policy in (("cpu", cpu, "serial"), ("gpu", gpu, "cuda")):
            z, o = tmp_path / ("%s_%s.zfp" % (name, who)), tmp_path / ("%s_%s.out" % (name, who))
            r = subprocess.run([exe] + args + ["-x", policy, "-i", str(raw), "-z", str(z), "-o", str(o)],
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, (name, who, r.stderr)
            got[who] = (z.read_bytes(), o.read_bytes())
        assert got["gpu"][0] == got["cpu"][0], "%s: compressed files differ" % name
        assert got["gpu"][1] == got["cpu"][1], "%s: decompressed files differ" % name


def test_reference_cli_on_libzfp_b200_alone(tmp_path):
    """utils/zfp.c (unmodified) linked against libzfp_b200.so and nothing else: `-x cuda` writes the stock
    CPU tool's files in fixed-rate and fixed-accuracy mode; `-x serial` / `-x omp`, which this library
    leaves to the reference, fail cleanly instead of producing anything."""
    import subprocess
    from oracle.oracle import HERE as ORACLE_DIR
    cpu, ours = os.path.join(ORACLE_DIR, "_ref", "zfp_ref"), os.path.join(ORACLE_DIR, "_ref", "zfp_b200_cli")
    if not (os.path.exists(cpu) and os.path.exists(ours)):
        pytest.skip("oracle/_ref/zfp_ref / zfp_b200_cli not built")
    a = analytic_field((72, 80, 88), np.float64)
    raw = tmp_path / "f.raw"
    a.tofile(raw)
    dims = ["-d", "-3", "88", "80", "72"]
    for name, args in (("rate", ["-r", "10", "-h"]), ("acc", ["-a", "1e-5", "-h"])):
        files = {}
        for who, exe, policy in (("cpu", cpu, "serial"), ("ours", ours, "cuda")):
            z, o = tmp_path / ("%s_%s.zfp" % (name, who)), tmp_path / ("%s_%s.out" % (name, who))
            r = subprocess.run([exe] + dims + args + ["-x", policy, "-i", str(raw), "-z", str(z), "-o", str(o)],
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, (name, who, r.stderr)
            files[who] = (z.read_bytes(), o.read_bytes())
        assert files["ours"] == files["cpu"], name
    for policy in ("serial", "omp"):
        r = subprocess.run([ours] + dims + ["-r", "8", "-x", policy, "-i", str(raw), "-z", str(tmp_path / "no.zfp")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode != 0 and not (tmp_path / "no.zfp").exists() or (tmp_path / "no.zfp").stat().st_size == 0, policy


def test_single_pass_variable_rate_encoder_is_bit_exact():
    """The single-pass variable-rate encoder (kernels_var1.cuh: decoupled look-back, no slot scratch; opt-in with
    ZFP_B200_VAR1=1 because it measured slower than the slot path) against the oracle, in a fresh process: smooth
    and noisy data (blocks that outgrow the shared-memory window are re-encoded into their holes), several tiles,
    a header before the payload, the probe / hand-over for data with mostly long blocks."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
import zfp_b200 as zb
from oracle.oracle import Port
from helpers import analytic_field, make_field
P = Port(); bad = []
for dtype, shape in ((np.float64, (68, 72, 80)), (np.float32, (40, 44, 52)), (np.int32, (36, 40, 44)), (np.float64, (260, 264, 272))):
    for kind in ("smooth", "noise"):
        a = make_field(shape, dtype, seed=7, kind=kind) if shape[0] < 200 else (analytic_field(shape, dtype) if kind == "smooth" else make_field(shape, dtype, seed=8, kind="noise"))
        x = torch.from_numpy(a).cuda()
        for mode in ({"precision": 19}, {"accuracy": 1e-4}, {"reversible": True}, {"expert": (64, 3000, 50, -60)}):
            if np.dtype(dtype).kind != "f" and "accuracy" in mode: continue
            if shape[0] > 200 and "expert" in mode: continue
            want = P.compress(a, **mode)
            c = zb.compress(x, header=False, **mode)
            ok = c.to_numpy().tobytes() == want.tobytes() and zb.decompress(c).cpu().numpy().tobytes() == P.decompress(want, a.shape, a.dtype, **mode).tobytes()
            buf = torch.zeros(zb.max_stream_words(x.shape, x.dtype, mode, 77), dtype=torch.int64, device="cuda")
            c2 = zb.compress(x, out=buf, start_bit=77, **mode)
            w2 = P.compress_raw(a.reshape(-1), 0, a.dtype, list(reversed(a.shape)) + [0], None, mode, start_bit=77)[0]
            ok = ok and (c2.to_numpy().tobytes() == w2[:len(c2.to_numpy())].tobytes())
            if not ok: bad.append((np.dtype(dtype).name, shape, kind, mode))
print("BAD", bad) if bad else print("VAR1 OK", zb.launch_count())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, ZFP_B200_VAR1="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "VAR1 OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
