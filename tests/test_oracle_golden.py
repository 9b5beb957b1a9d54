"""Pin the oracle (C restatement) before anything trusts it.

(1) the reference's own golden checksum tables on the reference's own seeded fields;
(2) byte-for-byte agreement with the compiled reference on reproducible synthetic inputs;
(3) the committed known-answer vectors (tests/golden/kat.json, made by make_kat.py).
CPU only.
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, MODE_ID, make_field, ref_key, ref_mode_cases, ref_table, sha

DTYPES = [np.float32, np.float64, np.int32, np.int64]


@pytest.mark.parametrize("dims", [1, 2, 3, 4])
@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_golden_tables(port, reftest, dtype, dims):
    """Equivalent of the reference's serial end-to-end tests (zfpEndtoendBase.c:339-464)."""
    a = reftest.smooth_field(dtype, dims)
    side = a.shape[0]
    table = ref_table(dtype, dims)
    assert port.hash_array(a) == table[ref_key(0, 0, 0, side, dims)]
    for name, p, mode in ref_mode_cases(dtype):
        words = port.compress(a, **mode)
        assert port.hash_stream(words) == table[ref_key(1, MODE_ID[name], p, side, dims)], (name, p)
        assert reftest.hash_stream(words) == port.hash_stream(words)
        if name == "reversible":
            back = port.decompress(words, a.shape, dtype, **mode)
            assert back.tobytes() == a.tobytes()
        else:
            back = port.decompress(words, a.shape, dtype, **mode)
            assert port.hash_array(back) == table[ref_key(2, MODE_ID[name], p, side, dims)], (name, p)
            if name == "accuracy":
                assert np.max(np.abs(back.astype(np.float64) - a.astype(np.float64))) <= mode["accuracy"]


def _kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def test_kat_vectors(port):
    n = 0
    for c in _kat():
        a = make_field(tuple(c["shape"]), c["dtype"], c["seed"], c["kind"])
        assert sha(a) == c["input"], "input recipe drifted: %r" % (c,)
        mode = dict(c["mode"])
        if "expert" in mode:
            mode["expert"] = tuple(mode["expert"])
        words = port.compress(a, **mode)
        assert words.nbytes == c["nbytes"], c
        assert sha(words) == c["stream"], c
        back = port.decompress(words, a.shape, a.dtype, **mode)
        assert sha(back) == c["decoded"], c
        n += 1
    assert n > 1000


@pytest.mark.parametrize("dtype", DTYPES)
def test_port_vs_reference_strided_and_offset(port, ref, dtype):
    """Negative / permuted / gapped strides and a non-zero start bit (header offset)."""
    rng = np.random.default_rng(7)
    for dims, n in [(1, [37, 0, 0, 0]), (2, [13, 10, 0, 0]), (3, [9, 6, 7, 0]), (4, [5, 6, 4, 7])]:
        total = int(np.prod([v for v in n if v]))
        base = make_field((3 * total + 11,), dtype, seed=dims, kind="smooth")
        layouts = []
        # reversed
        s = [0, 0, 0, 0]
        acc = 1
        for d in range(dims):
            s[d] = -acc
            acc *= n[d]
        layouts.append((s, total - 1 + 5))
        # interleaved (stride 2) with offset
        s = [0, 0, 0, 0]
        acc = 2
        for d in range(dims):
            s[d] = acc
            acc *= n[d]
        layouts.append((s, 3))
        # permuted (transpose of first and last axes)
        if dims > 1:
            order = list(range(dims))[::-1]
            s = [0, 0, 0, 0]
            acc = 1
            for d in order:
                s[d] = acc
                acc *= n[d]
            layouts.append((s, 0))
        for s, off in layouts:
            for mode in ({"rate": 6}, {"precision": 11}, {"reversible": True}):
                for start in (0, 96, 37):
                    prefix = rng.integers(0, 2 ** 63, size=2, dtype=np.uint64)
                    prefix[start // 64:] = 0
                    if start % 64:
                        prefix[start // 64] = rng.integers(0, 2 ** 63, dtype=np.uint64) & np.uint64((1 << (start % 64)) - 1)
                    wr, nbytes = ref.compress_raw(base, off, dtype, n, s, mode, start_bit=start, prefix_words=prefix)
                    wp, end = port.compress_raw(base, off, dtype, n, s, mode, start_bit=start, prefix_words=prefix)
                    assert nbytes == 8 * ((end + 63) // 64)
                    assert wr.tobytes() == wp.tobytes(), (dims, s, mode, start)
                    out_r = np.zeros_like(base)
                    out_p = np.zeros_like(base)
                    ref.decompress_raw(wr, out_r, off, dtype, n, s, mode, start_bit=start)
                    endp = port.decompress_raw(wr, out_p, off, dtype, n, s, mode, start_bit=start)
                    assert out_r.tobytes() == out_p.tobytes()
                    assert endp == end


def test_block_index_lengths(port):
    """The per-block bit lengths the port reports sum to the stream length (variable rate)."""
    a = make_field((20, 21, 22), np.float64, seed=3, kind="smooth")
    for mode in ({"accuracy": 1e-3}, {"precision": 12}, {"reversible": True}, {"rate": 8}):
        words, end, index = port.compress_raw(a.reshape(-1), 0, a.dtype, [22, 21, 20, 0], None, mode, want_index=True)
        assert int(index.astype(np.int64).sum()) == end
        if "rate" in mode:
            assert (index == 512).all()


def test_param_setters_match_reference(port, ref):
    for dtype in DTYPES:
        for dims in (1, 2, 3, 4):
            for mode in ({"rate": 0.3}, {"rate": 1}, {"rate": 8}, {"rate": 8, "align": True}, {"rate": 13.3, "align": True},
                         {"rate": 64}, {"precision": 0}, {"precision": 1}, {"precision": 33}, {"precision": 100},
                         {"accuracy": 0}, {"accuracy": 1e-6}, {"accuracy": 3.0}, {"accuracy": 1e300}, {"reversible": True}):
                assert port.params(mode, dtype, dims) == ref.params(mode, dtype, dims), (dtype, dims, mode)
                for n in ([17, 0, 0, 0], [17, 5, 0, 0], [17, 5, 9, 0], [17, 5, 9, 6]):
                    if sum(1 for v in n if v) == dims:
                        assert port.maximum_size(mode, dtype, n) == ref.maximum_size(mode, dtype, n)


def test_parallel_cpu_decompress_equals_serial(ref):
    """Row f4 of the scope table: the chunk-parallel fixed-rate CPU decompress (an OpenMP driver around the
    reference's own block API, oracle/ref_parallel_decompress.c) reproduces the reference's serial
    zfp_decompress bit for bit - every type, 1-4 D, partial blocks, 1 and 4 threads."""
    from oracle.oracle import REF_PDEC_SO, parallel_decompress
    if not os.path.exists(REF_PDEC_SO):
        pytest.skip("oracle/_ref/libzfp_ref_pdec.so not built")
    for dtype in (np.float32, np.float64, np.int32, np.int64):
        for shape in ((37,), (14, 19), (9, 10, 13), (5, 6, 7, 9)):
            a = make_field(shape, dtype, seed=17, kind="smooth")
            for rate in (4, 9.5, 20):
                mode = {"rate": rate}
                words = ref.compress(a, **mode)
                want = ref.decompress(words, a.shape, a.dtype, **mode)
                maxbits = ref.params(mode, a.dtype, a.ndim)[1]
                for threads in (1, 4):
                    got, used = parallel_decompress(np.concatenate([words, np.zeros(2, dtype=np.uint64)]), a.shape, a.dtype, maxbits, threads)
                    assert got.tobytes() == want.tobytes(), (np.dtype(dtype).name, shape, rate, threads)
                    assert (used + 63) // 64 * 8 == words.nbytes
