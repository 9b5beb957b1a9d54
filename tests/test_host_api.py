"""CPU-only tests of the product library's host side: the C-ABI library loads, exports every
symbol the headers declare, and its parameter / mode / header / field / bit-stream arithmetic
agrees with the unmodified reference (oracle/_ref).  No compute call is made (no GPU here)."""
import ctypes as C
import re
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DTYPES = [np.float32, np.float64, np.int32, np.int64]
TYPE_ID = {np.int32: 1, np.int64: 2, np.float32: 3, np.float64: 4}


@pytest.fixture(scope="module")
def lib():
    from zfp_b200 import api
    return api.load_library()


def test_library_exports_every_declared_symbol(lib):
    from zfp_b200 import api
    declared = set()
    for h in ("zfp_b200.h", "zfp_b200_backend.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        declared |= set(re.findall(r"\b((?:zfp|stream|cuda)_\w+)\s*\(", txt))
    declared -= {"zfp_exec_params_cuda"}
    assert len(declared) > 90
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(api.EXPORTED_SYMBOLS) - {"stream_word_bits", "zfp_codec_version", "zfp_library_version", "zfp_version_string"} <= declared
    assert C.c_size_t.in_dll(lib, "stream_word_bits").value == 64
    assert C.c_uint.in_dll(lib, "zfp_codec_version").value == 5


def _modes():
    return [{"rate": 0.3}, {"rate": 1}, {"rate": 8}, {"rate": 8, "align": True}, {"rate": 13.3, "align": True}, {"rate": 64},
            {"precision": 0}, {"precision": 1}, {"precision": 33}, {"precision": 100}, {"accuracy": 0}, {"accuracy": 1e-6},
            {"accuracy": 3.0}, {"accuracy": 1e300}, {"reversible": True}, {"expert": (5, 300, 40, -20)},
            {"expert": (1, 16658, 64, -1074)}, {"expert": (100, 40000, 64, -1080)}]


def test_parameters_modes_and_sizes_match_reference(lib, ref):
    from zfp_b200.api import _set_mode
    R = ref.L
    for dtype in DTYPES:
        for dims in (1, 2, 3, 4):
            for mode in _modes():
                za, zr = lib.zfp_stream_open(None), R.zfp_stream_open(None)
                _set_mode(lib, za, mode, TYPE_ID[dtype], dims)
                ref._set_mode(zr, mode, dtype, dims)
                pa = [C.c_uint(), C.c_uint(), C.c_uint(), C.c_int()]
                lib.zfp_stream_params(za, *[C.byref(v) for v in pa])
                assert tuple(v.value for v in pa) == ref.params(mode, dtype, dims)
                assert lib.zfp_stream_compression_mode(za) == R.zfp_stream_compression_mode(zr)
                assert lib.zfp_stream_mode(za) == R.zfp_stream_mode(zr)
                # compact mode round trip through both libraries
                zb2 = lib.zfp_stream_open(None)
                assert lib.zfp_stream_set_mode(zb2, R.zfp_stream_mode(zr)) == R.zfp_stream_compression_mode(zr)
                pb = [C.c_uint(), C.c_uint(), C.c_uint(), C.c_int()]
                lib.zfp_stream_params(zb2, *[C.byref(v) for v in pb])
                zr2 = R.zfp_stream_open(None)
                R.zfp_stream_set_mode(zr2, R.zfp_stream_mode(zr))
                pr = [C.c_uint(), C.c_uint(), C.c_uint(), C.c_int()]
                R.zfp_stream_params(zr2, *[C.byref(v) for v in pr])
                assert [v.value for v in pb] == [v.value for v in pr]
                for n in ([17, 0, 0, 0], [17, 5, 0, 0], [17, 5, 9, 0], [17, 5, 9, 6]):
                    if sum(1 for v in n if v) != dims:
                        continue
                    fa = getattr(lib, "zfp_field_%dd" % dims)(None, TYPE_ID[dtype], *n[:dims])
                    assert lib.zfp_stream_maximum_size(za, fa) == ref.maximum_size(mode, dtype, n)
                    fr, _ = ref._field(None, dtype, n, None)
                    assert lib.zfp_field_metadata(fa) == R.zfp_field_metadata(fr)
                    assert lib.zfp_field_blocks(fa) == int(np.prod([(v + 3) // 4 for v in n if v]))
                    lib.zfp_field_free(fa)
                    R.zfp_field_free(fr)
                for z, L in ((za, lib), (zb2, lib), (zr, R), (zr2, R)):
                    L.zfp_stream_close(z)


def test_execution_policy_contract(lib):
    z = lib.zfp_stream_open(None)
    assert lib.zfp_stream_execution(z) == 0
    assert lib.zfp_stream_set_execution(z, 2) == 1 and lib.zfp_stream_execution(z) == 2
    assert lib.zfp_stream_set_execution(z, 1) == 0  # OpenMP is not provided by this library
    assert lib.zfp_stream_set_execution(z, 7) == 0
    p = lib.zfp_stream_cuda_params(z)
    assert p and p.contents.magic == 0x7a66704232303021
    assert lib.zfp_stream_set_execution(z, 0) == 1
    assert not lib.zfp_stream_cuda_params(z)
    # serial policy: the array path of this library refuses (no CPU fallback), returns 0
    buf = np.zeros(64, dtype=np.uint64)
    a = np.ones((4, 4, 4))
    bs = lib.stream_open(buf.ctypes.data, buf.nbytes)
    lib.zfp_stream_set_bit_stream(z, bs)
    f = lib.zfp_field_3d(a.ctypes.data, 4, 4, 4, 4)
    assert lib.zfp_compress(z, f) == 0 and lib.zfp_decompress(z, f) == 0
    assert lib.stream_wtell(bs) == 0
    lib.zfp_field_free(f)
    lib.zfp_stream_close(z)
    lib.stream_close(bs)


def test_bitstream_matches_reference(lib, ref):
    """Random sequences of bit-stream operations give identical buffers and cursors."""
    R = ref.L
    R.stream_write_bits.restype = C.c_uint64
    R.stream_write_bits.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t]
    R.stream_read_bits.restype = C.c_uint64
    R.stream_read_bits.argtypes = [C.c_void_p, C.c_size_t]
    R.stream_pad.argtypes = [C.c_void_p, C.c_uint64]
    R.stream_skip.argtypes = [C.c_void_p, C.c_uint64]
    R.stream_flush.restype = C.c_size_t
    R.stream_flush.argtypes = [C.c_void_p]
    R.stream_align.restype = C.c_size_t
    R.stream_align.argtypes = [C.c_void_p]
    R.stream_size.restype = C.c_size_t
    R.stream_size.argtypes = [C.c_void_p]
    rng = np.random.default_rng(11)
    for trial in range(20):
        ba, br = np.zeros(256, dtype=np.uint64), np.zeros(256, dtype=np.uint64)
        sa, sr = lib.stream_open(ba.ctypes.data, ba.nbytes), R.stream_open(br.ctypes.data, br.nbytes)
        start = int(rng.integers(0, 130))
        lib.stream_wseek(sa, start)
        R.stream_wseek(sr, start)
        for _ in range(200):
            op = rng.integers(0, 10)
            if op < 7:
                n = int(rng.integers(0, 65))
                v = int(rng.integers(0, 2 ** 63)) * 2 + int(rng.integers(0, 2))
                assert lib.stream_write_bits(sa, v, n) == R.stream_write_bits(sr, v, n)
            elif op < 9:
                n = int(rng.integers(0, 150))
                lib.stream_pad(sa, n)
                R.stream_pad(sr, n)
            else:
                assert lib.stream_flush(sa) == R.stream_flush(sr)
            assert lib.stream_wtell(sa) == R.stream_wtell(sr)
        assert lib.stream_flush(sa) == R.stream_flush(sr)
        assert lib.stream_size(sa) == R.stream_size(sr)
        assert ba.tobytes() == br.tobytes()
        data = rng.integers(0, 2 ** 63, size=256, dtype=np.uint64)
        ba[:], br[:] = data, data
        start = int(rng.integers(0, 130))
        lib.stream_rseek(sa, start)
        R.stream_rseek(sr, start)
        for _ in range(200):
            op = rng.integers(0, 10)
            if op < 7:
                n = int(rng.integers(0, 65))
                assert lib.stream_read_bits(sa, n) == R.stream_read_bits(sr, n)
            elif op < 9:
                n = int(rng.integers(0, 150))
                lib.stream_skip(sa, n)
                R.stream_skip(sr, n)
            else:
                assert lib.stream_align(sa) == R.stream_align(sr)
            assert lib.stream_rtell(sa) == R.stream_rtell(sr)
        lib.stream_close(sa)
        R.stream_close(sr)


def test_header_matches_reference(lib, ref):
    from zfp_b200.api import _set_mode
    R = ref.L
    for dtype, n in ((np.float64, [100, 200, 30, 0]), (np.float32, [7, 0, 0, 0]), (np.int32, [9, 8, 7, 6]), (np.int64, [4000, 5000, 0, 0])):
        dims = sum(1 for v in n if v)
        for mode in _modes():
            ba, br = np.zeros(8, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
            sa, sr = lib.stream_open(ba.ctypes.data, ba.nbytes), R.stream_open(br.ctypes.data, br.nbytes)
            za, zr = lib.zfp_stream_open(sa), R.zfp_stream_open(sr)
            _set_mode(lib, za, mode, TYPE_ID[dtype], dims)
            ref._set_mode(zr, mode, dtype, dims)
            fa = getattr(lib, "zfp_field_%dd" % dims)(None, TYPE_ID[dtype], *n[:dims])
            fr, _ = ref._field(None, dtype, n, None)
            assert lib.zfp_write_header(za, fa, 7) == R.zfp_write_header(zr, fr, 7)
            lib.zfp_stream_flush(za)
            R.zfp_stream_flush.argtypes = [C.c_void_p]
            R.zfp_stream_flush(zr)
            assert ba.tobytes() == br.tobytes(), (dtype, mode)
            # read it back with our library into a blank stream/field
            zb2 = lib.zfp_stream_open(sa)
            lib.zfp_stream_rewind(zb2)
            fb = lib.zfp_field_alloc()
            bits = lib.zfp_read_header(zb2, fb, 7)
            assert bits > 0 and lib.zfp_field_metadata(fb) == R.zfp_field_metadata(fr)
            assert lib.zfp_stream_mode(zb2) == R.zfp_stream_mode(zr)
            for f in (fa, fb):
                lib.zfp_field_free(f)
            R.zfp_field_free(fr)
            lib.zfp_stream_close(za)
            lib.zfp_stream_close(zb2)
            R.zfp_stream_close(zr)
            lib.stream_close(sa)
            R.stream_close(sr)


def test_field_accessors(lib):
    a = np.zeros(1000)
    f = lib.zfp_field_3d(a.ctypes.data + 8 * 500, 4, 5, 6, 7)
    assert lib.zfp_field_dimensionality(f) == 3 and lib.zfp_field_precision(f) == 64 and lib.zfp_field_type(f) == 4
    assert lib.zfp_field_size(f, None) == 210 and lib.zfp_field_size_bytes(f) == 1680 and lib.zfp_field_is_contiguous(f) == 1
    st = (C.c_ssize_t * 4)()
    assert lib.zfp_field_stride(f, st) == 0 and list(st)[:3] == [1, 5, 30]
    lib.zfp_field_set_stride_3d(f, -1, -5, -30)
    assert lib.zfp_field_stride(f, st) == 1 and list(st)[:3] == [-1, -5, -30]
    assert lib.zfp_field_begin(f) == a.ctypes.data + 8 * (500 - 209) and lib.zfp_field_is_contiguous(f) == 1
    lib.zfp_field_set_stride_3d(f, 2, 10, 60)
    assert lib.zfp_field_is_contiguous(f) == 0 and lib.zfp_field_size_bytes(f) == 8 * (2 * 4 + 10 * 5 + 60 * 6 + 1)
    lib.zfp_field_free(f)


def test_headers_are_plain_c_and_a_c_program_links(tmp_path, lib):
    """The boundary is a C ABI: every header under include/ compiles as C99 and as C++17 with no CUDA or
    torch in sight, and a C program written against zfp.h (the reference's own include name) links
    against libzfp_b200.so and can set up a stream without touching the GPU."""
    import shutil
    import subprocess
    from zfp_b200 import build as _build
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    inc = os.path.join(ROOT, "include")
    for h in ("zfp.h", "zfp_b200.h", "zfp_b200_backend.h"):
        subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c", os.path.join(inc, h)], check=True)
        subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-I", inc, "-x", "c++", os.path.join(inc, h)], check=True)
    src = tmp_path / "prog.c"
    src.write_text(r'''
#include <stdio.h>
#include "zfp.h"
#include "zfp_b200_backend.h"
int main(void)
{
  double a[64] = { 0 };
  zfp_field* f = zfp_field_3d(a, zfp_type_double, 4, 4, 4);
  zfp_stream* z = zfp_stream_open(NULL);
  double rate = zfp_stream_set_rate(z, 8.0, zfp_type_double, 3, zfp_false);
  size_t cap = zfp_stream_maximum_size(z, f);
  int ok = zfp_stream_set_execution(z, zfp_exec_cuda);
  zfp_b200_desc d = { zfp_type_double, 3, { 4, 4, 4, 0 }, { 0, 0, 0, 0 }, 512, 512, 64, -1074 };
  printf("%g %zu %d %zu %d\n", rate, cap, ok, zfp_b200_blocks(&d), zfp_b200_is_fixed_rate(&d));
  zfp_field_free(f);
  zfp_stream_close(z);
  return 0;
}
''')
    exe = tmp_path / "prog"
    libdir = os.path.dirname(_build.LIB)
    subprocess.run([cc, "-std=c99", "-Wall", "-I", inc, "-o", str(exe), str(src), "-L", libdir, "-lzfp_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["8", "88", "1", "1", "1"], out   # (148 + 512 + 63) & ~63 bits = 704 bits = 88 bytes


def test_box_block_ranges_against_brute_force():
    """Host logic of the coordinate-box random access: the block ranges cover exactly the blocks that
    intersect the box (stream order: last dimension fastest)."""
    import itertools
    from zfp_b200.api import box_block_ranges
    rng = np.random.default_rng(7)
    for _ in range(200):
        dims = int(rng.integers(1, 5))
        shape = tuple(int(rng.integers(1, 23)) for _ in range(dims))
        lo = [int(rng.integers(-2, n + 2)) for n in shape]
        hi = [int(rng.integers(l, n + 4)) for l, n in zip(lo, shape)]
        nb = [(n + 3) // 4 for n in shape]
        want = set()
        for blk in itertools.product(*[range(m) for m in nb]):
            hit = all(4 * b < min(h, n) and 4 * b + 4 > max(l, 0) for b, l, h, n in zip(blk, lo, hi, shape))
            if hit:
                lin = 0
                for b, m in zip(blk, nb):
                    lin = lin * m + b
                want.add(lin)
        got = set()
        for b0, b1 in box_block_ranges(shape, lo, hi):
            assert 0 <= b0 < b1 <= int(np.prod(nb))
            got |= set(range(b0, b1))
        assert got == want, (shape, lo, hi)


def test_refused_parameter_regimes(lib):
    """The two parameter regimes the backend refuses instead of reproducing:
    a maxbits below the block header (the reference overruns zfp_stream_maximum_size there,
    src/template/encodef.c:77-82) and variable-rate minbits beyond the 16-bit block lengths of the index.
    Both are rejected before any CUDA call, with a message; ordinary parameters pass the same check."""
    from zfp_b200 import api
    dummy = C.c_void_p(0x1000)  # never dereferenced: the parameter check comes first
    end = C.c_uint64()

    def encode(ztype, dims, params):
        d = api.Desc()
        d.type, d.dims = ztype, dims
        for i in range(dims):
            d.n[i] = 8
        d.minbits, d.maxbits, d.maxprec, d.minexp = params
        return lib.zfp_b200_encode(C.byref(d), dummy, dummy, 0, C.byref(end), None, None)

    EINVAL = 1
    for ztype, header in ((3, 9), (4, 12)):                # float, double: 1 + exponent bits
        assert encode(ztype, 3, (1, header - 1, 64, -1074)) == EINVAL
        assert b"block header" in lib.zfp_b200_last_error()
        assert lib.zfp_stream_maximum_size  # (the setter itself accepts such values, as upstream does)
    assert encode(4, 3, (1, 18, 64, -1075)) == EINVAL      # reversible double: 1 + 1 + 11 + 6 = 19 header bits
    assert encode(1, 2, (70000, 80000, 32, -1074)) == EINVAL
    assert b"minbits" in lib.zfp_b200_last_error()
    # fixed rate with huge blocks is fine for the check (minbits == maxbits): the call gets past it and only
    # then fails on the fake pointers - which must not happen here, so ask for a descriptor error instead
    d = api.Desc()
    d.type, d.dims = 4, 0
    assert lib.zfp_b200_encode(C.byref(d), dummy, dummy, 0, C.byref(end), None, None) == EINVAL
    assert b"descriptor" in lib.zfp_b200_last_error()
