"""ctypes front ends for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Two checkers with one interface:

* ``Port``       - oracle/libzfp_oracle.so, our C restatement (zfp_oracle.c).
* ``Reference``  - oracle/_ref/libzfp_ref.so, the unmodified reference library compiled from
                   /root/reference by oracle/Makefile (serial and OpenMP execution policies).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under zfp_b200/ does.

Arrays follow zfpy's convention (reference python/zfpy.pyx:140-187): a C-ordered numpy array
of shape (nz, ny, nx) is a zfp field with nx = shape[-1] fastest.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libzfp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libzfp_ref.so")
REFTEST_SO = os.path.join(HERE, "_ref", "libzfp_reftest.so")

ZFP_TYPE = {np.dtype(np.int32): 1, np.dtype(np.int64): 2, np.dtype(np.float32): 3, np.dtype(np.float64): 4}
ZFP_MIN_EXP = -1074
ZFP_MAX_BITS = 16658


REF_CUDA_SO = os.path.join(HERE, "_ref", "libzfp_ref_cuda.so")
REF_PDEC_SO = os.path.join(HERE, "_ref", "libzfp_ref_pdec.so")  # chunk-parallel fixed-rate CPU decompress (row f4)
REF_CUDA_ALL_SO = os.path.join(HERE, "_ref", "libzfp_ref_cuda_all.so")    # reference + integration/zfp_cuda_dispatch.patch on our backend
REF_CUDA_ORIG_SO = os.path.join(HERE, "_ref", "libzfp_ref_cudaorig.so")   # reference with ITS OWN src/cuda_zfp (the backend replaced)


def build(ref=True):
    """(Re)build the oracle libraries with oracle/Makefile."""
    have_ref = ref and os.path.isdir(os.environ.get("ZFP_REFERENCE", "/root/reference"))
    targets = ["port"] + (["ref", "ref_pdec"] if have_ref else [])
    if have_ref and os.path.exists(os.path.join(HERE, "..", "zfp_b200", "lib", "libzfp_b200.so")):
        targets += ["ref_cuda", "ref_cli", "ref_cuda_all", "ref_cudaorig", "b200_cli"]
    subprocess.check_call(["make", "-s", "-C", HERE] + targets)


class _Params(C.Structure):
    _fields_ = [("type", C.c_int), ("dims", C.c_int), ("n", C.c_size_t * 4), ("s", C.c_ssize_t * 4),
                ("minbits", C.c_uint), ("maxbits", C.c_uint), ("maxprec", C.c_uint), ("minexp", C.c_int)]


def mode_params(mode, dtype, dims):
    """(minbits, maxbits, maxprec, minexp) for a mode dict, using the PORT's restated setters."""
    return Port().params(mode, dtype, dims)


def _shape_to_n(shape):
    n = list(reversed(shape)) + [0] * (4 - len(shape))
    return n


class Port:
    """The C restatement."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        L = C.CDLL(PORT_SO)
        L.zo_compress.restype = C.c_uint64
        L.zo_compress.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.zo_decompress.restype = C.c_uint64
        L.zo_decompress.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_uint64]
        L.zo_maximum_size.restype = C.c_size_t
        L.zo_maximum_size.argtypes = [C.POINTER(_Params)]
        L.zo_set_rate.argtypes = [C.POINTER(_Params), C.c_double, C.c_int]
        L.zo_set_precision.argtypes = [C.POINTER(_Params), C.c_uint]
        L.zo_set_accuracy.argtypes = [C.POINTER(_Params), C.c_double]
        L.zo_set_reversible.argtypes = [C.POINTER(_Params)]
        L.zo_hash_words64.restype = C.c_uint64
        L.zo_hash_words64.argtypes = [C.c_void_p, C.c_size_t]
        L.zo_hash_words32.restype = C.c_uint32
        L.zo_hash_words32.argtypes = [C.c_void_p, C.c_size_t]
        self.L = L

    # -- parameters -----------------------------------------------------------------------
    def _p(self, mode, dtype, n, s=None):
        p = _Params()
        p.type = ZFP_TYPE[np.dtype(dtype)]
        p.dims = sum(1 for v in n if v)
        for i in range(4):
            p.n[i] = n[i]
            p.s[i] = s[i] if s else 0
        if "rate" in mode:
            self.L.zo_set_rate(C.byref(p), float(mode["rate"]), int(bool(mode.get("align", False))))
        elif "precision" in mode:
            self.L.zo_set_precision(C.byref(p), int(mode["precision"]))
        elif "accuracy" in mode:
            self.L.zo_set_accuracy(C.byref(p), float(mode["accuracy"]))
        elif mode.get("reversible"):
            self.L.zo_set_reversible(C.byref(p))
        elif "expert" in mode:
            p.minbits, p.maxbits, p.maxprec, p.minexp = mode["expert"]
        else:
            raise ValueError("unknown mode %r" % (mode,))
        return p

    def params(self, mode, dtype, dims):
        p = self._p(mode, dtype, [4] * dims + [0] * (4 - dims))
        return p.minbits, p.maxbits, p.maxprec, p.minexp

    def maximum_size(self, mode, dtype, n):
        return self.L.zo_maximum_size(C.byref(self._p(mode, dtype, n)))

    # -- codec ----------------------------------------------------------------------------
    def compress_raw(self, buf, offset, dtype, n, s, mode, start_bit=0, prefix_words=None, want_index=False):
        """Compress the field at element `offset` of 1-D array `buf` with sizes n / strides s."""
        p = self._p(mode, dtype, n, s)
        cap = self.L.zo_maximum_size(C.byref(p)) // 8 + (start_bit + 63) // 64 + 2
        words = np.zeros(cap, dtype=np.uint64)
        if prefix_words is not None:
            words[:len(prefix_words)] = prefix_words
        nblocks = 1
        for v in n:
            nblocks *= (v + 3) // 4 if v else 1
        index = np.zeros(nblocks, dtype=np.uint16) if want_index else None
        base = buf.ctypes.data + offset * buf.dtype.itemsize
        end = self.L.zo_compress(C.byref(p), base, words.ctypes.data, start_bit,
                                 index.ctypes.data if want_index else None)
        out = words[:(end + 63) // 64].copy()
        return (out, end, index) if want_index else (out, end)

    def decompress_raw(self, words, buf, offset, dtype, n, s, mode, start_bit=0):
        p = self._p(mode, dtype, n, s)
        padded = np.concatenate([np.ascontiguousarray(words, dtype=np.uint64), np.zeros(4, dtype=np.uint64)])
        base = buf.ctypes.data + offset * buf.dtype.itemsize
        return self.L.zo_decompress(C.byref(p), base, padded.ctypes.data, start_bit)

    def compress(self, a, **mode):
        a = np.ascontiguousarray(a)
        words, _ = self.compress_raw(a.reshape(-1), 0, a.dtype, _shape_to_n(a.shape), None, mode)
        return words

    def decompress(self, words, shape, dtype, **mode):
        out = np.empty(shape, dtype=dtype)
        self.decompress_raw(words, out.reshape(-1), 0, dtype, _shape_to_n(shape), None, mode)
        return out

    # -- hashes used by the reference's golden tables -----------------------------------------
    def hash_stream(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        return self.L.zo_hash_words64(words.ctypes.data, words.size)

    def hash_array(self, a):
        a = np.ascontiguousarray(a)
        if a.dtype.itemsize == 4:
            return self.L.zo_hash_words32(a.ctypes.data, a.size)
        return self.L.zo_hash_words64(a.ctypes.data, a.size)


class Reference:
    """The unmodified reference library (serial / OpenMP), driven through its public C API
    exactly as utils/zfp.c does (reference utils/zfp.c:372-593)."""

    SERIAL, OMP = 0, 1

    def __init__(self, so=None):
        so = so or REF_SO
        if not os.path.exists(so):
            raise FileNotFoundError(so + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(so)
        vp, sz = C.c_void_p, C.c_size_t
        L.stream_open.restype = vp
        L.stream_open.argtypes = [vp, sz]
        L.stream_close.argtypes = [vp]
        L.stream_wseek.argtypes = [vp, C.c_uint64]
        L.stream_rseek.argtypes = [vp, C.c_uint64]
        L.stream_wtell.restype = C.c_uint64
        L.stream_wtell.argtypes = [vp]
        L.stream_rtell.restype = C.c_uint64
        L.stream_rtell.argtypes = [vp]
        L.zfp_stream_open.restype = vp
        L.zfp_stream_open.argtypes = [vp]
        L.zfp_stream_close.argtypes = [vp]
        L.zfp_stream_rewind.argtypes = [vp]
        L.zfp_stream_set_bit_stream.argtypes = [vp, vp]
        L.zfp_stream_set_rate.restype = C.c_double
        L.zfp_stream_set_rate.argtypes = [vp, C.c_double, C.c_int, C.c_uint, C.c_int]
        L.zfp_stream_set_precision.restype = C.c_uint
        L.zfp_stream_set_precision.argtypes = [vp, C.c_uint]
        L.zfp_stream_set_accuracy.restype = C.c_double
        L.zfp_stream_set_accuracy.argtypes = [vp, C.c_double]
        L.zfp_stream_set_reversible.argtypes = [vp]
        L.zfp_stream_set_params.restype = C.c_int
        L.zfp_stream_set_params.argtypes = [vp, C.c_uint, C.c_uint, C.c_uint, C.c_int]
        L.zfp_stream_params.argtypes = [vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_int)]
        L.zfp_stream_set_execution.restype = C.c_int
        L.zfp_stream_set_execution.argtypes = [vp, C.c_int]
        L.zfp_stream_set_omp_threads.restype = C.c_int
        L.zfp_stream_set_omp_threads.argtypes = [vp, C.c_uint]
        L.zfp_stream_maximum_size.restype = sz
        L.zfp_stream_maximum_size.argtypes = [vp, vp]
        L.zfp_stream_mode.restype = C.c_uint64
        L.zfp_stream_mode.argtypes = [vp]
        L.zfp_stream_set_mode.restype = C.c_int
        L.zfp_stream_set_mode.argtypes = [vp, C.c_uint64]
        L.zfp_stream_compression_mode.restype = C.c_int
        L.zfp_stream_compression_mode.argtypes = [vp]
        L.zfp_field_alloc.restype = vp
        L.zfp_field_free.argtypes = [vp]
        L.zfp_field_set_pointer.argtypes = [vp, vp]
        L.zfp_field_set_type.restype = C.c_int
        L.zfp_field_set_type.argtypes = [vp, C.c_int]
        L.zfp_field_set_size_1d.argtypes = [vp, sz]
        L.zfp_field_set_size_2d.argtypes = [vp, sz, sz]
        L.zfp_field_set_size_3d.argtypes = [vp, sz, sz, sz]
        L.zfp_field_set_size_4d.argtypes = [vp, sz, sz, sz, sz]
        ss = C.c_ssize_t
        L.zfp_field_set_stride_1d.argtypes = [vp, ss]
        L.zfp_field_set_stride_2d.argtypes = [vp, ss, ss]
        L.zfp_field_set_stride_3d.argtypes = [vp, ss, ss, ss]
        L.zfp_field_set_stride_4d.argtypes = [vp, ss, ss, ss, ss]
        L.zfp_field_metadata.restype = C.c_uint64
        L.zfp_field_metadata.argtypes = [vp]
        L.zfp_compress.restype = sz
        L.zfp_compress.argtypes = [vp, vp]
        L.zfp_decompress.restype = sz
        L.zfp_decompress.argtypes = [vp, vp]
        L.zfp_write_header.restype = sz
        L.zfp_write_header.argtypes = [vp, vp, C.c_uint]
        L.zfp_read_header.restype = sz
        L.zfp_read_header.argtypes = [vp, vp, C.c_uint]
        self.L = L

    def _field(self, ptr, dtype, n, s):
        L = self.L
        f = L.zfp_field_alloc()
        L.zfp_field_set_type(f, ZFP_TYPE[np.dtype(dtype)])
        L.zfp_field_set_pointer(f, ptr)
        dims = sum(1 for v in n if v)
        getattr(L, "zfp_field_set_size_%dd" % dims)(f, *n[:dims])
        if s and any(s):
            getattr(L, "zfp_field_set_stride_%dd" % dims)(f, *s[:dims])
        return f, dims

    def _set_mode(self, z, mode, dtype, dims):
        L = self.L
        if "rate" in mode:
            L.zfp_stream_set_rate(z, float(mode["rate"]), ZFP_TYPE[np.dtype(dtype)], dims, int(bool(mode.get("align", False))))
        elif "precision" in mode:
            L.zfp_stream_set_precision(z, int(mode["precision"]))
        elif "accuracy" in mode:
            L.zfp_stream_set_accuracy(z, float(mode["accuracy"]))
        elif mode.get("reversible"):
            L.zfp_stream_set_reversible(z)
        elif "expert" in mode:
            assert L.zfp_stream_set_params(z, *mode["expert"])
        else:
            raise ValueError("unknown mode %r" % (mode,))

    def params(self, mode, dtype, dims):
        z = self.L.zfp_stream_open(None)
        self._set_mode(z, mode, dtype, dims)
        a, b, c, d = C.c_uint(), C.c_uint(), C.c_uint(), C.c_int()
        self.L.zfp_stream_params(z, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.L.zfp_stream_close(z)
        return a.value, b.value, c.value, d.value

    def maximum_size(self, mode, dtype, n):
        z = self.L.zfp_stream_open(None)
        f, dims = self._field(None, dtype, n, None)
        self._set_mode(z, mode, dtype, dims)
        r = self.L.zfp_stream_maximum_size(z, f)
        self.L.zfp_field_free(f)
        self.L.zfp_stream_close(z)
        return r

    def compress_raw(self, buf, offset, dtype, n, s, mode, start_bit=0, prefix_words=None, policy=0, threads=0,
                     header_mask=0, out=None):
        L = self.L
        f, dims = self._field(buf.ctypes.data + offset * buf.dtype.itemsize, dtype, n, s)
        z = L.zfp_stream_open(None)
        self._set_mode(z, mode, dtype, dims)
        cap = L.zfp_stream_maximum_size(z, f) + 8 * ((start_bit + 63) // 64 + 2)
        words = out if out is not None else np.zeros(cap // 8, dtype=np.uint64)
        if prefix_words is not None:
            words[:len(prefix_words)] = prefix_words
        bs = L.stream_open(words.ctypes.data, words.nbytes)
        L.zfp_stream_set_bit_stream(z, bs)
        if policy:
            assert L.zfp_stream_set_execution(z, policy)
            if threads:
                L.zfp_stream_set_omp_threads(z, threads)
        L.stream_wseek(bs, start_bit)
        if header_mask:
            assert L.zfp_write_header(z, f, header_mask)
        nbytes = L.zfp_compress(z, f)
        L.zfp_field_free(f)
        L.zfp_stream_close(z)
        L.stream_close(bs)
        return words[:nbytes // 8], nbytes

    def decompress_raw(self, words, buf, offset, dtype, n, s, mode, start_bit=0):
        L = self.L
        words = np.concatenate([np.ascontiguousarray(words, dtype=np.uint64), np.zeros(4, dtype=np.uint64)])
        f, dims = self._field(buf.ctypes.data + offset * buf.dtype.itemsize, dtype, n, s)
        bs = L.stream_open(words.ctypes.data, words.nbytes)
        z = L.zfp_stream_open(bs)
        self._set_mode(z, mode, dtype, dims)
        L.stream_rseek(bs, start_bit)
        nbytes = L.zfp_decompress(z, f)
        L.zfp_field_free(f)
        L.zfp_stream_close(z)
        L.stream_close(bs)
        return nbytes

    def decompress_raw_noalloc(self, words, buf, dtype, n, mode):
        """decompress_raw without the defensive copy of the stream (for timing); `words` must be
        followed by at least one readable word (true for views of a capacity-sized buffer)."""
        L = self.L
        f, dims = self._field(buf.ctypes.data, dtype, n, None)
        bs = L.stream_open(words.ctypes.data, words.nbytes)
        z = L.zfp_stream_open(bs)
        self._set_mode(z, mode, dtype, dims)
        nbytes = L.zfp_decompress(z, f)
        L.zfp_field_free(f)
        L.zfp_stream_close(z)
        L.stream_close(bs)
        return nbytes

    def decompress_with_header(self, data):
        """What the reference's zfpy.decompress_numpy does (python/zfpy.pyx:300-340): read the full
        header, size the array from it, decompress serially."""
        L = self.L
        L.zfp_field_size.restype = C.c_size_t
        L.zfp_field_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.zfp_field_type.restype = C.c_int
        L.zfp_field_type.argtypes = [C.c_void_p]
        L.zfp_field_dimensionality.restype = C.c_uint
        L.zfp_field_dimensionality.argtypes = [C.c_void_p]
        words = np.zeros((len(data) + 15) // 8 + 1, dtype=np.uint64)
        words.view(np.uint8)[: len(data)] = np.frombuffer(data, dtype=np.uint8)
        bs = L.stream_open(words.ctypes.data, words.nbytes)
        z = L.zfp_stream_open(bs)
        f = L.zfp_field_alloc()
        assert L.zfp_read_header(z, f, 7), "reference could not read the header"
        dims = L.zfp_field_dimensionality(f)
        size = (C.c_size_t * 4)()
        L.zfp_field_size(f, size)
        shape = tuple(reversed([int(size[i]) for i in range(dims)]))
        inv = {v: k for k, v in ZFP_TYPE.items()}
        out = np.empty(shape, dtype=inv[L.zfp_field_type(f)])
        L.zfp_field_set_pointer(f, out.ctypes.data)
        nbytes = L.zfp_decompress(z, f)
        L.zfp_field_free(f)
        L.zfp_stream_close(z)
        L.stream_close(bs)
        return out, nbytes

    def compress(self, a, policy=0, threads=0, **mode):
        a = np.ascontiguousarray(a)
        words, _ = self.compress_raw(a.reshape(-1), 0, a.dtype, _shape_to_n(a.shape), None, mode,
                                     policy=policy, threads=threads)
        return words.copy()

    def decompress(self, words, shape, dtype, **mode):
        out = np.empty(shape, dtype=dtype)
        self.decompress_raw(words, out.reshape(-1), 0, dtype, _shape_to_n(shape), None, mode)
        return out


def parallel_decompress(words, shape, dtype, maxbits, threads, out=None, start_bit=0):
    """Chunk-parallel FIXED-RATE decompression on the host: oracle/ref_parallel_decompress.c,
    an OpenMP driver around the reference's own block API, for the CPU column of the baseline table."""
    if not os.path.exists(REF_PDEC_SO):
        raise FileNotFoundError(REF_PDEC_SO + " (run `make -C oracle ref_pdec` where /root/reference exists)")
    L = C.CDLL(REF_PDEC_SO)
    L.zfp_ref_parallel_decompress.restype = C.c_uint64
    L.zfp_ref_parallel_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.POINTER(C.c_size_t), C.c_uint,
                                              C.c_void_p, C.c_int]
    out = np.empty(shape, dtype=dtype) if out is None else out
    n = (C.c_size_t * 4)(*(list(reversed(shape)) + [0] * (4 - len(shape))))
    words = np.ascontiguousarray(words, dtype=np.uint64)
    used = L.zfp_ref_parallel_decompress(words.ctypes.data, words.nbytes, start_bit, ZFP_TYPE[np.dtype(dtype)], n, int(maxbits),
                                         out.ctypes.data, int(threads))
    return out, int(used)


class RefTestUtils:
    """The reference's own seeded smooth-field generator (tests/utils/genSmoothRandNums.c:863-923)
    and Jenkins hash (tests/utils/zfpHash.c), compiled unmodified into oracle/_ref."""

    def __init__(self):
        if not os.path.exists(REFTEST_SO):
            raise FileNotFoundError(REFTEST_SO)
        self.L = C.CDLL(REFTEST_SO)
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        self.L.hashBitstream.restype = C.c_uint64
        self.L.hashBitstream.argtypes = [C.c_void_p, C.c_size_t]

    def smooth_field(self, dtype, dims):
        """The array the reference's end-to-end tests compress (zfpEndtoendBase.c:60-100)."""
        dtype = np.dtype(dtype)
        ptr, side, total = C.c_void_p(), C.c_size_t(), C.c_size_t()
        if dtype.kind == "f":
            fn = self.L.generateSmoothRandFloats if dtype.itemsize == 4 else self.L.generateSmoothRandDoubles
            fn.argtypes = [C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
            fn(1000000, dims, C.byref(ptr), C.byref(side), C.byref(total))
        else:
            fn = self.L.generateSmoothRandInts32 if dtype.itemsize == 4 else self.L.generateSmoothRandInts64
            fn.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
            fn(4096, dims, 8 * dtype.itemsize - 2, C.byref(ptr), C.byref(side), C.byref(total))
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(total.value * dtype.itemsize,))
        out = arr.view(dtype).reshape((side.value,) * dims).copy()
        self.libc.free(ptr)
        return out

    def hash_stream(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        return self.L.hashBitstream(words.ctypes.data, words.nbytes)
