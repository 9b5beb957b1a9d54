"""ctypes binding of libzfp_b200.so - the host-side mirror of the reference's interface.

Everything here goes through the C ABI declared in include/zfp_b200.h / zfp_b200_backend.h:
the same calls, in the same order, that the reference's own callers make
(utils/zfp.c:372-593, python/zfpy.pyx:140-235):

    stream_open -> zfp_stream_open -> zfp_stream_set_{rate,precision,accuracy,reversible}
    -> zfp_stream_set_execution(zfp_exec_cuda) -> zfp_field_{1,2,3,4}d -> zfp_compress / zfp_decompress

torch is used for device memory only (tensors own the buffers; we pass raw pointers).
There is no CPU fallback: if the library is missing, or CUDA is unavailable at call time,
the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

ZFP_TYPE = {"int32": 1, "int64": 2, "float32": 3, "float64": 4}
ZFP_EXEC_CUDA = 2
ZFP_HEADER_FULL = 7

_vp, _sz, _ss = C.c_void_p, C.c_size_t, C.c_ssize_t


class Desc(C.Structure):
    """zfp_b200_desc (include/zfp_b200_backend.h)"""
    _fields_ = [("type", C.c_int), ("dims", C.c_uint), ("n", _sz * 4), ("s", _ss * 4),
                ("minbits", C.c_uint), ("maxbits", C.c_uint), ("maxprec", C.c_uint), ("minexp", C.c_int)]


class CudaParams(C.Structure):
    """zfp_exec_params_cuda (include/zfp_b200_backend.h)"""
    _fields_ = [("magic", C.c_uint64), ("cuda_stream", _vp), ("device_only_sync", C.c_int), ("index", _vp)]


_PROTOTYPES = {
    # bit stream
    "stream_open": (_vp, [_vp, _sz]), "stream_close": (None, [_vp]), "stream_size": (_sz, [_vp]),
    "stream_capacity": (_sz, [_vp]), "stream_data": (_vp, [_vp]), "stream_rewind": (None, [_vp]),
    "stream_rtell": (C.c_uint64, [_vp]), "stream_wtell": (C.c_uint64, [_vp]),
    "stream_rseek": (None, [_vp, C.c_uint64]), "stream_wseek": (None, [_vp, C.c_uint64]),
    "stream_read_bits": (C.c_uint64, [_vp, _sz]), "stream_write_bits": (C.c_uint64, [_vp, C.c_uint64, _sz]),
    "stream_read_bit": (C.c_uint, [_vp]), "stream_write_bit": (C.c_uint, [_vp, C.c_uint]),
    "stream_flush": (_sz, [_vp]), "stream_align": (_sz, [_vp]), "stream_pad": (None, [_vp, C.c_uint64]),
    "stream_skip": (None, [_vp, C.c_uint64]), "stream_alignment": (_sz, []),
    # fields
    "zfp_type_size": (_sz, [C.c_int]), "zfp_field_alloc": (_vp, []), "zfp_field_free": (None, [_vp]),
    "zfp_field_1d": (_vp, [_vp, C.c_int, _sz]), "zfp_field_2d": (_vp, [_vp, C.c_int, _sz, _sz]),
    "zfp_field_3d": (_vp, [_vp, C.c_int, _sz, _sz, _sz]), "zfp_field_4d": (_vp, [_vp, C.c_int, _sz, _sz, _sz, _sz]),
    "zfp_field_pointer": (_vp, [_vp]), "zfp_field_begin": (_vp, [_vp]), "zfp_field_type": (C.c_int, [_vp]),
    "zfp_field_precision": (C.c_uint, [_vp]), "zfp_field_dimensionality": (C.c_uint, [_vp]),
    "zfp_field_size": (_sz, [_vp, C.POINTER(_sz)]), "zfp_field_size_bytes": (_sz, [_vp]),
    "zfp_field_blocks": (_sz, [_vp]), "zfp_field_stride": (C.c_int, [_vp, C.POINTER(_ss)]),
    "zfp_field_is_contiguous": (C.c_int, [_vp]), "zfp_field_metadata": (C.c_uint64, [_vp]),
    "zfp_field_set_pointer": (None, [_vp, _vp]), "zfp_field_set_type": (C.c_int, [_vp, C.c_int]),
    "zfp_field_set_size_1d": (None, [_vp, _sz]), "zfp_field_set_size_2d": (None, [_vp, _sz, _sz]),
    "zfp_field_set_size_3d": (None, [_vp, _sz, _sz, _sz]), "zfp_field_set_size_4d": (None, [_vp, _sz, _sz, _sz, _sz]),
    "zfp_field_set_stride_1d": (None, [_vp, _ss]), "zfp_field_set_stride_2d": (None, [_vp, _ss, _ss]),
    "zfp_field_set_stride_3d": (None, [_vp, _ss, _ss, _ss]), "zfp_field_set_stride_4d": (None, [_vp, _ss, _ss, _ss, _ss]),
    "zfp_field_set_metadata": (C.c_int, [_vp, C.c_uint64]),
    # compressed stream
    "zfp_stream_open": (_vp, [_vp]), "zfp_stream_close": (None, [_vp]), "zfp_stream_bit_stream": (_vp, [_vp]),
    "zfp_stream_set_bit_stream": (None, [_vp, _vp]), "zfp_stream_rewind": (None, [_vp]),
    "zfp_stream_flush": (_sz, [_vp]), "zfp_stream_align": (_sz, [_vp]),
    "zfp_stream_compression_mode": (C.c_int, [_vp]), "zfp_stream_rate": (C.c_double, [_vp, C.c_uint]),
    "zfp_stream_precision": (C.c_uint, [_vp]), "zfp_stream_accuracy": (C.c_double, [_vp]),
    "zfp_stream_mode": (C.c_uint64, [_vp]),
    "zfp_stream_params": (None, [_vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_int)]),
    "zfp_stream_compressed_size": (_sz, [_vp]), "zfp_stream_maximum_size": (_sz, [_vp, _vp]),
    "zfp_stream_set_reversible": (None, [_vp]),
    "zfp_stream_set_rate": (C.c_double, [_vp, C.c_double, C.c_int, C.c_uint, C.c_int]),
    "zfp_stream_set_precision": (C.c_uint, [_vp, C.c_uint]), "zfp_stream_set_accuracy": (C.c_double, [_vp, C.c_double]),
    "zfp_stream_set_mode": (C.c_int, [_vp, C.c_uint64]),
    "zfp_stream_set_params": (C.c_int, [_vp, C.c_uint, C.c_uint, C.c_uint, C.c_int]),
    "zfp_stream_execution": (C.c_int, [_vp]), "zfp_stream_set_execution": (C.c_int, [_vp, C.c_int]),
    "zfp_stream_omp_threads": (C.c_uint, [_vp]), "zfp_stream_omp_chunk_size": (C.c_uint, [_vp]),
    "zfp_stream_set_omp_threads": (C.c_int, [_vp, C.c_uint]), "zfp_stream_set_omp_chunk_size": (C.c_int, [_vp, C.c_uint]),
    "zfp_compress": (_sz, [_vp, _vp]), "zfp_decompress": (_sz, [_vp, _vp]),
    "zfp_write_header": (_sz, [_vp, _vp, C.c_uint]), "zfp_read_header": (_sz, [_vp, _vp, C.c_uint]),
    # backend C ABI
    "cuda_compress": (_sz, [_vp, _vp]), "cuda_decompress": (None, [_vp, _vp]),
    "zfp_b200_compress_stream": (_sz, [_vp, _vp]), "zfp_b200_decompress_stream": (_sz, [_vp, _vp]),
    "zfp_stream_cuda_params": (C.POINTER(CudaParams), [_vp]),
    "zfp_b200_encode": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, C.POINTER(C.c_uint64), _vp, _vp]),
    "zfp_b200_encode_async": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, _vp, _vp, _vp]),
    "zfp_b200_decode": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, C.POINTER(C.c_uint64), _vp, _vp]),
    "zfp_b200_decode_async": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, _vp, _vp, _vp]),
    "zfp_b200_decode_box": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, C.POINTER(_sz), C.POINTER(_sz), _vp, _vp]),
    "zfp_b200_decode_blocks": (C.c_int, [C.POINTER(Desc), _vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "zfp_b200_bitcopy": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64, _vp]),
    "zfp_b200_bitcopy_ranked": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint, _vp, _vp, _vp]),
    "zfp_b200_multi_create": (_vp, [C.c_int, C.POINTER(C.c_int)]), "zfp_b200_multi_destroy": (None, [_vp]),
    "zfp_b200_multi_devices": (C.c_int, [_vp]), "zfp_b200_multi_stream": (_vp, [_vp, C.c_int]),
    "zfp_b200_multi_compress": (C.c_int, [_vp, C.POINTER(Desc), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zfp_b200_multi_decompress": (C.c_int, [_vp, C.POINTER(Desc), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "zfp_b200_is_fixed_rate": (C.c_int, [C.POINTER(Desc)]), "zfp_b200_blocks": (_sz, [C.POINTER(Desc)]),
    "zfp_b200_capacity": (_sz, [C.POINTER(Desc), C.c_uint64]),
    "zfp_b200_index_create": (_vp, []), "zfp_b200_index_destroy": (None, [_vp]), "zfp_b200_index_bits": (C.c_uint64, [_vp]),
    "zfp_b200_index_blocks": (_sz, [_vp]), "zfp_b200_index_export": (_sz, [_vp, _vp, _sz]),
    "zfp_b200_index_rebuild": (C.c_int, [_vp, _vp, C.c_uint64, _sz, _vp, _vp]),
    "zfp_b200_index_import": (C.c_int, [_vp, _vp, _sz]),
    "zfp_b200_last_error": (C.c_char_p, []), "zfp_b200_launch_count": (C.c_uint64, []),
    "zfp_b200_release_scratch": (None, []),
}

EXPORTED_SYMBOLS = sorted(_PROTOTYPES) + ["stream_word_bits", "zfp_codec_version", "zfp_library_version", "zfp_version_string"]

_lib = None


def load_library(build_if_missing=True):
    """Load libzfp_b200.so (building it with nvcc if absent).  Raises if it cannot be had."""
    global _lib
    if _lib is None:
        path = os.environ.get("ZFP_B200_LIB", _build.LIB)  # developer override for A/B experiments
        if not os.path.exists(path):
            if not build_if_missing:
                raise FileNotFoundError(path)
            _build.build_library()
        lib = C.CDLL(path)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error():
    return load_library().zfp_b200_last_error().decode()


def launch_count():
    return int(load_library().zfp_b200_launch_count())


# ---- mode handling -------------------------------------------------------------------------------
def _set_mode(L, z, mode, zfp_type, dims):
    if mode.get("rate") is not None:
        L.zfp_stream_set_rate(z, float(mode["rate"]), zfp_type, dims, int(bool(mode.get("align", False))))
    elif mode.get("precision") is not None:
        L.zfp_stream_set_precision(z, int(mode["precision"]))
    elif mode.get("accuracy") is not None:
        L.zfp_stream_set_accuracy(z, float(mode["accuracy"]))
    elif mode.get("reversible"):
        L.zfp_stream_set_reversible(z)
    elif mode.get("expert") is not None:
        if not L.zfp_stream_set_params(z, *mode["expert"]):
            raise ValueError("invalid expert parameters %r" % (mode["expert"],))
    else:
        raise ValueError("no compression mode given: %r" % (mode,))


def mode_params(mode, dtype_name, dims):
    """(minbits, maxbits, maxprec, minexp) the library derives for a mode (no GPU needed)."""
    L = load_library()
    z = L.zfp_stream_open(None)
    _set_mode(L, z, mode, ZFP_TYPE[str(dtype_name).replace("torch.", "")], dims)
    v = [C.c_uint(), C.c_uint(), C.c_uint(), C.c_int()]
    L.zfp_stream_params(z, *[C.byref(x) for x in v])
    L.zfp_stream_close(z)
    return tuple(x.value for x in v)


def is_fixed_rate_mode(mode):
    if mode.get("rate") is not None:
        return True
    if mode.get("expert") is not None:
        return mode["expert"][0] == mode["expert"][1]
    return False


def bitcopy(dst_words, dst_bit, src_words, src_bit, nbits, cuda_stream=None):
    """Device bit-granular copy between two int64 CUDA tensors of stream words."""
    rc = load_library().zfp_b200_bitcopy(dst_words.data_ptr(), dst_bit, src_words.data_ptr(), src_bit, nbits, cuda_stream)
    if rc:
        raise RuntimeError("zfp_b200_bitcopy failed: %s" % last_error())


def _make_field(L, ptr, zfp_type, shape, strides):
    """shape/strides in array order (slowest first), as numpy / torch report them (in elements)."""
    dims = len(shape)
    if not 1 <= dims <= 4:
        raise ValueError("zfp supports 1-4 dimensional arrays")
    n = list(reversed(shape))
    f = getattr(L, "zfp_field_%dd" % dims)(ptr, zfp_type, *n)
    if strides is not None:
        getattr(L, "zfp_field_set_stride_%dd" % dims)(f, *reversed(strides))
    return f


class Stream:
    """A zfp_stream bound to a bit stream over a caller-owned buffer (host or device)."""

    def __init__(self, buffer_ptr, buffer_bytes, mode, zfp_type, dims, cuda_stream=None, async_fixed_rate=False):
        L = load_library()
        self.L = L
        self.bs = L.stream_open(buffer_ptr, buffer_bytes)
        self.z = L.zfp_stream_open(self.bs)
        _set_mode(L, self.z, mode, zfp_type, dims)
        if not L.zfp_stream_set_execution(self.z, ZFP_EXEC_CUDA):
            raise RuntimeError("zfp_stream_set_execution(zfp_exec_cuda) failed")
        if cuda_stream is not None or async_fixed_rate:
            p = L.zfp_stream_cuda_params(self.z)
            p.contents.cuda_stream = cuda_stream
            p.contents.device_only_sync = int(async_fixed_rate)

    def close(self):
        if self.z:
            self.L.zfp_stream_close(self.z)
            self.L.stream_close(self.bs)
            self.z = self.bs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def index_bits(self):
        """Total coded bits of the last variable-rate compress on this stream (no copy of the index)."""
        p = self.L.zfp_stream_cuda_params(self.z)
        ix = p.contents.index if p else None
        return int(self.L.zfp_b200_index_bits(ix)) if ix else 0

    def index_lengths(self):
        """Per-block coded lengths (numpy uint16) of the last variable-rate compress, or None."""
        p = self.L.zfp_stream_cuda_params(self.z)
        ix = p.contents.index
        if not ix:
            return None
        n = self.L.zfp_b200_index_blocks(ix)
        out = np.empty(n, dtype=np.uint16)
        if self.L.zfp_b200_index_export(ix, out.ctypes.data, n) != n:
            raise RuntimeError(last_error())
        return out

    def set_index_lengths(self, lengths):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
        p = self.L.zfp_stream_cuda_params(self.z)
        if not p.contents.index:
            p.contents.index = self.L.zfp_b200_index_create()
        if self.L.zfp_b200_index_import(p.contents.index, lengths.ctypes.data, lengths.size):
            raise RuntimeError(last_error())


# ---- device-resident API (torch tensors) ----------------------------------------------------------
def _torch():
    import torch
    return torch


def _tensor_type(t):
    torch = _torch()
    return {torch.int32: 1, torch.int64: 2, torch.float32: 3, torch.float64: 4}[t.dtype]


class Compressed:
    """A compressed stream held in a device tensor of 64-bit words."""

    def __init__(self, words, nbytes, shape, dtype, mode, stream, start_bit=0):
        self.words, self.nbytes, self.shape, self.dtype, self.mode = words, nbytes, tuple(shape), dtype, dict(mode)
        self.stream, self.start_bit = stream, start_bit

    def to_numpy(self):
        return self.words[: self.nbytes // 8].cpu().numpy().view(np.uint64)


def max_stream_words(shape, dtype, mode, start_bit=0):
    """Capacity (64-bit words) a stream buffer needs for `shape` under `mode` (zfp_stream_maximum_size)."""
    L = load_library()
    torch = _torch()
    zt = {torch.int32: 1, torch.int64: 2, torch.float32: 3, torch.float64: 4}[dtype] if not isinstance(dtype, int) else dtype
    z = L.zfp_stream_open(None)
    _set_mode(L, z, mode, zt, len(shape))
    f = _make_field(L, None, zt, shape, None)
    n = L.zfp_stream_maximum_size(z, f)
    L.zfp_field_free(f)
    L.zfp_stream_close(z)
    return n // 8 + (start_bit + 63) // 64 + 1


def compress(x, out=None, start_bit=0, header=False, cuda_stream=None, async_fixed_rate=False, reuse=None, **mode):
    """Compress a CUDA tensor (any strides) into a device-resident zfp stream.

    Mirrors zfp_compress on a zfp_field over a device pointer (include/zfp.h:585-590).
    `reuse`: a previous `Compressed` of the same shape/dtype/mode on the same device whose zfp_stream, bit
    stream and block index are recycled (what a C caller does by keeping its zfp_stream and rewinding
    it); anything else - another shape, a buffer too small, explicit `out` / `cuda_stream` /
    `async_fixed_rate` arguments, which belong to a fresh stream - is refused, not silently ignored.
    """
    torch = _torch()
    if not x.is_cuda:
        raise ValueError("compress() wants a CUDA tensor; use compress_numpy() for host arrays")
    L = load_library()
    zt = _tensor_type(x)
    if reuse is not None:
        if out is not None or cuda_stream is not None or async_fixed_rate:
            raise ValueError("compress(reuse=...) recycles the previous call's stream; out / cuda_stream / async_fixed_rate do not apply")
        need = max_stream_words(x.shape, x.dtype, mode, start_bit)
        if not (reuse.mode == dict(mode) and reuse.dtype == x.dtype and tuple(reuse.shape) == tuple(x.shape) and
                reuse.words.device == x.device and reuse.words.numel() >= need):
            raise ValueError("compress(reuse=...) needs the same shape, dtype, mode and device as the recycled stream "
                             "(buffer of %d words, %d needed)" % (reuse.words.numel(), need))
        out, s = reuse.words, reuse.stream
        L.zfp_stream_rewind(s.z)
    else:
        if out is None:
            out = torch.empty(max_stream_words(x.shape, x.dtype, mode, start_bit), dtype=torch.int64, device=x.device)
        s = Stream(out.data_ptr(), out.numel() * 8, mode, zt, x.dim(), cuda_stream, async_fixed_rate)
    f = _make_field(L, x.data_ptr(), zt, tuple(x.shape), tuple(x.stride()))
    if start_bit:
        L.stream_wseek(s.bs, start_bit)
    if header:
        if not L.zfp_write_header(s.z, f, ZFP_HEADER_FULL):
            raise RuntimeError("zfp_write_header failed")
    nbytes = L.zfp_compress(s.z, f)
    L.zfp_field_free(f)
    if not nbytes:
        raise RuntimeError("zfp_compress failed: %s" % last_error())
    return Compressed(out, nbytes, x.shape, x.dtype, mode, s, start_bit)


def decompress(c, out=None, header=False):
    """Decompress a `Compressed` into a CUDA tensor; returns the tensor."""
    torch = _torch()
    L = load_library()
    if out is None:
        out = torch.empty(c.shape, dtype=c.dtype, device=c.words.device)
    s = c.stream
    f = _make_field(L, out.data_ptr(), _tensor_type(out), tuple(out.shape), tuple(out.stride()))
    L.stream_rseek(s.bs, c.start_bit)
    if header:
        if not L.zfp_read_header(s.z, f, ZFP_HEADER_FULL):
            raise RuntimeError("zfp_read_header failed")
        L.zfp_field_set_pointer(f, out.data_ptr())
    nbytes = L.zfp_decompress(s.z, f)
    L.zfp_field_free(f)
    if not nbytes:
        raise RuntimeError("zfp_decompress failed: %s" % last_error())
    if nbytes != c.nbytes:
        raise RuntimeError("zfp_decompress consumed %d bytes, compress produced %d" % (nbytes, c.nbytes))
    return out


def decompress_blocks(c, block0, block1, out):
    """Random access: decode blocks [block0, block1) (stream order, x fastest) of a `Compressed` into
    their places in the CUDA tensor `out` (same shape / dtype as the compressed array); everything
    else in `out` is left as it is.  Variable-rate streams use the block index kept with `c`."""
    L = load_library()
    zt = _tensor_type(out)
    minbits, maxbits, maxprec, minexp = mode_params(c.mode, str(out.dtype).split(".")[-1], out.dim())
    d = Desc()
    d.type, d.dims = zt, out.dim()
    for i, (n, st) in enumerate(zip(reversed(out.shape), reversed(out.stride()))):
        d.n[i], d.s[i] = n, st
    d.minbits, d.maxbits, d.maxprec, d.minexp = minbits, maxbits, maxprec, minexp
    index = None
    if minbits != maxbits:
        p = L.zfp_stream_cuda_params(c.stream.z)
        index = p.contents.index if p else None
    rc = L.zfp_b200_decode_blocks(C.byref(d), out.data_ptr(), c.words.data_ptr(), c.start_bit, block0, block1, index, None)
    if rc:
        raise RuntimeError("zfp_b200_decode_blocks failed (%d): %s" % (rc, last_error()))
    return out


def box_block_ranges(shape, lo, hi):
    """Block ranges [b0, b1) (stream order, last dimension fastest) of the 4^d blocks that intersect the
    box lo <= index < hi of an array of this shape: one range per row of blocks along the fastest dimension."""
    import itertools
    shape = tuple(int(n) for n in shape)
    dims = len(shape)
    if len(lo) != dims or len(hi) != dims:
        raise ValueError("lo / hi need one entry per dimension")
    nb = [(n + 3) // 4 for n in shape]                      # blocks per dimension, slowest first
    b_lo = [max(0, int(l)) // 4 for l in lo]
    b_hi = [min((min(int(h), n) + 3) // 4, m) for h, n, m in zip(hi, shape, nb)]
    if any(a >= b for a, b in zip(b_lo, b_hi)):
        return []
    ranges = []
    for idx in itertools.product(*[range(a, b) for a, b in zip(b_lo[:-1], b_hi[:-1])]):
        base = 0
        for i, m in zip(idx, nb[:-1]):
            base = base * m + i
        base *= nb[-1]
        ranges.append((base + b_lo[-1], base + b_hi[-1]))
    return ranges


def decompress_box(c, lo, hi, out):
    """Random access by coordinates: decode every block that intersects the box lo <= index < hi
    (one (lo, hi) pair per array dimension, slowest first, like the tensor's shape) into `out`.
    Blocks are decoded whole, so values of `out` up to 3 positions outside the box along each
    dimension are written too.  ONE kernel launch over the list of blocks (zfp_b200_decode_box)."""
    L = load_library()
    dims = out.dim()
    if len(lo) != dims or len(hi) != dims:
        raise ValueError("lo / hi need one entry per dimension")
    minbits, maxbits, maxprec, minexp = mode_params(c.mode, str(out.dtype).split(".")[-1], dims)
    d = Desc()
    d.type, d.dims = _tensor_type(out), dims
    for i, (n, st) in enumerate(zip(reversed(out.shape), reversed(out.stride()))):
        d.n[i], d.s[i] = n, st
    d.minbits, d.maxbits, d.maxprec, d.minexp = minbits, maxbits, maxprec, minexp
    blo, bhi = (_sz * 4)(), (_sz * 4)()
    for i, (l, h) in enumerate(zip(reversed(lo), reversed(hi))):
        blo[i], bhi[i] = max(0, int(l)), max(0, int(h))
    index = None
    if minbits != maxbits:
        p = L.zfp_stream_cuda_params(c.stream.z)
        index = p.contents.index if p else None
    rc = L.zfp_b200_decode_box(C.byref(d), out.data_ptr(), c.words.data_ptr(), c.start_bit, blo, bhi, index, None)
    if rc:
        raise RuntimeError("zfp_b200_decode_box failed: %s" % last_error())
    return out


# ---- host arrays (numpy), same calls with host pointers: the backend stages through the device ----
def compress_numpy(a, start_bit=0, prefix_words=None, want_index=False, **mode):
    """zfpy.compress_numpy analogue without header: returns (uint64 words, nbytes[, block lengths])."""
    L = load_library()
    a = np.asarray(a)
    zt = ZFP_TYPE[a.dtype.name]
    strides = tuple(s // a.itemsize for s in a.strides)
    z0 = L.zfp_stream_open(None)
    _set_mode(L, z0, mode, zt, a.ndim)
    f = _make_field(L, a.ctypes.data, zt, a.shape, strides)
    cap = L.zfp_stream_maximum_size(z0, f) + 8 * ((start_bit + 63) // 64 + 2)
    L.zfp_stream_close(z0)
    words = np.zeros(cap // 8, dtype=np.uint64)
    if prefix_words is not None:
        words[: len(prefix_words)] = prefix_words
    s = Stream(words.ctypes.data, words.nbytes, mode, zt, a.ndim)
    if start_bit:
        L.stream_wseek(s.bs, start_bit)
    nbytes = L.zfp_compress(s.z, f)
    L.zfp_field_free(f)
    if not nbytes:
        raise RuntimeError("zfp_compress failed: %s" % last_error())
    index = s.index_lengths() if want_index else None
    s.close()
    out = words[: nbytes // 8]
    return (out, nbytes, index) if want_index else (out, nbytes)


def decompress_numpy(words, shape, dtype, out=None, start_bit=0, index=None, **mode):
    L = load_library()
    words = np.concatenate([np.ascontiguousarray(words, dtype=np.uint64), np.zeros(2, dtype=np.uint64)])
    if out is None:
        out = np.empty(shape, dtype=dtype)
    zt = ZFP_TYPE[out.dtype.name]
    strides = tuple(s // out.itemsize for s in out.strides)
    s = Stream(words.ctypes.data, words.nbytes, mode, zt, out.ndim)
    if index is not None:
        s.set_index_lengths(index)
    f = _make_field(L, out.ctypes.data, zt, out.shape, strides)
    L.stream_rseek(s.bs, start_bit)
    nbytes = L.zfp_decompress(s.z, f)
    L.zfp_field_free(f)
    s.close()
    if not nbytes:
        raise RuntimeError("zfp_decompress failed: %s" % last_error())
    return out, nbytes
