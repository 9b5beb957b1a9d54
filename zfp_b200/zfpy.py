"""zfpy-compatible entry points on top of the B200 backend.

Same call signatures and stream format as the reference's Cython module (python/zfpy.pyx:140-235):
``compress_numpy(arr, tolerance=-1, rate=-1, precision=-1, write_header=True)`` returns ``bytes``
that the reference's ``zfpy.decompress_numpy`` / ``zfp -d`` can read (full header: magic, field
metadata, mode), and ``decompress_numpy(data)`` reads streams written by the reference.

Two additions for a GPU box:

* ``compress_tensor`` / ``decompress_tensor`` work on CUDA tensors without a host round trip;
* variable-rate streams may carry the block-offset index as a trailer AFTER the zfp stream
  (``index_trailer=True``).  Readers of the zfp format stop at the end of the stream and never see
  it; this module finds it from the end of the buffer and decodes in parallel instead of re-parsing
  the stream sequentially.  Layout (little endian), appended at the 8-byte aligned stream end:
      b"ZFPB200I" | u64 blocks | u64 stream_bytes | u16 length[blocks] | pad to 8 | u64 trailer_bytes
"""
import ctypes as C
import struct

import numpy as np

from . import api

HEADER_FULL = 7
_MAGIC = b"ZFPB200I"
_NP_TYPE = {1: np.int32, 2: np.int64, 3: np.float32, 4: np.float64}


def _mode_from_args(tolerance, rate, precision):
    given = [tolerance >= 0, rate >= 0, precision >= 0]
    if sum(given) > 1:
        raise ValueError("Only one of tolerance, rate, or precision can be specified")
    if tolerance >= 0:
        return {"accuracy": tolerance}
    if rate >= 0:
        return {"rate": rate}
    if precision >= 0:
        return {"precision": precision}
    return {"reversible": True}


def pack_index(lengths, stream_bytes):
    lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
    body = _MAGIC + struct.pack("<QQ", lengths.size, stream_bytes) + lengths.tobytes()
    body += b"\0" * (-len(body) % 8)
    return body + struct.pack("<Q", len(body) + 8)


def unpack_index(data):
    """(lengths, stream_bytes) if `data` ends with an index trailer, else None."""
    if len(data) < 40:
        return None
    (size,) = struct.unpack_from("<Q", data, len(data) - 8)
    if size < 32 or size > len(data) or size % 8:
        return None
    start = len(data) - size
    if bytes(data[start:start + 8]) != _MAGIC:
        return None
    blocks, stream_bytes = struct.unpack_from("<QQ", data, start + 8)
    if 24 + 2 * blocks > size - 8 or stream_bytes > start:
        return None
    lengths = np.frombuffer(data, dtype=np.uint16, count=blocks, offset=start + 24).copy()
    return lengths, stream_bytes


def compress_numpy(arr, tolerance=-1, rate=-1, precision=-1, write_header=True, index_trailer=False):
    """Compress a host array on the GPU; returns bytes in zfp's stream format."""
    arr = np.asarray(arr)
    if arr.dtype.name not in api.ZFP_TYPE:
        raise TypeError("Unknown dtype: %s" % arr.dtype)
    mode = _mode_from_args(tolerance, rate, precision)
    L = api.load_library()
    zt = api.ZFP_TYPE[arr.dtype.name]
    strides = tuple(s // arr.itemsize for s in arr.strides)
    f = api._make_field(L, arr.ctypes.data, zt, arr.shape, strides)
    z0 = L.zfp_stream_open(None)
    api._set_mode(L, z0, mode, zt, arr.ndim)
    cap = L.zfp_stream_maximum_size(z0, f) + 32
    L.zfp_stream_close(z0)
    buf = np.zeros(cap // 8 + 1, dtype=np.uint64)
    s = api.Stream(buf.ctypes.data, buf.nbytes, mode, zt, arr.ndim)
    try:
        if write_header and not L.zfp_write_header(s.z, f, HEADER_FULL):
            raise RuntimeError("zfp_write_header failed (array dimensions do not fit the header)")
        nbytes = L.zfp_compress(s.z, f)
        if not nbytes:
            raise RuntimeError("zfp_compress failed: %s" % api.last_error())
        out = buf.view(np.uint8)[:nbytes].tobytes()
        if index_trailer and not api.is_fixed_rate_mode(mode):
            out += pack_index(s.index_lengths(), nbytes)
        return out
    finally:
        L.zfp_field_free(f)
        s.close()


def decompress_numpy(compressed_data):
    """Decompress a byte stream WITH header (as written by compress_numpy / zfpy / `zfp -h`)."""
    L = api.load_library()
    data = compressed_data if isinstance(compressed_data, (bytes, bytearray)) else bytes(compressed_data)
    trailer = unpack_index(data)
    padded = np.zeros((len(data) + 15) // 8 + 1, dtype=np.uint64)
    padded.view(np.uint8)[: len(data)] = np.frombuffer(data, dtype=np.uint8)
    # parameters are placeholders until the header has been read
    s = api.Stream(padded.ctypes.data, padded.nbytes, {"reversible": True}, 4, 1)
    f = L.zfp_field_alloc()
    try:
        if not L.zfp_read_header(s.z, f, HEADER_FULL):
            raise ValueError("Failed to read required zfp header")
        dims = L.zfp_field_dimensionality(f)
        size = (C.c_size_t * 4)()
        L.zfp_field_size(f, size)
        shape = tuple(reversed([int(size[i]) for i in range(dims)]))
        out = np.empty(shape, dtype=_NP_TYPE[L.zfp_field_type(f)])
        L.zfp_field_set_pointer(f, out.ctypes.data)
        if trailer is not None and L.zfp_stream_compression_mode(s.z) != 2:
            s.set_index_lengths(trailer[0])
        if not L.zfp_decompress(s.z, f):
            raise RuntimeError("zfp_decompress failed: %s" % api.last_error())
        return out
    finally:
        L.zfp_field_free(f)
        s.close()


def compress_tensor(x, tolerance=-1, rate=-1, precision=-1, write_header=True):
    """Device-resident variant: CUDA tensor in, (uint8 CUDA tensor holding the stream, Compressed) out."""
    import torch
    mode = _mode_from_args(tolerance, rate, precision)
    c = api.compress(x, header=write_header, **mode)
    return c.words.view(torch.uint8)[: c.nbytes], c


def decompress_tensor(c, header=True):
    """Inverse of compress_tensor for the `Compressed` handle it returned (keeps the block index)."""
    return api.decompress(c, header=header)
