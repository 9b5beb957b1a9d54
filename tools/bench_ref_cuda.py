"""Time the backend this project replaces - the UNMODIFIED reference src/cuda_zfp, compiled for sm_100
into oracle/_ref/libzfp_ref_cudaorig.so by `make -C oracle ref_cudaorig` - on the same device-resident
fields as ours, and check that its fixed-rate streams equal ours (developer/reporting tool).

The reference backend takes device pointers for the field and for the stream buffer
(reference src/cuda_zfp/cuZFP.cu: setup_device_field_* / setup_device_stream_*), synchronises at the
end of every call and launches on the legacy default stream, so each call is timed with CUDA events
recorded on that stream by this process (torch's current stream is set to the default stream).
Prints one JSON object per configuration."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import zfp_b200 as zb
from oracle.oracle import Reference, ZFP_TYPE
from test_gpu_fullsize import device_field

SO = os.path.join(ROOT, "oracle", "_ref", "libzfp_ref_cudaorig.so")
PEAK = 6540.8
NP = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64}


class RefCuda(Reference):
    """zfp_compress / zfp_decompress of the reference library under zfp_exec_cuda on device buffers."""

    def setup(self, x, rate, words):
        L = self.L
        n = tuple(reversed(x.shape)) + (0,) * (4 - x.dim())
        f, dims = self._field(x.data_ptr(), NP[x.dtype], n, None)
        z = L.zfp_stream_open(None)
        L.zfp_stream_set_rate(z, float(rate), ZFP_TYPE[np.dtype(NP[x.dtype])], dims, 0)
        bs = L.stream_open(words.data_ptr(), words.numel() * 8)
        L.zfp_stream_set_bit_stream(z, bs)
        assert L.zfp_stream_set_execution(z, 2), "reference built without CUDA?"
        return z, f, bs

    def compress(self, z, f):
        self.L.zfp_stream_rewind(z)
        return self.L.zfp_compress(z, f)

    def decompress(self, z, f):
        self.L.zfp_stream_rewind(z)
        return self.L.zfp_decompress(z, f)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ref = RefCuda(SO)
    side = int(os.environ.get("SIDE", "1024"))
    configs = [
        ("3D f64 %d^3" % side, (side, side, side), torch.float64, [4, 8, 16]),
        ("3D f32 %d^3" % side, (side, side, side), torch.float32, [4, 8, 16]),
        ("2D f32 %d^2" % (side * 16), (side * 16, side * 16), torch.float32, [8]),
        ("1D f64 2^%d" % (18 + side // 100), (1 << (18 + side // 100),), torch.float64, [8]),
    ]
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, shape, dtype, rates in configs:
        if only and only not in name:
            continue
        x = device_field(shape, dtype)
        raw = x.numel() * x.element_size()
        for rate in rates:
            ours = zb.compress(x, rate=rate)
            words = torch.zeros(ours.words.numel() + 16, dtype=torch.int64, device=x.device)
            z, f, bs = ref.setup(x, rate, words)
            nbytes = ref.compress(z, f)
            same_stream = bool(nbytes == ours.nbytes and torch.equal(words[: nbytes // 8], ours.words[: nbytes // 8].view(torch.int64)))
            y = torch.empty_like(x)
            ref.L.zfp_field_set_pointer(f, y.data_ptr())
            ref.decompress(z, f)
            mine = zb.decompress(ours)
            same_array = bool(torch.equal(y.view(torch.uint8), mine.view(torch.uint8)))
            ref.L.zfp_field_set_pointer(f, x.data_ptr())
            tc = timeit(lambda: ref.compress(z, f))
            ref.L.zfp_field_set_pointer(f, y.data_ptr())
            td = timeit(lambda: ref.decompress(z, f))
            oc = timeit(lambda: zb.compress(x, reuse=ours, rate=rate))
            od = timeit(lambda: zb.decompress(ours, out=mine))
            print(json.dumps({"config": name, "rate": rate, "stream_equal": same_stream, "array_equal": same_array,
                              "ref_cuda_compress_ms": round(tc, 3), "ref_cuda_decompress_ms": round(td, 3),
                              "ref_cuda_gbs": round(2 * raw / (tc + td) / 1e6, 1),
                              "ref_cuda_compress_hbm_frac": round((raw + nbytes) / tc / 1e6 / PEAK, 3),
                              "ref_cuda_decompress_hbm_frac": round((raw + nbytes) / td / 1e6 / PEAK, 3),
                              "ours_compress_ms": round(oc, 3), "ours_decompress_ms": round(od, 3),
                              "ours_gbs": round(2 * raw / (oc + od) / 1e6, 1)}), flush=True)
            ref.L.zfp_field_free(f)
            ref.L.zfp_stream_close(z)
            ref.L.stream_close(bs)
            del words, y, mine, ours
        del x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
