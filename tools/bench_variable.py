"""Variable-rate timing with one reused zfp_stream (no per-call index allocation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import zfp_b200 as zb
from zfp_b200.api import Stream, _make_field, _tensor_type
from test_gpu_fullsize import device_field
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = zb.load_library()
for dtype, mode in ((torch.float64, {"accuracy": 1e-6}), (torch.float64, {"precision": 32}), (torch.int32, {"reversible": True}), (torch.float64, {"reversible": True})):
    x = device_field((side, side, side), dtype)
    words = torch.empty(zb.max_stream_words(x.shape, x.dtype, mode), dtype=torch.int64, device="cuda")
    s = Stream(words.data_ptr(), words.numel() * 8, mode, _tensor_type(x), 3)
    f = _make_field(L, x.data_ptr(), _tensor_type(x), tuple(x.shape), None)
    y = torch.empty_like(x)
    g = _make_field(L, y.data_ptr(), _tensor_type(x), tuple(x.shape), None)
    def comp():
        L.zfp_stream_rewind(s.z); return L.zfp_compress(s.z, f)
    def dec():
        L.zfp_stream_rewind(s.z); return L.zfp_decompress(s.z, g)
    nb = comp(); dec(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = td = 0
    for _ in range(3):
        e[0].record(); comp(); e[1].record(); dec(); e[2].record(); torch.cuda.synchronize()
        tc += e[0].elapsed_time(e[1]) / 3; td += e[1].elapsed_time(e[2]) / 3
    raw = x.numel() * x.element_size()
    print("%d^3 %s %s: ratio %.2f compress %.2f ms %.0f GB/s | decompress %.2f ms %.0f GB/s | roundtrip ok %s" % (
        side, str(dtype).split(".")[-1], mode, raw / nb, tc, raw / tc / 1e6, td, raw / td / 1e6,
        bool(torch.equal(x, y)) if "reversible" in mode else "-"), flush=True)
    del x, y, words
