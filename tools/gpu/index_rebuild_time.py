"""Time the device-side index rebuild of a variable-rate stream that arrives without block lengths
(zfp_b200_decode with index == NULL: one thread walks the stream, index_scan_kernel)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import zfp_b200 as zb
from helpers import analytic_field
for side in (128, 256):
    a = analytic_field((side, side, side), np.float64)
    for mode in ({"accuracy": 1e-6}, {"precision": 32}):
        words, nbytes = zb.compress_numpy(a, **mode)
        t0 = time.perf_counter(); got, used = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode); t1 = time.perf_counter()
        _, _, lengths = zb.compress_numpy(a, want_index=True, **mode)
        t2 = time.perf_counter(); got2, _ = zb.decompress_numpy(words, a.shape, a.dtype, index=lengths, **mode); t3 = time.perf_counter()
        nb = (side // 4) ** 3
        print("%d^3 %s: %d blocks, without index %.3f s (%.2f us/block), with index %.3f s; identical %s" %
              (side, mode, nb, t1 - t0, (t1 - t0) / nb * 1e6, t3 - t2, got.tobytes() == got2.tobytes()), flush=True)
