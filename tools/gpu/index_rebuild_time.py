"""Time zfp_decompress of a variable-rate stream that arrives without block lengths: the segment-parallel
speculative index rebuild (zfp_b200_index_rebuild; default from 16384 blocks / 4 Mbit on) against the sequential
walk (ZFP_B200_SERIAL_INDEX=1: one thread, index_scan_kernel) and against a decode with the index at hand."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import zfp_b200 as zb
from helpers import analytic_field
for side in (128, 256, 512):
    a = analytic_field((side, side, side), np.float64)
    for mode in ({"accuracy": 1e-6}, {"precision": 32}):
        words, nbytes = zb.compress_numpy(a, **mode)
        t0 = time.perf_counter(); got, used = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode); t1 = time.perf_counter()
        ts = float("nan")
        if side <= 256 or "accuracy" in mode:
            os.environ["ZFP_B200_SERIAL_INDEX"] = "1"
            t4 = time.perf_counter(); zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode); ts = time.perf_counter() - t4
            del os.environ["ZFP_B200_SERIAL_INDEX"]
        _, _, lengths = zb.compress_numpy(a, want_index=True, **mode)
        t2 = time.perf_counter(); got2, _ = zb.decompress_numpy(words, a.shape, a.dtype, index=lengths, **mode); t3 = time.perf_counter()
        nb = (side // 4) ** 3
        print("%d^3 %s: %d blocks, %.1f Mbit: without index %.3f s (sequential walk %.3f s), with index %.3f s; identical %s" %
              (side, mode, nb, nbytes * 8 / 2**20, t1 - t0, ts, t3 - t2, got.tobytes() == got2.tobytes()), flush=True)
