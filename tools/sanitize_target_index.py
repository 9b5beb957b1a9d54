"""compute-sanitizer target for the speculative index rebuild (zfp_b200_index_rebuild): foreign variable-rate
streams in buffers that end right behind the stream, so that a walker reading past the end would be caught."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import zfp_b200 as zb
from helpers import make_field
n = 0
for dtype, shape, mode in ((np.float64, (128, 128, 132), {"accuracy": 1e-5}), (np.float32, (1100, 1024), {"precision": 14}),
                           (np.int32, (64, 64, 68), {"reversible": True})):
    a = make_field(shape, dtype, seed=3, kind="smooth")
    words, nbytes = zb.compress_numpy(a, **mode)
    words = np.ascontiguousarray(words[: (nbytes + 7) // 8])          # exactly the stream
    want, _ = zb.decompress_numpy(words, a.shape, a.dtype, index=zb.compress_numpy(a, want_index=True, **mode)[2], **mode)
    l0 = zb.launch_count()
    got, _ = zb.decompress_numpy(words, a.shape, a.dtype, index=None, **mode)
    assert got.tobytes() == want.tobytes()
    print(shape, mode, "stream Mbit %.1f" % (nbytes * 8 / 2**20), "launches", zb.launch_count() - l0, flush=True)
    n += 1
print("cases", n)
