"""zfp_b200 - B200 (sm_100a) execution backend for zfp's whole-array compress / decompress path.

The package holds the CUDA kernels + C-ABI library (csrc/ -> lib/libzfp_b200.so) and a thin
ctypes mirror of the reference's host interface (api.py).  See DESIGN.md and INTEGRATION.md.
"""
from .api import (Compressed, Stream, compress, compress_numpy, decompress, decompress_blocks, decompress_box, decompress_numpy,  # noqa: F401
                  last_error, launch_count, load_library, max_stream_words)
from .build import build_library  # noqa: F401
