"""Seeded random sweep of (shape, type, data kind, mode) against the oracle on the GPU: small
arrays, many parameter combinations - odd rates (blocks sharing words), tiny and huge budgets,
expert parameter sets that bind in several ways at once, partial blocks in every dimension."""
import numpy as np
import pytest

from helpers import make_field

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import zfp_b200
    zfp_b200.load_library(build_if_missing=False)
    return zfp_b200


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    dtypes = [np.float32, np.float64, np.int32, np.int64]
    kinds = ["smooth", "noise", "sparse", "tiny", "huge", "special"]
    for _ in range(count):
        dims = int(rng.integers(1, 5))
        hi = {1: 3000, 2: 70, 3: 24, 4: 10}[dims] * (3 if rng.integers(0, 8) == 0 else 1)   # now and then a larger array
        shape = tuple(int(rng.integers(1, hi)) for _ in range(dims))
        dtype = dtypes[int(rng.integers(0, 4))]
        kind = kinds[int(rng.integers(0, len(kinds)))]
        if np.dtype(dtype).kind != "f" and kind in ("tiny", "huge", "special"):
            kind = "noise"
        intprec = 8 * np.dtype(dtype).itemsize
        pick = int(rng.integers(0, 6))
        if pick == 0:
            mode = {"rate": float(rng.integers(1, 4 * intprec)) / 4.0}
        elif pick == 1:
            mode = {"rate": int(rng.integers(1, intprec + 1))}
        elif pick == 2:
            mode = {"precision": int(rng.integers(1, intprec + 1))}
        elif pick == 3 and np.dtype(dtype).kind == "f":
            mode = {"accuracy": float(2.0 ** int(rng.integers(-40, 8)))}
        elif pick == 4:
            mode = {"reversible": True}
        else:
            # (maxbits below the block header is refused by the backend: upstream's own size bound
            # does not hold there, see backend.cu check_params)
            maxbits = int(rng.integers(20, 4 ** dims * intprec + 100))
            minbits = int(rng.integers(1, maxbits + 1)) if rng.integers(0, 2) else 1
            mode = {"expert": (minbits, maxbits, int(rng.integers(1, intprec + 1)), int(rng.integers(-1074, 20)))}
        if kind == "special" and "reversible" not in mode:
            kind = "smooth"   # NaN / infinity are undefined behaviour in the lossy modes upstream
        yield shape, dtype, kind, mode, int(rng.integers(0, 1 << 30))


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_random_parameter_sweep(zb, port, seed):
    import torch
    bad = []
    for shape, dtype, kind, mode, fseed in _cases(seed, 70):
        a = make_field(shape, dtype, fseed, kind)
        x = torch.from_numpy(a).cuda()
        c = zb.compress(x, **mode)
        want = port.compress(a, **mode)
        got = c.to_numpy()
        ok = got.tobytes() == want.tobytes()
        if ok:
            back = zb.decompress(c).cpu().numpy()
            ok = back.tobytes() == port.decompress(want, a.shape, a.dtype, **mode).tobytes()
        if not ok:
            bad.append((shape, np.dtype(dtype).name, kind, mode, fseed))
    assert not bad, "%d mismatches, first: %r" % (len(bad), bad[:5])
