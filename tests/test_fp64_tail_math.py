"""The arithmetic claim behind the FP64-pipe tail of the fp64 decode kernels (zfp_b200/csrc/codec.cuh,
inv_lift_f64 / lift_gain), checked with exact Python integers and with numpy doubles:

for coefficients that are multiples of 2^L (L >= 10 + 2 dims) whose gain-weighted magnitude sum stays
below 2^63, the reference's inverse lifting (src/template/decode.c:13-45: wrapping int64 adds,
arithmetic shifts) never drops a bit in a shift, never leaves the int64 range where it matters, and
equals the same formulas evaluated in double precision, bit for bit.
"""
import numpy as np
import pytest

MASK = (1 << 64) - 1


def wrap(v):
    v &= MASK
    return v - (1 << 64) if v >> 63 else v


def inv_lift_int64(p, i, s):
    """decode.c inv_lift on p[i], p[i+s], p[i+2s], p[i+3s] with two's complement wrap-around."""
    x, y, z, w = p[i], p[i + s], p[i + 2 * s], p[i + 3 * s]
    y = wrap(y + (w >> 1)); w = wrap(w - (y >> 1))
    y = wrap(y + w); w = wrap(w << 1); w = wrap(w - y)
    z = wrap(z + x); x = wrap(x << 1); x = wrap(x - z)
    y = wrap(y + z); z = wrap(z << 1); z = wrap(z - y)
    w = wrap(w + x); x = wrap(x << 1); x = wrap(x - w)
    p[i], p[i + s], p[i + 2 * s], p[i + 3 * s] = x, y, z, w


def inv_lift_f64(p, i, s):
    x, y, z, w = p[i], p[i + s], p[i + 2 * s], p[i + 3 * s]
    y = w * 0.5 + y; w = y * -0.5 + w    # (products by powers of two are exact: same as one fma)
    y = y + w; w = w * 2.0 - y
    z = z + x; x = x * 2.0 - z
    y = y + z; z = z * 2.0 - y
    w = w + x; x = x * 2.0 - w
    p[i], p[i + s], p[i + 2 * s], p[i + 3 * s] = x, y, z, w


def xform_inv(p, dims, lift):
    n = 4 ** dims
    for axis in reversed(range(dims)):
        s = 4 ** axis
        for i in range(n):
            if (i >> (2 * axis)) & 3 == 0:
                lift(p, i, s)


def gain(i, dims):
    g = 1.0
    for d in range(dims):
        g *= (1.0, 1.5, 1.0, 1.25)[(i >> (2 * d)) & 3]
    return g


@pytest.mark.parametrize("dims", [1, 2, 3])
def test_fp64_inverse_lift_equals_wrapping_int64_when_the_bound_holds(dims):
    rng = np.random.default_rng(17 + dims)
    n = 4 ** dims
    lmin = 10 + 2 * dims
    taken = 0
    for trial in range(600):
        L = int(rng.integers(lmin, 60))
        # a dominant DC term plus decaying detail, scaled to sit around the bound
        mag = 2.0 ** rng.uniform(40, 63.2)
        decay = rng.uniform(0.0, 1.0)
        c = []
        for i in range(n):
            m = mag * (decay ** bin(i).count("1")) * rng.uniform(-1, 1)
            v = int(m) >> L << L
            c.append(max(-(1 << 63), min((1 << 63) - (1 << L), v)))
        S = sum(abs(v) * gain(i, dims) for i, v in enumerate(c))
        if not S < float.fromhex("0x1.fffffffp+62"):
            continue
        taken += 1
        a = list(c)
        xform_inv(a, dims, inv_lift_int64)
        d = [float(v) for v in c]
        assert all(int(f) == v for f, v in zip(d, c))
        xform_inv(d, dims, inv_lift_f64)
        assert [int(f) for f in d] == a, (dims, L, trial)
        assert all(np.float64(f) == np.float64(v) for f, v in zip(d, a))
    assert taken > 100


def test_the_bound_is_needed():
    """Above the bound the two arithmetics do part ways (so the kernels must test it)."""
    c = [0] * 64
    c[0] = (1 << 62) + (1 << 61)
    c[1] = (1 << 62)
    a = list(c)
    xform_inv(a, 3, inv_lift_int64)
    d = [float(v) for v in c]
    xform_inv(d, 3, inv_lift_f64)
    assert [int(f) for f in d] != a


def test_lift_gain_is_the_largest_column_entry_of_what_the_lift_shifts_or_returns():
    """gain() (= lift_gain in codec.cuh): per axis 1, 3/2, 1, 5/4 for the inputs x, y, z, w - derived here from the
    lift itself: every value it shifts ("y + (w >> 1)") or returns, as a linear form of the four inputs."""
    from fractions import Fraction as F
    forms = []
    x, y, z, w = ([F(int(i == j)) for j in range(4)] for i in range(4))
    add = lambda a, b: [p + q for p, q in zip(a, b)]
    sub = lambda a, b: [p - q for p, q in zip(a, b)]
    half = lambda a: [p / 2 for p in a]
    dbl = lambda a: [2 * p for p in a]
    y = add(y, half(w)); forms.append(y)             # y += w >> 1     (shifted next)
    w = sub(w, half(y))                              # w -= y >> 1
    y = add(y, w); w = sub(dbl(w), y)
    z = add(z, x); x = sub(dbl(x), z)
    y = add(y, z); z = sub(dbl(z), y)
    w = add(w, x); x = sub(dbl(x), w)
    forms += [x, y, z, w]                            # returned: shifted by the next pass or converted at the end
    col = [max(abs(f[j]) for f in forms) for j in range(4)]
    assert col == [F(1), F(3, 2), F(1), F(5, 4)]
    assert [gain(j, 1) for j in range(4)] == [float(c) for c in col]
    assert gain(1 + 4 * 3 + 16 * 1, 3) == 1.5 * 1.25 * 1.5
