"""Shared helpers for the test-suite: bit-reproducible synthetic fields, golden-table keys.

Inputs are produced from integer recipes only (counter-based splitmix64 + integer prefix sums +
power-of-two scaling), so every platform generates the same bytes and committed hashes in
tests/golden/kat.json stay valid without shipping the arrays themselves.
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def splitmix64(idx, seed):
    """Counter-based 64-bit hash of idx (uint64 array), wrap-around arithmetic."""
    with np.errstate(over="ignore"):
        z = (idx.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def smooth_ints(shape, seed, step_bits=4):
    """Integer random walk along every axis: smooth, exactly reproducible, int64."""
    n = int(np.prod(shape))
    r = splitmix64(np.arange(n, dtype=np.uint64), seed)
    steps = (r & np.uint64((1 << (step_bits + 1)) - 1)).astype(np.int64) - (1 << step_bits)
    a = steps.reshape(shape)
    for ax in range(len(shape)):
        a = np.cumsum(a, axis=ax, dtype=np.int64)
    return a


def make_field(shape, dtype, seed=1, kind="smooth"):
    """kind: smooth | noise | sparse | tiny | huge | special (special only meaningful for floats)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    if kind == "smooth":
        a = smooth_ints(shape, seed)
        if dtype.kind == "f":
            return (a.astype(np.float64) * 2.0 ** -7).astype(dtype)
        if dtype.itemsize == 4:
            return (a * 4096).astype(np.int32)
        return a * (1 << 40)
    r = splitmix64(np.arange(n, dtype=np.uint64), seed).reshape(shape)
    if kind == "noise":
        if dtype.kind == "f":
            m = (r >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 - 0.5
            return m.astype(dtype)
        if dtype.itemsize == 4:
            return (r >> np.uint64(34)).astype(np.int64).astype(np.int32) - (1 << 28)
        return (r >> np.uint64(3)).astype(np.int64) - (1 << 59)
    if kind == "sparse":
        a = smooth_ints(shape, seed)
        mask = (r >> np.uint64(60)) == 0  # ~1/16 of the values survive
        a = np.where(mask, a, 0)
        # zero out whole slabs so that some blocks are entirely zero
        a.reshape(-1)[: n // 3] = 0
        return (a.astype(np.float64) * 2.0 ** -3).astype(dtype) if dtype.kind == "f" else a.astype(dtype)
    if kind in ("tiny", "huge") and dtype.kind == "f":
        a = smooth_ints(shape, seed).astype(np.float64)
        info = np.finfo(dtype)
        e = info.minexp + 3 if kind == "tiny" else info.maxexp - 24
        return np.ldexp(a, e - 12).astype(dtype)
    if kind == "special" and dtype.kind == "f":
        a = (smooth_ints(shape, seed).astype(np.float64) * 2.0 ** -7).astype(dtype)
        flat = a.reshape(-1)
        info = np.finfo(dtype)
        specials = [0.0, -0.0, info.smallest_subnormal, -info.smallest_subnormal, info.max, -info.max,
                    np.inf, -np.inf, np.nan, info.tiny, 1.0, -1.5]
        pos = (splitmix64(np.arange(len(specials) * 8, dtype=np.uint64), seed + 99) % np.uint64(n)).astype(np.int64)
        for i, p in enumerate(pos):
            flat[p] = specials[i % len(specials)]
        return a
    raise ValueError((kind, dtype))


def analytic_field(shape, dtype):
    """SURVEY section 8(d) 'S1' smooth analytic field (libm-dependent: never hashed into goldens,
    only fed identically to both sides of a parity check)."""
    axes = [np.linspace(0.0, 1.0, n) for n in shape]
    g = np.meshgrid(*axes, indexing="ij", sparse=True)
    if len(shape) == 1:
        f = np.sin(6 * np.pi * g[0]) + 0.25 * np.sin(14 * np.pi * g[0] ** 2)
    elif len(shape) == 2:
        y, x = g
        f = np.sin(2 * np.pi * (3 * x + 0.5 * y)) + 0.25 * np.sin(14 * np.pi * x * y)
    elif len(shape) == 3:
        z, y, x = g
        f = np.sin(2 * np.pi * (3 * x + 0.5 * y)) * np.cos(4 * np.pi * z) + 0.25 * np.sin(14 * np.pi * x * y * z)
    else:
        w, z, y, x = g
        f = np.sin(2 * np.pi * (3 * x + 0.5 * y)) * np.cos(4 * np.pi * z) * np.cos(2 * np.pi * w) + 0.25 * np.sin(14 * np.pi * x * y * z * w)
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return np.ascontiguousarray(np.broadcast_to(f, shape)).astype(dtype)
    scale = 2 ** 20 if dtype.itemsize == 4 else 2 ** 40
    return np.round(np.broadcast_to(f, shape) * scale).astype(dtype)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()[:32]


# ---- reference golden tables (tests/golden/ref_checksums.json) ---------------------------
MODE_ID = {"rate": 2, "precision": 3, "accuracy": 4, "reversible": 5}  # zfp_mode enum, include/zfp.h:96-103
TYPE_NAME = {np.dtype(np.int32): "Int32", np.dtype(np.int64): "Int64", np.dtype(np.float32): "Float",
             np.dtype(np.float64): "Double"}


def ref_table(dtype, dims):
    with open(os.path.join(GOLDEN, "ref_checksums.json")) as f:
        t = json.load(f)["%dd%s" % (dims, TYPE_NAME[np.dtype(dtype)])]
    return {(int(k1, 16), int(k2, 16)): int(c, 16) for k1, k2, c in t}


def ref_key(subject, mode_id, param, side, dims):
    """subject: 0 original input, 1 compressed stream, 2 decompressed array (array-level tests)."""
    key1 = ((((2 << 2) | subject) << 3 | mode_id) << 4) | param
    shift = {1: 0, 2: 24, 3: 16, 4: 12}[dims]
    key2 = 0
    for _ in range(dims):
        key2 = (key2 << shift) + (side - 1)
    return key1, key2


def ref_mode_cases(dtype):
    """(mode_name, param_index, mode kwargs) as swept by the reference's end-to-end tests
    (tests/utils/zfpCompressionParams.c:4-20)."""
    cases = []
    for p in range(3):
        cases.append(("rate", p, {"rate": 1 << (p + 3)}))
        cases.append(("precision", p, {"precision": 1 << (p + 3)}))
        if np.dtype(dtype).kind == "f":
            cases.append(("accuracy", p, {"accuracy": 2.0 ** -(1 << p)}))
    cases.append(("reversible", 0, {"reversible": True}))
    return cases
