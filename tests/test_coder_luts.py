"""The generated plane-string tables (zfp_b200/csrc/coder_luts.h) against straight transcriptions of the
reference's coder loops (src/template/encode.c:91-130 and decode.c:79-130), entry by entry:
every encoder string must decode - with the reference's decoder loop - to the plane bits and the
significance count it was made from, every decoder entry must agree with that loop on the bits it
consumes, and the committed header must be what the generator writes."""
import importlib.util
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "zfp_b200", "csrc", "coder_luts.h")


def tables():
    src = open(HEADER).read()
    out = {}
    for m in re.finditer(r"const uint(?:32|16)_t (\w+)\[(\d+)\] = \{(.*?)\};", src, re.S):
        vals = [int(v.rstrip("u"), 16) for v in re.findall(r"0x[0-9a-fA-F]+u?", m.group(3))]
        assert len(vals) == int(m.group(2))
        out[m.group(1)] = vals
    return out


def ref_decode_plane(bits_in, n, size, budget):
    """decode.c:96-117: one bit plane; returns (plane word, n after, bits consumed)."""
    pos = 0

    def read():
        nonlocal pos
        b = bits_in[pos] if pos < len(bits_in) else 0
        pos += 1
        return b

    bits = budget
    m = min(n, bits)
    bits -= m
    x = 0
    for i in range(m):
        x |= read() << i
    while bits and n < size:
        bits -= 1
        if read():
            while bits and n < size - 1:
                bits -= 1
                if read():
                    break
                n += 1
            x += 1 << n
            n += 1
        else:
            break
    return x, n, budget - bits


def ref_encode_plane(x, n, size):
    """encode.c:108-124 (unbudgeted): the plane's bit string."""
    out = [(x >> i) & 1 for i in range(n)]
    x >>= n
    while n < size:
        out.append(1 if x else 0)
        if not x:
            break
        while n < size - 1:
            b = x & 1
            out.append(b)
            if b:
                break
            x >>= 1
            n += 1
        x >>= 1
        n += 1
    return out, n


def test_encoder_tables_decode_with_the_reference_loop():
    t = tables()
    for n in range(9):
        for byte in range(256):
            e = t["kEncLut8"][n * 256 + byte]
            s, ln, n2 = e & 0x1FFFF, (e >> 17) & 31, (e >> 22) & 15
            bits = [(s >> i) & 1 for i in range(ln)]
            want, n_after = ref_encode_plane(byte, n, 64)
            assert bits == want and n2 == n_after, (n, byte)
            x, n3, used = ref_decode_plane(bits, n, 64, 1000)
            assert (x, n3, used) == (byte, n2, ln), (n, byte)
    assert all(v == 9 << 22 for v in t["kEncLut8"][9 * 256:])          # idle row
    for n in range(5):
        for nib in range(16):
            e = t["kEncLut4"][n * 16 + nib]
            s, ln, n2 = e & 0xFF, (e >> 8) & 15, (e >> 12) & 7
            bits = [(s >> i) & 1 for i in range(ln)]
            want, n_after = ref_encode_plane(nib, n, 4)
            assert bits == want and n2 == n_after, (n, nib)
            x, n3, used = ref_decode_plane(bits, n, 4, 1000)
            assert (x, n3, used) == (nib, n2, ln), (n, nib)
    assert all(v == 5 << 12 for v in t["kEncLut4"][5 * 16:])


def test_decoder_table_agrees_with_the_reference_loop():
    t = tables()
    escapes = 0
    for n in range(9):
        for w in range(512):
            e = t["kDecLut8h"][n * 512 + w]
            e32 = t["kDecLut8"][n * 512 + w]
            stream = [(w >> i) & 1 for i in range(9)] + [1] * 80   # what follows the nine bits must not matter ...
            x, n2, used = ref_decode_plane([0] * n + stream, n, 64, 1000)
            used -= n
            fits = used <= 9 and x < 256 and n2 <= 8
            if e == 0:
                escapes += 1
                assert e32 >> 31 and not fits, (n, w)
                continue
            assert fits and (e & 15, (e >> 4) & 0xFF, e >> 12) == (used, x, n2), (n, w)
            x0, n0, u0 = ref_decode_plane([0] * n + [(w >> i) & 1 for i in range(9)] + [0] * 80, n, 64, 1000)
            assert (x0, n0, u0 - n) == (x, n2, used), (n, w)                # ... either way
            assert (e32 & 15, (e32 >> 4) & 0xFF, (e32 >> 12) & 15) == (used, x, n2)
    assert 0 < escapes < 9 * 512 // 2


def test_four_value_decoder_table_with_every_budget():
    t = tables()["kDecLut4"]
    for a in range(8):
        for n in range(5):
            for w in range(1 << a):
                e = t[5 * ((1 << a) - 1) + (n << a) + w]
                for tail in (0, 1):                                          # bits beyond the budget are never looked at
                    stream = [0] * n + [(w >> i) & 1 for i in range(a)] + [tail] * 20
                    budgets = (n + a,) if a < 7 else (n + 7, n + 9, n + 40)  # seven bits always suffice
                    for budget in budgets:
                        x, n2, used = ref_decode_plane(stream, n, 4, budget)
                        assert (e & 15, (e >> 4) & 15, e >> 8) == (used - n, x, n2), (a, n, w, tail, budget)


def test_committed_header_is_what_the_generator_writes(tmp_path):
    spec = importlib.util.spec_from_file_location("gen_coder_luts", os.path.join(ROOT, "tools", "gen_coder_luts.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.OUT = str(tmp_path / "coder_luts.h")
    gen.main()
    assert open(gen.OUT).read() == open(HEADER).read()
