#!/usr/bin/env python3
"""Generate zfp_b200/csrc/coder_luts.h: tables for the "small universe" steps of the plane-lockstep coders.

While only the first eight coefficients (sequency order) of a block of 64 have bits in a plane, the plane's
whole code string is a function of (n, byte): n = coefficients already significant (0..8), byte = the plane's
bits of coefficients 0..7.  The strings follow the reference's embedded coder (src/template/encode.c:91-130,
decode.c:79-130, restated in oracle/zfp_oracle.c): the n verbatim bits, then the group-tested rest - a test
bit "any one-bit left?", and after a '1' the remaining bits up to and including the next one-bit.

  kEncLut8[n * 256 + byte]  (n = 0..9; row 9 is the idle row: empty string, n stays 9)
      bits  0..16  the string, first bit in bit 0          (at most n + 2 (8 - n) + 1 <= 17 bits)
      bits 17..21  its length
      bits 22..25  n after the plane
  kEncLut4[n * 16 + nibble] (n = 0..5; blocks of FOUR values, every plane; row 5 idle)
      bits 0..7 the string, bits 8..11 its length (<= 8), bits 12..14 n after the plane
  kDecLut8[n * 512 + t]     (n = 0..8; t = the nine stream bits that follow the n verbatim bits)
      bits  0..3   stream bits the group-tested part consumed (1..9)
      bits  4..11  the one-bits it deposits, at their coefficient positions
      bits 12..15  n after the plane
      bit  31      escape: the part does not end within nine bits, or reaches beyond coefficient 7
  kDecLut8h: the same in 16 bits (consumed | deposited << 4 | n after << 12; 0 = escape)
  kDecLut4[5 * (2^a - 1) + (n << a) + t]  blocks of FOUR values, a = min(bits of budget left, 7), t = the next a bits:
      consumed | deposited << 4 | n after << 8 (exact for every plane, budget end included)
"""
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zfp_b200", "csrc", "coder_luts.h")
SIZE = 64  # coefficients per block


def plane_string(n, word, size):
    """(bits, n after the plane) of a plane whose coefficient bits are `word`, n already significant,
    in a block of `size` coefficients (encode.c:108-124)"""
    bits = [(word >> i) & 1 for i in range(n)]
    x = word >> n
    pos = n
    while pos < size:
        bits.append(1 if x else 0)
        if not x:
            break
        while pos < size - 1:
            b = x & 1
            bits.append(b)
            if b:
                break
            x >>= 1
            pos += 1
        x >>= 1
        pos += 1
    return bits, pos


def enc_entry(n, byte):
    bits, pos = plane_string(n, byte, SIZE)
    assert pos <= 8 and len(bits) <= 17
    s = sum(b << i for i, b in enumerate(bits))
    return s | (len(bits) << 17) | (pos << 22)


def enc4_entry(n, nibble):
    """blocks of 4 values (1-D): the table covers every plane.  bits 0..7 string, 8..11 length, 12..14 n after"""
    bits, pos = plane_string(n, nibble, 4)
    assert pos <= 4 and len(bits) <= 8
    s = sum(b << i for i, b in enumerate(bits))
    return s | (len(bits) << 8) | (pos << 12)


def dec4_table():
    """blocks of 4 values, the group-tested part of a plane under a bit budget (decode.c:96-117): entry
    5 * (2^a - 1) + (n << a) + t for a = min(budget, 7) available bits t (first bit in bit 0), n = 0..4:
    bits consumed | deposited one-bits << 4 | n after << 8.  Seven bits always suffice for four coefficients."""
    tab = []
    for a in range(8):
        for n in range(5):
            for t in range(1 << a):
                bits, used, x, pos = a, 0, 0, n
                def read():
                    nonlocal used
                    b = (t >> used) & 1
                    used += 1
                    return b
                while bits and pos < 4:
                    bits -= 1
                    if read():
                        while bits and pos < 3:
                            bits -= 1
                            if read():
                                break
                            pos += 1
                        x |= 1 << pos
                        pos += 1
                    else:
                        break
                assert used <= 7 and (a < 7 or bits >= 0)
                tab.append(used | (x << 4) | (pos << 8))
    assert len(tab) == 5 * 255
    return tab


def dec_entry(n, t):
    avail = 9
    used = 0
    x = 0
    pos = n

    def read():
        nonlocal used
        if used >= avail:
            raise OverflowError
        b = (t >> used) & 1
        used += 1
        return b

    try:
        while pos < SIZE:
            if not read():
                break
            while pos < SIZE - 1:
                if read():
                    break
                pos += 1
            if pos > 7:
                raise OverflowError
            x |= 1 << pos
            pos += 1
    except OverflowError:
        return 1 << 31
    return used | (x << 4) | (pos << 12)


def main():
    enc = [enc_entry(n, b) for n in range(9) for b in range(256)] + [9 << 22] * 256
    dec = [dec_entry(n, t) for n in range(9) for t in range(512)]
    # every encoder string must decode to itself
    for n in range(9):
        for b in range(256):
            e = enc[n * 256 + b]
            s, ln, n2 = e & 0x1FFFF, (e >> 17) & 31, (e >> 22) & 15
            t = (s >> n) & 0x1FF
            d = dec[n * 512 + t]
            if d >> 31:
                assert ln - n > 9
                continue
            assert (d & 15) == ln - n and ((d >> 4) & 0xFF) == (b >> n << n) and ((d >> 12) & 15) == n2, (n, b)
    with open(OUT, "w") as f:
        f.write("/* GENERATED by tools/gen_coder_luts.py - do not edit.\n"
                " * Code strings of a bit plane whose one-bits all lie in coefficients 0..7 of a 64-value block\n"
                " * (layout of the entries: see the generator). */\n#pragma once\n#include <cstdint>\n\nnamespace zb {\n\n")
        enc4 = [enc4_entry(n, x) for n in range(5) for x in range(16)] + [5 << 12] * 16   # row 5: idle (empty string)
        for name, tab in (("kEncLut8", enc), ("kDecLut8", dec), ("kEncLut4", enc4)):
            f.write("static __device__ __align__(16) const uint32_t %s[%d] = {\n" % (name, len(tab)))
            for i in range(0, len(tab), 8):
                f.write("  " + ", ".join("0x%08xu" % v for v in tab[i:i + 8]) + ",\n")
            f.write("};\n\n")
        # the decoder's table in 16 bits: consumed (0 = escape) | deposited bits << 4 | n after << 12
        dech = [0 if d >> 31 else (d & 15) | (((d >> 4) & 0xFF) << 4) | (((d >> 12) & 15) << 12) for d in dec]
        assert all(0 <= v < 65536 for v in dech) and all((v & 15) != 0 or d >> 31 for v, d in zip(dech, dec))
        dec4 = dec4_table() + [0] * 5                                       # (padded to whole 16-byte groups)
        for name, tab in (("kDecLut8h", dech), ("kDecLut4", dec4)):
            f.write("static __device__ __align__(16) const uint16_t %s[%d] = {\n" % (name, len(tab)))
            for i in range(0, len(tab), 12):
                f.write("  " + ", ".join("0x%04x" % v for v in tab[i:i + 12]) + ",\n")
            f.write("};\n\n")
        f.write("constexpr int kEncLut8Words = %d;\nconstexpr int kDecLut8Words = %d;\nconstexpr int kEncLut4Words = %d;\nconstexpr int kDecLut8hBytes = %d;\nconstexpr int kDecLut4Bytes = %d;\n\n}  // namespace zb\n" % (len(enc), len(dec), len(enc4), 2 * len(dec), 2 * len(dec4)))
    print("wrote", OUT, len(enc), len(dec))


if __name__ == "__main__":
    main()
