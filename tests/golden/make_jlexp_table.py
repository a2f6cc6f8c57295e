#!/usr/bin/env python
"""Regenerates the 256-entry table of merzbild.jl_b200/csrc/mb_jlexp.h (exp(x) as Julia evaluates it) from its definition:
entry j packs 2^(j/256) as a rounded-DOWN double head (its 52 mantissa bits; the exponent bits are implied, 0x3FF) and the top 12
significant bits of the remainder (bits 55..44 of the remainder's double pattern; its leading exponent byte 0x3C is implied).
Needs mpmath.  Run:  python tests/golden/make_jlexp_table.py  -> prints the C initialiser; `table()` is imported by the tests."""
import math
import struct


def _bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def table():
    import mpmath as mp

    mp.mp.prec = 200
    out = []
    for j in range(256):
        val = mp.mpf(2) ** (mp.mpf(j) / 256)
        head = float(val)
        if mp.mpf(head) > val:
            head = math.nextafter(head, -math.inf)
        tail = float(val - mp.mpf(head))
        out.append((((_bits(tail) >> 44) & 0xFFF) << 52) | (_bits(head) & ((1 << 52) - 1)))
    return out


if __name__ == "__main__":
    t = table()
    for i in range(0, 256, 4):
        print("    " + ", ".join("0x%016xull" % v for v in t[i:i + 4]) + ",")
