"""CPU model of K1's seed refinement (nb_force.cu: w_from_seed, w_from_seed_uni): from a
MUFU.RSQ64H-like seed y0 = d2^(-1/2) (1 + delta) — high word only, |delta| up to 2^-20 — the seven
(six) FP64 instructions must deliver m_j d2^(-3/2) to < 3 ulp.  Emulated with exactly rounded
arithmetic (one rounding per __dmul_rn / __fma_rn, as compiled with -fmad=false) against a 200-bit
reference.  This is the accuracy budget behind the 1e-12 force tolerance (measured on the device:
3e-16 normwise)."""
import math
import struct
from fractions import Fraction

import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

mp = pytest.importorskip("mpmath")
mp.mp.prec = 200


def rn(q):
    return float(q)                      # Fraction -> nearest double


def dmul(a, b):
    return rn(Fraction(a) * Fraction(b))


def fma(a, b, c):
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def high_word_only(x):
    """What MUFU.RSQ64H leaves in the register pair: the high 32 bits, low word zero."""
    bits = struct.unpack("<Q", struct.pack("<d", x))[0] & 0xFFFFFFFF00000000
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def w_from_seed(y0, d2, mj):
    u = dmul(y0, y0)
    e = fma(-d2, u, 1.0)
    pp = fma(1.875, e, 1.5)
    c = fma(e, pp, 1.0)
    t = dmul(mj, y0)
    tu = dmul(t, u)
    return dmul(tu, c)


def w_from_seed_uni(y0, d2):
    u = dmul(y0, y0)
    e = fma(-d2, u, 1.0)
    pp = fma(1.875, e, 1.5)
    c = fma(e, pp, 1.0)
    s = dmul(y0, u)
    return dmul(s, c)


def seed(d2, delta):
    return high_word_only(float(mp.mpf(1) / mp.sqrt(mp.mpf(d2)) * (1 + mp.mpf(delta))))


case = st.tuples(st.floats(1e-20, 1e40), st.floats(-2.0 ** -20, 2.0 ** -20), st.floats(1e-10, 1e32))


@settings(max_examples=1500, deadline=None, derandomize=True)
@given(case)
def test_cubic_refinement_delivers_under_three_ulp(c):
    d2, delta, mj = c
    y0 = seed(d2, delta)
    exact = mp.mpf(mj) * mp.mpf(d2) ** mp.mpf(-1.5)
    w = w_from_seed(y0, d2, mj)
    # worst case by rounding analysis: 4 roundings of 2^-53 on the product chain, half of u's (the
    # correction c undoes the rest), c's own, the cubic's truncation: < 5.2e-16; observed <= 4e-16
    assert abs((mp.mpf(w) - exact) / exact) <= 6e-16
    wu = w_from_seed_uni(y0, d2)
    exact_u = mp.mpf(d2) ** mp.mpf(-1.5)
    assert abs((mp.mpf(wu) - exact_u) / exact_u) <= 6e-16
    # hoisting the mass costs one more rounding at most: m * w_uni vs the per-pair product
    assert abs((mp.mpf(dmul(mj, wu)) - exact) / exact) <= 7.2e-16


def test_truncation_error_of_the_cubic_is_far_below_an_ulp_at_the_seed_accuracy():
    # (1 - e)^(-3/2) - (1 + e(3/2 + 15/8 e)) = 35/16 e^3 + ...: with |e| ~ 2|delta| <= 2^-19 that is 1.5e-17
    for delta in (2.0 ** -20, -2.0 ** -20, 2.0 ** -22):
        e = 1 - (1 + mp.mpf(delta)) ** 2
        trunc = (1 - e) ** mp.mpf(-1.5) - (1 + e * (mp.mpf(1.5) + mp.mpf(1.875) * e))
        assert abs(trunc) < 2e-17
    # a seed whose low word is NOT zero changes nothing beyond that (the kernel's register-pairing trick)
    d2, mj = 123.456, 7e24
    y_hi = seed(d2, 3e-7)
    bits = struct.unpack("<Q", struct.pack("<d", y_hi))[0] | 0xDEADBEEF
    y_garbage = struct.unpack("<d", struct.pack("<Q", bits))[0]
    exact = mp.mpf(mj) * mp.mpf(d2) ** mp.mpf(-1.5)
    for y in (y_hi, y_garbage):
        assert abs((mp.mpf(w_from_seed(y, d2, mj)) - exact) / exact) <= 6e-16
