"""CPU model of K1's integer screen (nb_force.cu, screen_threshold): the fast pass never evaluates
the reference's overlap predicate; it keeps the minimum of the high word of its FUSED d2 per
(body, tile) and redoes the tile exactly only if that minimum falls below
    hi((r_i + rmax_tile)^2 * (1 + 2^-18)) + 2.
The screen must have no false negative: whenever the reference's own arithmetic — unfused
dist = sqrt(fl(fl(dx*dx + dy*dy) + dz*dz)), predicate !(dist > r_i + r_j) (cmd/body/body.go:192-225) —
sees an overlap, the pair must be screened.  Checked here with exactly rounded arithmetic
(fractions) on adversarial geometry: distances within a few ulp .. 1e-3 of the sum of the radii.
"""
import math
import struct
from fractions import Fraction

from hypothesis import given, settings
from hypothesis import strategies as st


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))     # one rounding


def hi(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0] >> 32


def screen_threshold(sr):
    t2 = (sr * sr) * (1.0 + 1.0 / 262144.0)
    return min(hi(t2) + 2, 0x7FF00000)


def fused_d2(dx, dy, dz):                                     # fast pass: DMUL + 2 DFMA
    return fma(dz, dz, fma(dy, dy, dx * dx))


def reference_overlap(dx, dy, dz, ri, rj):                    # body.go:194-201, 214-219
    dist = math.sqrt(dx * dx + dy * dy + dz * dz)
    return not (dist > ri + rj)


unit = st.tuples(st.floats(-1, 1), st.floats(-1, 1), st.floats(-1, 1)).filter(
    lambda v: 0.05 < math.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2))
case = st.tuples(
    unit,
    st.floats(1e-3, 1e6), st.floats(1e-3, 1e6),               # radii
    st.sampled_from([0.0, 1e-16, -1e-16, 3e-16, -3e-16, 1e-15, -1e-15, 1e-12, -1e-12, 1e-9, -1e-6, 1e-6, -1e-3, 1e-3]),
    st.floats(1.0, 4.0),                                      # rmax_tile / r_j  (>= 1)
    st.floats(-1e7, 1e7), st.floats(-1e7, 1e7), st.floats(-1e7, 1e7))   # where the pair sits


@settings(max_examples=3000, deadline=None, derandomize=True)
@given(case)
def test_screen_has_no_false_negative_near_the_threshold(c):
    (ux, uy, uz), ri, rj, eps, k, ox, oy, oz = c
    norm = math.sqrt(ux * ux + uy * uy + uz * uz)
    d = (ri + rj) * (1.0 + eps)
    xi, yi, zi = ox, oy, oz
    xj, yj, zj = ox + ux / norm * d, oy + uy / norm * d, oz + uz / norm * d
    dx, dy, dz = xj - xi, yj - yi, zj - zi                    # what both paths compute first (DADD)
    rmax = rj * k
    screened = hi(fused_d2(dx, dy, dz)) < screen_threshold(ri + rmax)
    if reference_overlap(dx, dy, dz, ri, rj):
        assert screened, (dx, dy, dz, ri, rj)
    # the per-pair refinement used when one radius dwarfs the others is conservative as well
    if reference_overlap(dx, dy, dz, ri, rj):
        assert hi(fused_d2(dx, dy, dz)) < screen_threshold(ri + rj)


def test_screen_is_tight_enough_to_be_rare():
    # a pair 0.1 % outside contact is not screened by its own radii: the margin is 2^-18 + 2 units of the high word
    ri = rj = 3.15
    d = (ri + rj) * 1.001
    assert not hi(fused_d2(d, 0.0, 0.0)) < screen_threshold(ri + rj)
    # degenerate inputs: coincident centres are always screened; inf / NaN radii screen every finite pair
    assert hi(fused_d2(0.0, 0.0, 0.0)) < screen_threshold(1e-300 + 1e-300)
    assert hi(fused_d2(1e100, 0.0, 0.0)) < screen_threshold(math.inf)
