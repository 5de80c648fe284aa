"""Second, separately written restatement of the Go standard library's math.Sin/Cos/Tan/Atan/
Asin/Acos/Atan2 (math/sin.go, tan.go, atan.go, asin.go, atan2.go — Cephes translations), in pure
Python: IEEE-754 double, one rounding per operation, no FMA, left to right — the FP model of
Go gc/amd64 at GOAMD64=v1.  It exists so that oracle/gomath.c is checked against something other
than itself (tests/test_gomath.py compares the two bit for bit): a compiler that contracted or
re-associated the C polynomials would show up here.  Not reference output (parity unpinned).
"""
import math

PI = math.pi                      # float64(Pi)
PI_2 = math.pi / 2                # exact halvings of float64(Pi) == float64(Pi/2), float64(Pi/4)
PI_4 = math.pi / 4
THREE_PI_4 = float.fromhex("0x1.2d97c7f3321d2p+1")   # float64(3*Pi/4), rounded once
M4PI = float.fromhex("0x1.45f306dc9c883p+0")         # float64(4/Pi), rounded once

SIN = (1.58962301576546568060e-10, -2.50507477628578072866e-8, 2.75573136213857245213e-6,
       -1.98412698295895385996e-4, 8.33333333332211858878e-3, -1.66666666666666307295e-1)
COS = (-1.13585365213876817300e-11, 2.08757008419747316778e-9, -2.75573141792967388112e-7,
       2.48015872888517045348e-5, -1.38888888888730564116e-3, 4.16666666666665929218e-2)
PI4A, PI4B, PI4C = 7.85398125648498535156e-1, 3.77489470793079817668e-8, 2.69515142907905952645e-15
TAN_P = (-1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7)
TAN_Q = (1.0, 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7,
         -5.38695755929454629881e7)
REDUCE_THRESHOLD = float(1 << 29)


def _octant(x):
    j = int(x * M4PI)
    y = float(j)
    if j & 1 == 1:
        j += 1
        y += 1
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    return j, z


def _sinpoly(z, zz):
    return z + z * zz * ((((((SIN[0] * zz) + SIN[1]) * zz + SIN[2]) * zz + SIN[3]) * zz + SIN[4]) * zz + SIN[5])


def _cospoly(zz):
    return 1.0 - 0.5 * zz + zz * zz * ((((((COS[0] * zz) + COS[1]) * zz + COS[2]) * zz + COS[3]) * zz + COS[4]) * zz + COS[5])


def sin(x):
    if x == 0 or math.isnan(x):
        return x
    if math.isinf(x):
        return math.nan
    sign = False
    if x < 0:
        x, sign = -x, True
    if x >= REDUCE_THRESHOLD:
        return math.nan      # trigReduce (Payne-Hanek) is not restated
    j, z = _octant(x)
    j &= 7
    if j > 3:
        sign, j = not sign, j - 4
    zz = z * z
    y = _cospoly(zz) if j in (1, 2) else _sinpoly(z, zz)
    return -y if sign else y


def cos(x):
    if math.isnan(x) or math.isinf(x):
        return math.nan
    sign = False
    x = abs(x)
    if x >= REDUCE_THRESHOLD:
        return math.nan
    j, z = _octant(x)
    j &= 7
    if j > 3:
        j, sign = j - 4, not sign
    if j > 1:
        sign = not sign
    zz = z * z
    y = _sinpoly(z, zz) if j in (1, 2) else _cospoly(zz)
    return -y if sign else y


def tan(x):
    if x == 0 or math.isnan(x):
        return x
    if math.isinf(x):
        return math.nan
    sign = False
    if x < 0:
        x, sign = -x, True
    if x >= REDUCE_THRESHOLD:
        return math.nan
    j, z = _octant(x)
    zz = z * z
    if zz > 1e-14:
        y = z + z * (zz * (((TAN_P[0] * zz) + TAN_P[1]) * zz + TAN_P[2]) /
                     ((((zz + TAN_Q[1]) * zz + TAN_Q[2]) * zz + TAN_Q[3]) * zz + TAN_Q[4]))
    else:
        y = z
    if j & 2 == 2:
        y = -1 / y if y != 0 else -math.copysign(math.inf, y)   # Go and C divide by zero quietly
    return -y if sign else y


def _xatan(x):
    P0, P1, P2, P3, P4 = (-8.750608600031904122785e-01, -1.615753718733365076637e+01, -7.500855792314704667340e+01,
                          -1.228866684490136173410e+02, -6.485021904942025371773e+01)
    Q0, Q1, Q2, Q3, Q4 = (2.485846490142306297962e+01, 1.650270098316988542046e+02, 4.328810604912902668951e+02,
                          4.853903996359136964868e+02, 1.945506571482613964425e+02)
    z = x * x
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4)
    return x * z + x


def _satan(x):
    morebits, tan3pio8 = 6.123233995736765886130e-17, 2.41421356237309504880
    if x <= 0.66:
        return _xatan(x)
    if x > tan3pio8:
        return PI_2 - _xatan(1 / x) + morebits
    return PI_4 + _xatan((x - 1) / (x + 1)) + 0.5 * morebits


def atan(x):
    if x == 0:
        return x
    return _satan(x) if x > 0 else -_satan(-x)


def asin(x):
    if x == 0:
        return x
    sign = False
    if x < 0:
        x, sign = -x, True
    if x > 1:
        return math.nan
    temp = math.sqrt(1 - x * x)
    if x > 0.7:
        temp = PI_2 - _satan(temp / x)
    else:
        temp = _satan(x / temp) if temp != 0 else _satan(math.inf)
    return -temp if sign else temp


def acos(x):
    return PI_2 - asin(x)


def atan2(y, x):
    if math.isnan(y) or math.isnan(x):
        return math.nan
    if y == 0:
        if x >= 0 and math.copysign(1, x) > 0:
            return math.copysign(0.0, y)
        return math.copysign(PI, y)
    if x == 0:
        return math.copysign(PI_2, y)
    if math.isinf(x):
        if x > 0:
            return math.copysign(PI_4, y) if math.isinf(y) else math.copysign(0.0, y)
        return math.copysign(THREE_PI_4, y) if math.isinf(y) else math.copysign(PI, y)
    if math.isinf(y):
        return math.copysign(PI_2, y)
    q = atan(y / x)
    if x < 0:
        return q + PI if q <= 0 else q - PI
    return q
