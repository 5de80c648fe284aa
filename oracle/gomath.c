/*
 * gomath.c — restatement of the Go standard library's math.Sin/Cos/Tan/Asin/Acos/Atan/Atan2.
 *
 * TEST INFRASTRUCTURE ONLY (see nbody_oracle.h).
 *
 * Why: calcElasticCollision (cmd/body/collisioncalc.go:104-160 of the reference) calls these
 * functions, and they live in a dependency that is not under /root/reference: the Go standard
 * library, package math, toolchain go1.26.3 (go.mod:3; README: developed on go1.20.1).  On
 * amd64 the gc toolchain has no assembly for them (only s390x does), so what the reference
 * executes is the portable Go source: math/sin.go, math/tan.go, math/asin.go, math/atan.go,
 * math/atan2.go — translations of the Cephes library (sin.c, tan.c, atan.c; S. Moshier) that
 * have not changed since go1.0 apart from the Payne-Hanek reduction for huge arguments added in
 * go1.12.  This file restates that published algorithm in C, evaluated like gc/amd64 does at the
 * default GOAMD64=v1: IEEE double, no FMA contraction (-ffp-contract=off), left-to-right.
 *
 * PARITY UNPINNED: there is no Go toolchain in this image, so nothing here was compared with a
 * run of the real library.  What was checked (tests/test_gomath.py): the decimal coefficients
 * against the bit patterns the Go source prints beside them; every function against glibc over
 * the argument ranges the collision path can produce (1-3 ulp apart — a mistyped coefficient
 * would show as a systematic error orders of magnitude larger); the special cases of atan2; and
 * bit for bit against a second restatement written separately in pure Python
 * (tests/golden/gomath_py.py), which rules out contraction / re-association by the C compiler.
 *
 * Not restated: trigReduce (Payne-Hanek, |x| >= 2^29).  The collision path only ever passes
 * angles in [-pi, 3pi/2]; beyond 2^29 these functions return NaN so that a use outside the
 * restated range cannot pass unnoticed.
 */
#include "gomath.h"

#include <math.h>
#include <stdint.h>

/* Go evaluates constant expressions such as Pi/2, 3*Pi/4 or 4/Pi exactly and rounds them once
 * when they meet a float64; these are those correctly rounded values (computed with 400-bit
 * arithmetic, tests/test_gomath.py re-derives them). */
static const double GO_PI = 0x1.921fb54442d18p+1;      /* Pi     */
static const double GO_PI_2 = 0x1.921fb54442d18p+0;    /* Pi/2   */
static const double GO_PI_4 = 0x1.921fb54442d18p-1;    /* Pi/4   */
static const double GO_3PI_4 = 0x1.2d97c7f3321d2p+1;   /* 3*Pi/4 */
static const double FOUR_OVER_PI = 0x1.45f306dc9c883p+0; /* 4/Pi  */

/* math/sin.go: sin and cos coefficients (Cephes sincof / coscof) */
static const double SIN_C[6] = {
    1.58962301576546568060e-10, /* 0x3de5d8fd1fd19ccd */
    -2.50507477628578072866e-8, /* 0xbe5ae5e5a9291f5d */
    2.75573136213857245213e-6,  /* 0x3ec71de3567d48a1 */
    -1.98412698295895385996e-4, /* 0xbf2a01a019bfdf03 */
    8.33333333332211858878e-3,  /* 0x3f8111111110f7d0 */
    -1.66666666666666307295e-1, /* 0xbfc5555555555548 */
};
static const double COS_C[6] = {
    -1.13585365213876817300e-11, /* 0xbda8fa49a0861a9b */
    2.08757008419747316778e-9,   /* 0x3e21ee9d7b4e3f05 */
    -2.75573141792967388112e-7,  /* 0xbe927e4f7eac4bc6 */
    2.48015872888517045348e-5,   /* 0x3efa01a019c844f5 */
    -1.38888888888730564116e-3,  /* 0xbf56c16c16c14f91 */
    4.16666666666665929218e-2,   /* 0x3fa555555555554b */
};
/* Pi/4 split into three parts (Cephes DP1..DP3) */
static const double PI4A = 7.85398125648498535156e-1;  /* 0x3fe921fb40000000 */
static const double PI4B = 3.77489470793079817668e-8;  /* 0x3e64442d00000000 */
static const double PI4C = 2.69515142907905952645e-15; /* 0x3ce8469898cc5170 */
static const double REDUCE_THRESHOLD = 536870912.0;    /* 1 << 29 */

/* math/tan.go: Cephes tan P, Q */
static const double TAN_P[3] = {
    -1.30936939181383777646e4, /* 0xc0c992d8d24f3f38 */
    1.15351664838587416140e6,  /* 0x413199eca5fc9ddd */
    -1.79565251976484877988e7, /* 0xc1711fead3299176 */
};
static const double TAN_Q[5] = {
    1.0,
    1.36812963470692954678e4,  /* 0x40cab8a5eeb36572 */
    -1.32089234440210967447e6, /* 0xc13427bc582abc96 */
    2.50083801823357915839e7,  /* 0x4177d98fc2ead8ef */
    -5.38695755929454629881e7, /* 0xc189afe03cbe5a31 */
};

/* the octant reduction shared by sin, cos and tan: j = octant (before "& 7"), returns z */
static double reduce_pi4(double x, uint64_t *jout)
{
    uint64_t j = (uint64_t)(x * FOUR_OVER_PI); /* integer part of x/(Pi/4) */
    double y = (double)j;
    if (j & 1) { /* map zeros to origin */
        j++;
        y++;
    }
    *jout = j;
    return ((x - y * PI4A) - y * PI4B) - y * PI4C; /* extended-precision modular arithmetic */
}

static double sin_poly(double z, double zz)
{
    return z + z * zz * ((((((SIN_C[0] * zz) + SIN_C[1]) * zz + SIN_C[2]) * zz + SIN_C[3]) * zz + SIN_C[4]) * zz + SIN_C[5]);
}
static double cos_poly(double zz)
{
    return 1.0 - 0.5 * zz +
           zz * zz * ((((((COS_C[0] * zz) + COS_C[1]) * zz + COS_C[2]) * zz + COS_C[3]) * zz + COS_C[4]) * zz + COS_C[5]);
}

/* math/sin.go: func sin */
double go_sin(double x)
{
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) {
        x = -x;
        sign = 1;
    }
    if (x >= REDUCE_THRESHOLD) return NAN; /* trigReduce not restated */
    uint64_t j;
    const double z = reduce_pi4(x, &j);
    j &= 7; /* octant modulo 2Pi */
    if (j > 3) { /* reflect in x axis */
        sign = !sign;
        j -= 4;
    }
    const double zz = z * z;
    double y = (j == 1 || j == 2) ? cos_poly(zz) : sin_poly(z, zz);
    return sign ? -y : y;
}

/* math/sin.go: func cos */
double go_cos(double x)
{
    if (isnan(x) || isinf(x)) return NAN;
    int sign = 0;
    x = fabs(x);
    if (x >= REDUCE_THRESHOLD) return NAN; /* trigReduce not restated */
    uint64_t j;
    const double z = reduce_pi4(x, &j);
    j &= 7;
    if (j > 3) {
        j -= 4;
        sign = !sign;
    }
    if (j > 1) sign = !sign;
    const double zz = z * z;
    double y = (j == 1 || j == 2) ? sin_poly(z, zz) : cos_poly(zz);
    return sign ? -y : y;
}

/* math/tan.go: func tan */
double go_tan(double x)
{
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) {
        x = -x;
        sign = 1;
    }
    if (x >= REDUCE_THRESHOLD) return NAN; /* trigReduce not restated */
    uint64_t j;
    const double z = reduce_pi4(x, &j);
    const double zz = z * z;
    double y;
    if (zz > 1e-14)
        y = z + z * (zz * (((TAN_P[0] * zz) + TAN_P[1]) * zz + TAN_P[2]) /
                     ((((zz + TAN_Q[1]) * zz + TAN_Q[2]) * zz + TAN_Q[3]) * zz + TAN_Q[4]));
    else
        y = z;
    if ((j & 2) == 2) y = -1 / y;
    return sign ? -y : y;
}

/* math/atan.go: xatan evaluates the rational approximation on [0, 0.66] */
static double xatan(double x)
{
    static const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01,
                        P2 = -7.500855792314704667340e+01, P3 = -1.228866684490136173410e+02,
                        P4 = -6.485021904942025371773e+01;
    static const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02,
                        Q2 = +4.328810604912902668951e+02, Q3 = +4.853903996359136964868e+02,
                        Q4 = +1.945506571482613964425e+02;
    double z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    z = x * z + x;
    return z;
}

/* math/atan.go: satan reduces a positive argument to [0, 0.66] */
static double satan(double x)
{
    static const double Morebits = 6.123233995736765886130e-17; /* pi/2 = PIO2 + Morebits */
    static const double Tan3pio8 = 2.41421356237309504880;      /* tan(3*pi/8) */
    if (x <= 0.66) return xatan(x);
    if (x > Tan3pio8) return GO_PI_2 - xatan(1 / x) + Morebits;
    return GO_PI_4 + xatan((x - 1) / (x + 1)) + 0.5 * Morebits;
}

/* math/atan.go: func atan */
double go_atan(double x)
{
    if (x == 0) return x;
    if (x > 0) return satan(x);
    return -satan(-x);
}

/* math/asin.go: func asin */
double go_asin(double x)
{
    if (x == 0) return x; /* special case */
    int sign = 0;
    if (x < 0) {
        x = -x;
        sign = 1;
    }
    if (x > 1) return NAN; /* special case (NaN compares false and falls through to a NaN result) */
    double temp = sqrt(1 - x * x);
    if (x > 0.7)
        temp = GO_PI_2 - satan(temp / x);
    else
        temp = satan(x / temp);
    return sign ? -temp : temp;
}

/* math/asin.go: func acos */
double go_acos(double x) { return GO_PI_2 - go_asin(x); }

/* math/atan2.go: func atan2 */
double go_atan2(double y, double x)
{
    const double pi = GO_PI;
    if (isnan(y) || isnan(x)) return NAN;
    if (y == 0) {
        if (x >= 0 && !signbit(x)) return copysign(0, y);
        return copysign(pi, y);
    }
    if (x == 0) return copysign(GO_PI_2, y);
    if (isinf(x)) {
        if (x > 0) return isinf(y) ? copysign(GO_PI_4, y) : copysign(0, y);
        return isinf(y) ? copysign(GO_3PI_4, y) : copysign(pi, y);
    }
    if (isinf(y)) return copysign(GO_PI_2, y);
    /* call atan and determine the quadrant */
    const double q = go_atan(y / x);
    if (x < 0) return q <= 0 ? q + pi : q - pi;
    return q;
}
