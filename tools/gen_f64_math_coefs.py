#!/usr/bin/env python
"""Generates the polynomial coefficients of the fp64 elementary functions in quadsim_b200/csrc/qs_core.cuh
(qs::f64m: atan2, sincos) with mpmath: Chebyshev-interpolation fits at 60 digits, rounded to double.

    python tools/gen_f64_math_coefs.py          # prints the C initialiser lists

atan(u) = u + u*z*Q(z),  z = u^2 in [0, tan^2(pi/8)]        (11 coefficients, |rel err| < 1e-17)
sin(r)  = r + r*z*S(z),  z = r^2 in [0, (pi/4)^2]           ( 6 coefficients)
cos(r)  = 1 - z/2 + z^2*C(z)                                ( 6 coefficients)
asin(u) = u + u*z*R(z),  z = u^2 in [0, 1/4]                (13 coefficients); |x| > 1/2: pi/2 - 2 asin(sqrt((1-|x|)/2))
pi/2    = P1 + P2 + P3 with P1, P2 carrying 33 significant bits (Cody-Waite: k*P1, k*P2 exact for |k| < 2^20)
"""
import mpmath as mp

mp.mp.dps = 60


def fit(f, a, b, n):
    c, err = mp.chebyfit(f, [a, b], n, error=True)
    return [float(x) for x in c][::-1], err


def q_atan(z):
    if z == 0:
        return -mp.mpf(1) / 3
    u = mp.sqrt(z)
    return (mp.atan(u) - u) / u ** 3


def s_sin(z):
    if z == 0:
        return -mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.sin(r) - r) / r ** 3


def c_cos(z):
    if z == 0:
        return mp.mpf(1) / 24
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / z ** 2


def r_asin(z):
    if z == 0:
        return mp.mpf(1) / 6
    u = mp.sqrt(z)
    return (mp.asin(u) - u) / u ** 3


def split33(x):
    """leading 33 significant bits of x as a double"""
    m, e = mp.frexp(x)
    return mp.ldexp(mp.floor(mp.ldexp(m, 33)), e - 33)


def show(name, c):
    print("#define %s {%s}" % (name, ", ".join(float(x).hex() for x in c)))


if __name__ == "__main__":
    zmax = mp.tan(mp.pi / 8) ** 2 * mp.mpf("1.0001")
    c, err = fit(q_atan, 0, zmax, 11)
    show("QS_F64M_ATAN", c)
    print("// atan: |Q err| %s" % mp.nstr(err, 3))
    zs = (mp.pi / 4) ** 2 * mp.mpf("1.001")
    c, err = fit(s_sin, 0, zs, 6)
    show("QS_F64M_SIN", c)
    print("// sin: |S err| %s" % mp.nstr(err, 3))
    c, err = fit(c_cos, 0, zs, 6)
    show("QS_F64M_COS", c)
    print("// cos: |C err| %s" % mp.nstr(err, 3))
    c, err = fit(r_asin, 0, mp.mpf("0.2501"), 13)
    show("QS_F64M_ASIN", c)
    print("// asin: |R err| %s" % mp.nstr(err, 3))
    p = mp.pi / 2
    p1 = split33(p)
    p2 = split33(p - p1)
    p3 = p - p1 - p2
    print("// pi/2 = P1 + P2 + P3:", float(p1).hex(), float(p2).hex(), float(p3).hex())
    for nm, v in (("pi/4", mp.pi / 4), ("pi/2", mp.pi / 2), ("pi", mp.pi)):
        hi = mp.mpf(float(v))
        print("// %s hi/lo:" % nm, float(hi).hex(), float(v - hi).hex())
    print("// 2/pi:", float(2 / mp.pi).hex(), " tan(pi/8):", float(mp.tan(mp.pi / 8)).hex())
