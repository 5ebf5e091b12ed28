#!/usr/bin/env python3
"""Near-minimax polynomial coefficients (Chebyshev-node interpolation in 60-digit arithmetic) for the
fp64 kernels of include/cpprob/math/dmath.hpp.  Prints the coefficient tables and the max error of the
double-rounded polynomial.  Re-run to regenerate: python tools/gen_poly.py
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_interp(f, a, b, n):
    """coefficients c0..cn (monomial basis, low order first) of the degree-n interpolant at Chebyshev nodes"""
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
    A = mp.matrix(n + 1, n + 1)
    y = mp.matrix(n + 1, 1)
    for i, x in enumerate(xs):
        for j in range(n + 1):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n + 1)]


def max_err(f, coefs, a, b, rel=True, samples=4001):
    cd = [mp.mpf(float(c)) for c in coefs]
    worst = mp.mpf(0)
    for k in range(samples):
        x = a + (b - a) * mp.mpf(k) / (samples - 1)
        p = mp.mpf(0)
        for c in reversed(cd):
            p = p * x + c
        fx = f(x)
        e = abs(p - fx)
        if rel and fx != 0:
            e /= abs(fx)
        worst = max(worst, e)
    return worst


def show(name, coefs):
    print(f"// {name}")
    for c in coefs:
        print(f"    {float(c).hex()},   // {mp.nstr(c, 20)}")


def main():
    # log: atanh(s)/s = 1 + z*P(z), z = s^2 in [0, zmax], s = (m-1)/(m+1), m in [sqrt(.5), sqrt(2)]
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    zmax = smax ** 2

    def f_log(z):
        if z == 0:
            return mp.mpf(1) / 3
        s = mp.sqrt(z)
        return (mp.atanh(s) / s - 1) / z

    for n in (6, 7, 8):
        c = cheb_interp(f_log, mp.mpf(0), zmax, n)
        # error relative to the full atanh(s)/s (=1+zP): abs err of P times z
        e = max_err(f_log, c, mp.mpf(0), zmax, rel=False) * zmax
        print(f"log  deg {n}: err contribution to atanh(s)/s = {mp.nstr(e, 5)}  (2^{mp.nstr(mp.log(e, 2), 5)})")
    show("LOG_P (deg 6 in z)", cheb_interp(f_log, mp.mpf(0), zmax, 6))

    # exp(r) = 1 + r + r^2 Q(r), r in [-ln2/2, ln2/2]
    h = mp.log(2) / 2 * mp.mpf("1.0001")

    def f_exp(r):
        if abs(r) < mp.mpf(10) ** -12:
            return mp.mpf(1) / 2 + r / 6 + r * r / 24
        return (mp.exp(r) - 1 - r) / (r * r)

    for n in (9, 10, 11):
        c = cheb_interp(f_exp, -h, h, n)
        e = max_err(f_exp, c, -h, h, rel=False) * h * h
        print(f"exp  deg {n}: abs err in exp(r) = {mp.nstr(e, 5)}  (2^{mp.nstr(mp.log(e, 2), 5)})")
    show("EXP_Q (deg 9 in r)", cheb_interp(f_exp, -h, h, 9))
    show("EXP_Q (deg 10 in r)", cheb_interp(f_exp, -h, h, 10))

    # sinpi(r) = r*(pi + z*S(z)), cospi(r) = 1 + z*C(z), z = r^2, r in [-1/4, 1/4]
    q = mp.mpf(1) / 16

    def f_sin(z):
        if z == 0:
            return -(mp.pi ** 3) / 6
        r = mp.sqrt(z)
        return (mp.sin(mp.pi * r) / r - mp.pi) / z

    def f_cos(z):
        if z == 0:
            return -(mp.pi ** 2) / 2
        r = mp.sqrt(z)
        return (mp.cos(mp.pi * r) - 1) / z

    for n in (5, 6):
        c = cheb_interp(f_sin, mp.mpf(0), q, n)
        e = max_err(f_sin, c, mp.mpf(0), q, rel=False) * q / mp.pi
        print(f"sinpi deg {n}: rel err = {mp.nstr(e, 5)}  (2^{mp.nstr(mp.log(e, 2), 5)})")
        c = cheb_interp(f_cos, mp.mpf(0), q, n)
        e = max_err(f_cos, c, mp.mpf(0), q, rel=False) * q / mp.cos(mp.pi / 4)
        print(f"cospi deg {n}: rel err = {mp.nstr(e, 5)}  (2^{mp.nstr(mp.log(e, 2), 5)})")
    show("SINPI_S (deg 5 in z)", cheb_interp(f_sin, mp.mpf(0), q, 5))
    show("COSPI_C (deg 6 in z)", cheb_interp(f_cos, mp.mpf(0), q, 6))


if __name__ == "__main__":
    main()
