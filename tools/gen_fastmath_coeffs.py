"""Generates the polynomial coefficients of csrc/fastmath64.cuh with mpmath (80 digits): Chebyshev-node interpolation
(near-minimax) of the kernel functions below, followed by an accuracy sweep of the double-precision evaluation against
mpmath.  Output: C initialisers (hex-exact via %a-style repr) printed to stdout.

    sin(r) = r + r*z*S(z)            z = r^2, |r| <= pi/4        S: degree 5
    cos(r) = 1 - z/2 + z^2*C(z)                                   C: degree 5
    exp(r) = 1 + r + r^2*E(r)        |r| <= ln2/2                 E: degree 9
    log(m) = 2s + s*z*L(z)           s = (m-1)/(m+1), z = s^2, m in [sqrt(1/2), sqrt(2))   L: degree 6
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 80


def cheb_fit(f, a, b, deg):
    """coefficients (ascending powers) of the degree-`deg` polynomial interpolating f at Chebyshev nodes of [a,b]."""
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]


def S(z):
    r = mp.sqrt(z)
    return (mp.sin(r) / r - 1) / z


def C(z):
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / (z * z)


def E(r):
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


def L(z):
    s = mp.sqrt(z)
    return (mp.log((1 + s) / (1 - s)) - 2 * s) / (s * z)


def main():
    eps = mp.mpf(10) ** -30
    qp = (mp.pi / 4) ** 2 * mp.mpf("1.02")
    cs = cheb_fit(S, eps, qp, 5)
    cc = cheb_fit(C, eps, qp, 5)
    h = mp.log(2) / 2 * mp.mpf("1.01")
    ce = cheb_fit(E, -h, h, 9)
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    cl = cheb_fit(L, eps, (smax * mp.mpf("1.01")) ** 2, 6)

    def emit(name, c):
        vals = [float(x) for x in c]
        print(f"// {name}")
        print("    " + ", ".join(v.hex() for v in vals) + ",")
        return vals

    vs, vc, ve, vl = emit("S", cs), emit("C", cc), emit("E", ce), emit("L", cl)

    # accuracy sweep of the double evaluation (no FMA emulation: pessimistic) against mpmath
    rng = np.random.default_rng(0)

    def horner(c, x):
        p = np.full_like(x, c[-1])
        for k in c[-2::-1]:
            p = p * x + k
        return p

    def ulp_err(got, ref):
        ref64 = np.array([float(r) for r in ref])
        ulp = np.spacing(np.abs(ref64))
        return np.max(np.abs(np.array([float(mp.mpf(float(g)) - r) for g, r in zip(got, ref)])) / ulp)

    r = rng.uniform(-np.pi / 4, np.pi / 4, 4000)
    z = r * r
    sin_ = r + r * z * horner(vs, z)
    cos_ = 1.0 - 0.5 * z + z * z * horner(vc, z)
    print("// sin max ulp", ulp_err(sin_, [mp.sin(mp.mpf(float(x))) for x in r]),
          "cos max ulp", ulp_err(cos_, [mp.cos(mp.mpf(float(x))) for x in r]))
    r = rng.uniform(-np.log(2) / 2, np.log(2) / 2, 4000)
    exp_ = 1.0 + r + r * r * horner(ve, r)
    print("// exp kernel max ulp", ulp_err(exp_, [mp.exp(mp.mpf(float(x))) for x in r]))
    m = rng.uniform(np.sqrt(0.5), np.sqrt(2), 4000)
    f = m - 1.0
    s = f / (2.0 + f)
    z = s * s
    log_ = 2 * s + s * z * horner(vl, z)
    refs = [mp.log(mp.mpf(float(x))) for x in m]
    print("// log kernel (naive 2s + s z L) max ulp", ulp_err(log_, refs))
    # fdlibm-style combination: f - (hfsq - s*(hfsq + R)),  R = z*L(z) ... here  log = f - hfsq + s*(hfsq + z*L)
    hfsq = 0.5 * f * f
    log2_ = f - (hfsq - s * (hfsq + z * horner(vl, z)))
    print("// log kernel (f - (hfsq - s*(hfsq+R))) max ulp", ulp_err(log2_, refs))


if __name__ == "__main__":
    main()
