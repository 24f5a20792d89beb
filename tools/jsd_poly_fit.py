"""Fit the two polynomial pieces used by the JSD kernel (JSD section).

Per dimension  t = a ln(a/h) + b ln(b/h),  h=(a+b)/2,  x=(a-b)/(a+b), u=x^2, r=|x|, w=1-r:
    t = h * f(x),  f(x) = (1+x)ln(1+x) + (1-x)ln(1-x) = u * G(u),
    G(u) = sum_{n>=1} u^(n-1) / (n (2n-1))            (regime A, u <= u0)
    f    = E(w) + w ln w,  E(w) = (2-w) ln(2-w)        (regime B, u >  u0)
Chebyshev-node interpolation in float64; reports the max relative error of each
piece evaluated in float32 Horner arithmetic.

``--sinh``: the reciprocal-free alternative that was considered and rejected.  With g = 1/sqrt(ab)
(a product of two per-operand values, no MUFU per term), y = d g = 2 sinh(delta/2), v = y^2 = d^2/(ab):
    d x G(u) = d y Phi(v),   Phi(v) = G(v / (4 + v)) / sqrt(4 + v),   ratio <= r  <=>  v <= r + 1/r - 2
The operation count per term is 6 + degree, like the u form's, but Phi needs one degree more than G
for the same accuracy on the same ratio range (singularity at v = -4), so the form costs 11 FP32-pipe
operations where the u form costs 10 -- it only trades the MUFU for an FMA-pipe operation.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
import sys

def G_exact(u):
    u = np.asarray(u, dtype=np.float64)
    x = np.sqrt(u)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (1 + x) * np.log1p(x) + (1 - x) * np.log1p(-x)
        g = f / u
    series = sum(u ** (n - 1) / (n * (2 * n - 1)) for n in range(1, 40))
    return np.where(u < 1e-2, series, g)

def E_exact(w):
    return (2 - w) * np.log(2 - w)

def fit(fn, lo, hi, deg):
    k = np.arange(deg + 1)
    nodes = np.cos(np.pi * (k + 0.5) / (deg + 1))
    xs = 0.5 * (hi - lo) * nodes + 0.5 * (hi + lo)
    cheb = C.chebfit(nodes, fn(xs), deg)
    # convert to monomial in the original variable
    pol_t = C.cheb2poly(cheb)                       # in t in [-1,1]
    # t = (2x - (hi+lo))/(hi-lo)
    a = 2.0 / (hi - lo); b = -(hi + lo) / (hi - lo)
    mono = np.zeros(1)
    base = np.ones(1)
    for c in pol_t:
        mono = P.polyadd(mono, c * base)
        base = P.polymul(base, np.array([b, a]))
    return mono

def horner32(coef, x):
    x = x.astype(np.float32)
    acc = np.full_like(x, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        acc = (acc * x + np.float32(c)).astype(np.float32)
    return acc

def Phi_exact(v):
    v = np.asarray(v, dtype=np.float64)
    return G_exact(v / (4 + v)) / np.sqrt(4 + v)


def sinh_form():
    for ratio, degs in ((3.0, (3, 4, 5)), (5.8, (5, 6, 7))):
        hi = ratio + 1 / ratio - 2
        for deg in degs:
            co = fit(Phi_exact, 0.0, hi * 1.001, deg)
            vs = np.linspace(0, hi, 400001)
            e32 = np.abs(horner32(co, vs).astype(np.float64) / Phi_exact(vs) - 1).max()
            e64 = np.abs(P.polyval(vs, co) / Phi_exact(vs) - 1).max()
            print("Phi ratio<=%.1f (v<=%.4f) deg=%d  relerr f32=%.2e f64=%.2e" % (ratio, hi, deg, e32, e64))
    for u0, degs in ((0.25, (3, 4)), (0.4983, (5, 6))):
        for deg in degs:
            co = fit(G_exact, 0.0, u0, deg)
            us = np.linspace(0, u0, 200001)
            print("G   u<=%.4f deg=%d  relerr f32=%.2e f64=%.2e" % (
                u0, deg, np.abs(horner32(co, us).astype(np.float64) / G_exact(us) - 1).max(),
                np.abs(P.polyval(us, co) / G_exact(us) - 1).max()))


if __name__ == "__main__":
    if "--sinh" in sys.argv:
        sinh_form()
        sys.exit(0)
    u0 = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
    for deg in range(4, 12):
        co = fit(G_exact, 0.0, u0, deg)
        us = np.linspace(0, u0, 200001)
        err = np.abs(horner32(co, us).astype(np.float64) / G_exact(us) - 1).max()
        err64 = np.abs(P.polyval(us, co) / G_exact(us) - 1).max()
        print("G  u0=%.3f deg=%d  relerr f32=%.2e f64=%.2e" % (u0, deg, err, err64))
    w0 = 1 - np.sqrt(u0)
    for deg in range(3, 10):
        co = fit(E_exact, 0.0, w0, deg)
        ws = np.linspace(0, w0, 200001)
        ex = E_exact(ws)
        err = np.abs(horner32(co, ws).astype(np.float64) - ex).max()
        print("E  w0=%.3f deg=%d  abserr f32=%.2e (f >= %.3f)" % (w0, deg, err, u0 * G_exact(np.array([u0]))[0]))
