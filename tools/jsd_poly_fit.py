"""Fit the two polynomial pieces used by the JSD kernel (JSD section).

Per dimension  t = a ln(a/h) + b ln(b/h),  h=(a+b)/2,  x=(a-b)/(a+b), u=x^2, r=|x|, w=1-r:
    t = h * f(x),  f(x) = (1+x)ln(1+x) + (1-x)ln(1-x) = u * G(u),
    G(u) = sum_{n>=1} u^(n-1) / (n (2n-1))            (regime A, u <= u0)
    f    = E(w) + w ln w,  E(w) = (2-w) ln(2-w)        (regime B, u >  u0)
Chebyshev-node interpolation in float64; reports the max relative error of each
piece evaluated in float32 Horner arithmetic.
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

if __name__ == "__main__":
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
