"""Fit of the GELU used by the fused conv prologue:  GELU(y) = max(y, 0) - |y| * 2^P(|y|),  P = log2 Phi(-|y|) as a polynomial.
Weighted minimax (iteratively re-weighted least squares, weight = |y| Phi(-|y|) ln 2 = d GELU / d P), then the fp32 Horner
evaluation is replayed in numpy (fp32 FMA emulated through float64) against the float64 erf GELU.
usage: python tools/fit_gelu_poly.py [degree]"""
import sys
import numpy as np
from scipy.special import log_ndtr, ndtr, erf

LN2 = np.log(2.0)


def fit(deg, xmax=7.0, iters=400, n=60001):
    xs = np.linspace(0, xmax, n)
    f = log_ndtr(-xs) / LN2
    w0 = xs * ndtr(-xs) * LN2 + 1e-10
    wts = w0.copy()
    V = np.vander(xs / xmax, deg + 1, increasing=True)
    best = (1e9, None)
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(V * wts[:, None], f * wts, rcond=None)
        e = np.abs(xs * (2.0 ** (V @ c) - ndtr(-xs)))
        if e.max() < best[0]:
            best = (e.max(), c.copy())
        wts = wts * (1 + 2.0 * e / e.max()) ** 0.5
    return best[0], best[1] / (xmax ** np.arange(deg + 1))


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def gelu_fp32(y, coef):
    y = y.astype(np.float32)
    nay = -np.abs(y)
    c = [np.float32(v * (-1) ** i) for i, v in enumerate(coef)]  # polynomial in -|y|
    p = np.full_like(y, c[-1])
    for ck in c[-2::-1]:
        p = fma32(p, nay, np.full_like(y, ck))
    e = np.exp2(p.astype(np.float64)).astype(np.float32)  # ex2.approx: 2 ulp, flushes denormals
    e = np.where(e < np.float32(1.18e-38), np.float32(0), e)
    return fma32(nay, e, np.maximum(y, np.float32(0)))


if __name__ == "__main__":
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    err, coef = fit(deg)
    print("degree", deg, "float64 max |error|", err)
    print("coefficients of P(|y|), low to high:", ", ".join("%.9ef" % np.float32(v) for v in coef))
    y = np.concatenate([np.linspace(-12, 12, 2000001), np.array([-1e4, -100.0, -30.0, 30.0, 100.0, 1e4, 0.0, -0.0])])
    ref = 0.5 * y * (1 + erf(y / np.sqrt(2)))
    got = gelu_fp32(y, coef).astype(np.float64)
    d = np.abs(got - np.float32(y).astype(np.float64) * 0 - ref)
    print("fp32 replay max |error| on [-12, 12] + outliers:", d.max(), "at y =", y[d.argmax()])
