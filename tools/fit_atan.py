#!/usr/bin/env python
"""Minimax fit of atan(t) / t as a polynomial in t^2 on [0, 1] (Lawson-weighted least squares on a dense grid) and its error
when evaluated in fp32 Horner / FMA form: the coefficients of fast_atan2f in csrc/qx_model.cuh (degree 7).  CPU only."""
import numpy as np
from numpy.polynomial import chebyshev as C
from numpy.polynomial import polynomial as P

t = np.linspace(1e-9, 1, 20001)
s, f = t * t, np.arctan(t) / t


def fit(deg: int) -> np.ndarray:
    x, w = 2 * s - 1, np.ones_like(s)
    for _ in range(200):
        c = C.chebfit(x, f, deg, w=np.sqrt(w))
        err = np.abs((C.chebval(x, c) - f) * t)
        w = w * (err / err.max() + 1e-3)
        w /= w.sum()
    acc, pw = np.zeros(deg + 1), np.array([1.0])
    for ck in C.cheb2poly(c):
        acc[:len(pw)] += ck * pw
        pw = P.polymul(pw, np.array([-1.0, 2.0]))
    return acc


if __name__ == "__main__":
    for deg in (5, 6, 7, 8):
        a = fit(deg)
        tt = np.linspace(0, 1, 400001).astype(np.float32)
        ss = (tt * tt).astype(np.float32)
        r = np.full_like(ss, np.float32(a[-1]))
        for ck in a[-2::-1]:
            r = (r.astype(np.float64) * ss + np.float64(np.float32(ck))).astype(np.float32)  # one rounding per FMA
        val = (r.astype(np.float64) * tt).astype(np.float32)
        e = np.abs(val.astype(np.float64) - np.arctan(tt.astype(np.float64)))
        print(deg, "max |err| in fp32:", e.max(), "coefficients (t^0 .. ):", [float(np.float32(c)) for c in a])
