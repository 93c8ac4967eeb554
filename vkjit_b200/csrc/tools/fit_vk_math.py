#!/usr/bin/env python
"""Derives the polynomial coefficients of vk_math.h (provenance; not needed at build time).

Each kernel is fitted in float64 by iteratively re-weighted least squares on Chebyshev nodes (converges to the
minimax fit of the RELATIVE error), every coefficient is then rounded to f32 one at a time from the lowest order up,
refitting the remaining ones, so that the f32 coefficient set as a whole is near-optimal.  Prints C initialisers.
"""
import numpy as np


def fit(fun, lo, hi, powers, fixed=None, weight=None, iters=60, npts=4001):
    """minimise max |w(x) * (fixed(x) + sum c_k x^k - fun(x))| over [lo, hi]; returns f64 coefficients"""
    k = np.arange(npts)
    x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / npts)
    x = x[np.abs(x) > 1e-300]
    y = fun(x) - (fixed(x) if fixed else 0.0)
    w = weight(x) if weight else np.ones_like(x)
    A = np.stack([x ** p for p in powers], axis=1)
    rw = np.ones_like(x)
    c = None
    for _ in range(iters):
        W = (w * rw)[:, None]
        c, *_ = np.linalg.lstsq(A * W, y * w * rw, rcond=None)
        err = np.abs(w * (A @ c - y))
        rw = rw * (0.2 + err / err.max()) ** 0.5   # Lawson-style reweighting
        rw /= rw.max()
    err = np.abs(w * (A @ c - y))
    return c, err.max()


def fit_f32(fun, lo, hi, powers, fixed=None, weight=None):
    """round coefficients to f32 one by one (lowest power first), refitting the rest"""
    done = {}
    remaining = list(powers)
    while remaining:
        def fx(x, done=dict(done)):
            base = fixed(x) if fixed else 0.0
            return base + sum(c * x ** p for p, c in done.items())
        c, e = fit(fun, lo, hi, remaining, fixed=fx, weight=weight)
        p0 = remaining.pop(0)
        done[p0] = float(np.float32(c[0]))
    def total(x):
        return (fixed(x) if fixed else 0.0) + sum(c * x ** p for p, c in done.items())
    k = np.arange(20001)
    x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / 20001)
    x = x[np.abs(x) > 1e-300]
    w = weight(x) if weight else 1.0
    return done, np.max(np.abs(w * (total(x) - fun(x))))


def show(name, coeffs):
    print(f"// {name}")
    for p, c in coeffs.items():
        print(f"//   x^{p}: {c!r}f  bits 0x{np.float32(c).view(np.uint32):08x}")


if __name__ == "__main__":
    L = 0.5 * np.log(2.0) * 1.0005
    # exp(r) = 1 + r + r^2 q(r)
    for deg in (5, 6):
        c, e = fit_f32(np.exp, -L, L, list(range(2, deg + 1)), fixed=lambda x: 1.0 + x, weight=lambda x: np.exp(-x))
        print(f"exp degree {deg}: max rel err {e:.3e} ({e / 2**-24:.4f} ulp-ish)")
        show("exp", c)
    # log1p(f) = f - f^2/2 + f^3 P(f), f in [sqrt(.5)-1, sqrt(2)-1]
    lo, hi = np.sqrt(0.5) - 1.0, np.sqrt(2.0) - 1.0
    for deg in (8, 9, 10, 11):
        c, e = fit_f32(np.log1p, lo, hi, list(range(3, deg + 1)), fixed=lambda x: x - 0.5 * x * x, weight=lambda x: 1.0 / np.abs(np.log1p(x)))
        print(f"log1p degree {deg}: max rel err {e:.3e} ({e / 2**-24:.4f} ulp-ish)")
        show("log1p", c)
    q = np.pi / 4 * 1.0005
    for deg in (7, 9):
        c, e = fit_f32(np.sin, 1e-9, q, list(range(3, deg + 1, 2)), fixed=lambda x: x, weight=lambda x: 1.0 / np.sin(x))
        print(f"sin degree {deg}: max rel err {e:.3e} ({e / 2**-24:.4f} ulp-ish)")
        show("sin", c)
    for deg in (6, 8):
        c, e = fit_f32(np.cos, 0.0, q, list(range(4, deg + 1, 2)), fixed=lambda x: 1.0 - 0.5 * x * x, weight=lambda x: 1.0 / np.cos(x))
        print(f"cos degree {deg}: max rel err {e:.3e} ({e / 2**-24:.4f} ulp-ish)")
        show("cos", c)
