"""Generates the polynomial coefficients used by include/mbd_fp32.h.

Least-squares fits on Chebyshev nodes (near-minimax) in float64; the printed
values are pasted into the header as float literals.  Run: python scripts/gen_fp32_coeffs.py
"""
import numpy as np
from numpy.polynomial import Polynomial as Poly
from numpy.polynomial import chebyshev as C


def cheb_nodes(a, b, n):
    k = np.arange(n)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (k + 0.5) / n)


def fit(f, a, b, deg, n=6000):
    """Polynomial P (power basis, in x) of degree `deg` approximating f on [a,b]."""
    x = cheb_nodes(a, b, n)
    u = (2 * x - (a + b)) / (b - a)
    c = np.linalg.lstsq(C.chebvander(u, deg), f(x), rcond=None)[0]
    P = Poly(C.cheb2poly(c))(Poly([-(a + b) / (b - a), 2 / (b - a)]))
    err = np.max(np.abs(P(x) - f(x)))
    return P.coef, err


def show(name, coef, err):
    print(f"// {name}: max abs fit error {err:.3e}")
    for i, c in enumerate(coef):
        print(f"//   c{i} = {np.float32(c)!r}")
    print("  " + ", ".join(f"{float(np.float32(c)):.9e}f" for c in coef))


if __name__ == "__main__":
    # atan(t) = t * P(t^2), t in [0,1]
    c, e = fit(lambda z: np.arctan(np.sqrt(z)) / np.sqrt(z), 1e-12, 1.0, 8)
    show("ATAN  P(z), z=t^2 in [0,1]", c, e)
    # log(1+f) = f - f^2/2 + f^3 * P(f), f in [sqrt(.5)-1, sqrt(2)-1]
    lo, hi = np.sqrt(0.5) - 1, np.sqrt(2) - 1

    def g(f):
        f = np.where(np.abs(f) < 1e-9, 1e-9, f)
        return (np.log1p(f) - f + 0.5 * f * f) / f ** 3

    c, e = fit(g, lo, hi, 7)
    show("LOG   P(f), f in [sqrt(.5)-1, sqrt(2)-1]", c, e * hi ** 3)
    # exp(r) = 1 + r + r^2 * P(r), r in [-ln2/2, ln2/2]
    h = np.log(2) / 2

    def ge(r):
        r = np.where(np.abs(r) < 1e-9, 1e-9, r)
        return (np.exp(r) - 1 - r) / r ** 2

    c, e = fit(ge, -h, h, 4)
    show("EXP   P(r), r in [-ln2/2, ln2/2]", c, e * h ** 2)
    # sin(r) = r + r^3 * P(r^2), cos(r) = 1 - r^2/2 + r^4 * P(r^2); r in [-pi/4, pi/4]
    q = (np.pi / 4) ** 2

    def gs(z):
        r = np.sqrt(z)
        return (np.sin(r) - r) / r ** 3

    def gc(z):
        r = np.sqrt(z)
        return (np.cos(r) - 1 + 0.5 * z) / z ** 2

    c, e = fit(gs, 1e-6, q, 3)
    show("SIN   P(z), z=r^2 in [0,(pi/4)^2]", c, e)
    c, e = fit(gc, 1e-4, q, 3)
    show("COS   P(z), z=r^2 in [0,(pi/4)^2]", c, e)
