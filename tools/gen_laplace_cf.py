"""
Coefficients of the K-level Laplace continued fraction of the Faddeeva function,

    w(z) = (i/sqrt(pi)) / (z - (1/2)/(z - 1/(z - (3/2)/(z - ...)))),

collapsed into the rational form  w(z) ~= (i/sqrt(pi)) p(u) / (z q(u)),  u = z^2, with real
polynomials p, q of degree K/2 (K even).  Used by the channel loop for |z|^2 >= 64, where
K = 6 reproduces scipy.special.wofz to <= 3e-10 relative in Re w (checked below and by
tests/test_faddeeva.py).  Prints the C initialisers for rjp_device.cuh.
"""
import numpy as np
from numpy.polynomial import polynomial as P


def cf_polys(K):
    n, d = np.array([0.0, 1.0]), np.array([1.0])       # t_K = z / 1  (ascending powers of z)
    for k in range(K, 0, -1):
        n, d = P.polysub(P.polymul([0.0, 1.0], n), (k / 2.0) * d), n
    # w = (i/sqrt(pi)) d / n ;  n is odd in z, d even
    assert np.allclose(n[0::2], 0) and np.allclose(d[1::2], 0)
    return d[0::2], n[1::2]      # p(u) coefficients, q(u) coefficients (ascending in u)


def eval_re(x, y, p, q):
    ur, ui = x * x - y * y, 2 * x * y
    def horner(c):
        ar, ai = np.full_like(ur, c[-1]), np.zeros_like(ur)
        for ck in c[-2::-1]:
            ar, ai = ar * ur - ai * ui + ck, ar * ui + ai * ur
        return ar, ai
    ar, ai = horner(p)
    qr, qi = horner(q)
    br, bi = x * qr - y * qi, x * qi + y * qr
    return (ar * bi - ai * br) / (br * br + bi * bi) / np.sqrt(np.pi)


if __name__ == "__main__":
    from scipy.special import wofz
    for K in (4, 6, 8):
        p, q = cf_polys(K)
        worst = 0
        for y in (1e-4, 1e-3, 1e-2, 0.1, 1, 3, 7.9, 20, 100):
            x = np.linspace(0, 60, 6001)
            m = x * x + y * y >= 64.0
            if not m.any():
                continue
            ref = wofz(x[m] + 1j * y).real
            worst = max(worst, np.max(np.abs(eval_re(x[m], y, p, q) - ref) / ref))
        print(f"K={K}: p={p.tolist()} q={q.tolist()}  worst rel err for |z|>=8: {worst:.2e}")
