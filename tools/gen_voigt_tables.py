"""
Generate rajepy_b200/csrc/rjp_voigt_tables.inc: the constants of the mixed-precision
Voigt evaluation of the channel loop (rjp_integrate.cu, `voigt_fast_*`), which replaces
scipy.special.wofz at maths/rrls.py:353 for cells with a small Lorentz/Gauss ratio
y = (dnu_L / 2) / (sigma sqrt2) <= RJP_VT_Y_MAX (97 % of the cells of the BASELINE jets).

Decomposition (z = x + iy, Daw = complex Dawson function):

    w(z) = exp(-z^2) + (2i/sqrt(pi)) Daw(z)
    K(x, y) = Re w = exp(y^2 - x^2) cos(2xy) - H(x, y),
    H(x, y) = (2/sqrt(pi)) Im Daw(x + iy)
            = (2/sqrt(pi)) [ y D1(x) - y^3 D3(x)/6 + y^5 D5(x)/120 - ... ],
    D1 = 1 - 2x F(x),  D_{n+1} = -2n D_{n-1} - 2x D_n   (F = Dawson's integral).

The Gaussian term is ill-conditioned (d ln / d ln x = -2x^2) and is evaluated in fp64;
H is smooth, proportional to y, and evaluated in fp32 (the products are accumulated in
fp64).  Three y-terms give <= 3e-8 relative truncation error for y <= 0.1.

Everything is tabulated in the scaled coordinate X = kappa x, kappa = sqrt(log2 e), so
that the Gaussian is 2^(Y^2 - X^2) with Y = kappa y.

* core, X < 8 (x < 6.66): 16 intervals of width 1/2 in X; per interval Chebyshev-fitted
  polynomials in t in [-1, 1) (monomial coefficients, fp32) of
      A(x) = (2/sqrt pi) D1,  B(x) = -(2/sqrt pi) D3/6,  C(x) = (2/sqrt pi) D5/120,
      H / y = A + y^2 (B + y^2 C).
* wings, x >= 5.3: with U = 1/X^2 (u = 1/x^2 = kappa^2 U)
      K = y u [ g1(u) + (y^2 u) g3(u) + (y^2 u)^2 g5(u) ]      (Gaussian < 1e-8 K)
  g_n fitted as polynomials in U (fp32 monomial coefficients, normalised by G1(0) which is
  kept in fp64 as RJP_VT_G10); the kernel combines them
  per cell into one degree-6 polynomial P(U) = sum_k (G1_k + Y^2 G3_{k-1} + Y^4 G5_{k-2}) U^k.
* exp2: degree-8 polynomial of 2^f on [-1/2, 1/2], pre-multiplied by (1 + 2^-25) so that the
  kernel's truncating double->float bit conversion rounds to nearest.

High-precision function values come from mpmath (50 digits).  tests/test_faddeeva.py
evaluates the generated file with a numpy fp32 emulation of the kernel's arithmetic against
scipy.special.wofz and checks that this generator reproduces the committed file.
"""
import os

import mpmath as mp
import numpy as np
from numpy.polynomial import chebyshev as C
from numpy.polynomial import polynomial as P

mp.mp.dps = 50
KAPPA = float(mp.sqrt(1 / mp.log(2)))      # sqrt(log2 e)
N_INT = 16                                 # core intervals
WIDTH = 0.5                                # in X
DEG_CORE = (7, 5, 3)                       # A, B, C
X_WING_MIN = 5.3                           # in x: wing polynomials valid for x >= this
DEG_WING = (6, 3, 1)                       # G1, G3, G5
DEG_EXP2 = 8
Y_MAX = 0.1
Y_MIN = 1e-9
NODES = np.cos(np.pi * (np.arange(96) + 0.5) / 96)


def _dawson_derivs(x, nmax=5):
    x = mp.mpf(x)
    f = mp.sqrt(mp.pi) / 2 * mp.exp(-x * x) * mp.erfi(x)
    d = [f, 1 - 2 * x * f]
    for n in range(1, nmax):
        d.append(-2 * n * d[n - 1] - 2 * x * d[n])
    return d


def _abc(x):
    d = _dawson_derivs(x)
    c = 2 / mp.sqrt(mp.pi)
    return float(c * d[1]), float(-c * d[3] / 6), float(c * d[5] / 120)


def _g135(u):
    if u == 0:
        s = float(1 / mp.sqrt(mp.pi))
        return s, -s, s
    x = 1 / mp.sqrt(mp.mpf(u))
    d = _dawson_derivs(x)
    c = 2 / mp.sqrt(mp.pi)
    return (float(-c * d[1] / u), float(c * d[3] / 6 / u ** 2),
            float(-c * d[5] / 120 / u ** 3))


def core_table():
    """(N_INT, 20) float32: A[0..7], B[0..5], C[0..3], 2 pad; ascending powers of t."""
    rows = []
    for i in range(N_INT):
        xs = (i + 0.5 + 0.5 * NODES) * WIDTH / KAPPA
        vals = np.array([_abc(x) for x in xs])
        row = []
        for k, dg in enumerate(DEG_CORE):
            row += list(C.cheb2poly(C.chebfit(NODES, vals[:, k], dg)))
        row += [0.0, 0.0]
        rows.append(row)
    return np.array(rows, dtype=np.float64).astype(np.float32)


def wing_polys():
    """Monomial coefficients in U = 1/X^2 (ascending) of G1, G3, G5 on [0, 1/(kappa 5.3)^2]."""
    umax = 1.0 / (KAPPA * X_WING_MIN) ** 2
    us = (NODES + 1) / 2 * umax
    vals = np.array([_g135(u * KAPPA ** 2) for u in us])
    out = []
    for k, dg in enumerate(DEG_WING):
        p = C.cheb2poly(C.chebfit(NODES, vals[:, k], dg))
        poly = np.array([0.0])
        for j in range(dg, -1, -1):              # substitute s = 2U/umax - 1
            poly = P.polyadd(P.polymul(poly, [-1.0, 2.0 / umax]), [p[j]])
        out.append(poly)
    # normalised so that the leading coefficient is exactly 1.0f; the fp64 lead factor of the
    # wing term carries G1(0) ~ 1/sqrt(pi) (a rounded fp32 constant term would bias every
    # wing value by its rounding error)
    g10 = out[0][0]
    return g10, [(p_ / g10).astype(np.float32) for p_ in out]


def exp2_poly():
    """Ascending monomial coefficients of 2^f on [-1/2, 1/2], times (1 + 2^-25)."""
    fs = 0.5 * NODES
    c = C.chebfit(NODES, np.array([float(mp.mpf(2) ** mp.mpf(float(f))) for f in fs]), DEG_EXP2)
    poly = np.array([0.0])
    p = C.cheb2poly(c)
    for j in range(DEG_EXP2, -1, -1):            # s = 2 f
        poly = P.polyadd(P.polymul(poly, [0.0, 2.0]), [p[j]])
    return poly * (1.0 + 2.0 ** -25)


def render():
    tab, (g10, wing), e2 = core_table(), wing_polys(), exp2_poly()
    fh = lambda v: float(v).hex()
    lines = ["// generated by tools/gen_voigt_tables.py -- do not edit",
             f"#define RJP_VT_KAPPA {fh(KAPPA)}  /* sqrt(log2 e) = {KAPPA!r} */",
             f"#define RJP_VT_NI {N_INT}",
             f"#define RJP_VT_ROW 20",
             f"#define RJP_VT_XWING2 {fh((KAPPA * X_WING_MIN) ** 2)}  /* (kappa 5.3)^2 */",
             f"#define RJP_VT_XCORE2 {fh((N_INT * WIDTH) ** 2)}  /* table end, X^2 */",
             f"#define RJP_VT_Y_MAX {fh(Y_MAX)}",
             f"#define RJP_VT_Y_MIN {fh(Y_MIN)}",
             "// core table rows: A t^0..t^7, B t^0..t^5, C t^0..t^3, 2 pad (fp32)",
             "#define RJP_VT_CORE \\"]
    lines.append(", \\\n".join("  " + ", ".join(fh(v) + "f" for v in row) for row in tab))
    lines.append(f"#define RJP_VT_G10 {fh(g10)}  /* G1(0) ~ 1/sqrt(pi); G1, G3, G5 are divided by it */")
    for name, poly in zip(("G1", "G3", "G5"), wing):
        lines.append(f"#define RJP_VT_{name} " + ", ".join(fh(v) + "f" for v in poly))
    lines.append("#define RJP_VT_EXP2 " + ", ".join(fh(v) for v in e2))
    return "\n".join(lines) + "\n"


def main():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "rajepy_b200", "csrc", "rjp_voigt_tables.inc")
    with open(out, "wt") as f:
        f.write(render())
    print("wrote", out)


if __name__ == "__main__":
    main()
