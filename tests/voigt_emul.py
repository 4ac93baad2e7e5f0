"""numpy emulation of the kernel's mixed-precision Voigt evaluation (`voigt_fast_core` /
`voigt_fast_wing` in rajepy_b200/csrc/rjp_device.cuh), operation by operation: fp32
products and FMAs are formed exactly in float64 and rounded once to float32, the integer /
exponent-bit tricks are done on the bit patterns.  Reads the constants from the generated
rjp_voigt_tables.inc, so the test checks the file the kernels are compiled from."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "rajepy_b200", "csrc", "rjp_voigt_tables.inc")
f32 = np.float32


def _hexes(s):
    return [float.fromhex(h) for h in re.findall(r"(-?0x[0-9a-f.]+p[+-]\d+)", s)]


def load_tables(path=INC):
    txt = open(path).read()
    defs = {}
    for m in re.finditer(r"#define (RJP_VT_\w+) (.*?)(?=\n#define|\n//|\Z)", txt, re.S):
        defs[m.group(1)] = m.group(2)
    t = {
        "kappa": _hexes(defs["RJP_VT_KAPPA"])[0],
        "ni": int(defs["RJP_VT_NI"].split()[0]),
        "xwing2": _hexes(defs["RJP_VT_XWING2"])[0],
        "xcore2": _hexes(defs["RJP_VT_XCORE2"])[0],
        "y_max": _hexes(defs["RJP_VT_Y_MAX"])[0],
        "y_min": _hexes(defs["RJP_VT_Y_MIN"])[0],
        "g10": _hexes(defs["RJP_VT_G10"])[0],
        "g1": np.array(_hexes(defs["RJP_VT_G1"]), dtype=f32),
        "g3": np.array(_hexes(defs["RJP_VT_G3"]), dtype=f32),
        "g5": np.array(_hexes(defs["RJP_VT_G5"]), dtype=f32),
        "exp2": np.array(_hexes(defs["RJP_VT_EXP2"])),
    }
    core = np.array(_hexes(defs["RJP_VT_CORE"]), dtype=f32)
    t["core"] = core.reshape(t["ni"], 20)
    return t


def fmaf(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64)
            + np.asarray(c, np.float64)).astype(f32)


def mulf(a, b):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64)).astype(f32)


def horner32(coeffs_ascending, t):
    r = np.broadcast_to(f32(coeffs_ascending[-1]), np.shape(t)).astype(f32)
    for c in coeffs_ascending[-2::-1]:
        r = fmaf(r, t, f32(c))
    return r


def float_to_double_bits(v):
    """(double) of a positive normal float through the exponent re-bias the kernel uses."""
    b = np.asarray(v, f32).view(np.uint32).astype(np.uint64)
    hi = (b >> np.uint64(3)) + np.uint64(0x38000000)
    lo = (b << np.uint64(29)) & np.uint64(0xFFFFFFFF)
    return ((hi << np.uint64(32)) | lo).view(np.float64)


def double_to_float_trunc(v):
    """float of a positive normal double by dropping mantissa bits (no rounding)."""
    b = np.asarray(v, np.float64).view(np.uint64)
    hi = (b >> np.uint64(32)) - np.uint64(0x38000000)
    lo = b & np.uint64(0xFFFFFFFF)
    return (((hi << np.uint64(3)) | (lo >> np.uint64(29))) & np.uint64(0xFFFFFFFF)) \
        .astype(np.uint32).view(f32)


def core_threshold(y, t):
    """X^2 below which a cell of Lorentz ratio y takes the core path: the larger of the
    wing polynomials' validity limit and the point where exp(-x^2) < 1e-8 K
    (vt_cell_constants: fp32 logs, nudged to the safe side)."""
    lc = f32(np.log(f32(1e8) * f32(1.7724538509055159) / f32(y))) + f32(1e-3)
    x2 = lc + f32(3.3322045)
    for _ in range(2):
        x2 = lc + f32(np.log(x2))
    return float(min(max(t["xwing2"], float(x2) * t["kappa"] ** 2), t["xcore2"] - 0.01))


def cell_constants(y, t):
    """Per-cell constants of make_fast_entry (fp64 unless noted)."""
    k = t["kappa"]
    yy = (k * y) ** 2
    y2f = f32(y * y)
    g1, g3, g5 = t["g1"], t["g3"], t["g5"]
    yyf = f32(yy)
    c = g1.copy()
    c[1] = fmaf(yyf, g3[0], g1[1])
    c[2] = fmaf(yyf, fmaf(yyf, g5[0], g3[1]), g1[2])
    c[3] = fmaf(yyf, fmaf(yyf, g5[1], g3[2]), g1[3])
    c[4] = fmaf(yyf, g3[3], g1[4])
    return {"yy": yy, "yf": f32(y), "y2f": y2f, "ya": f32(2.0 * y / k), "c": c,
            "lead_wing": y * k * k * t["g10"], "xc2": core_threshold(y, t)}


def voigt_fast(x, y, t=None):
    """Re w(x + iy) for arrays x (true units) and a scalar y in [y_min, y_max]."""
    t = t or load_tables()
    k = t["kappa"]
    cc = cell_constants(y, t)
    X = np.asarray(x, np.float64) * k
    X2 = X * X
    core = X2 < cc["xc2"]
    out = np.empty_like(X)
    # ---- wings: K = y kappa^2 U P(U), U = 1/X^2 in fp64, P in fp32
    Xw2 = np.where(core, 1e3, X2)
    U = 1.0 / Xw2
    Uf = double_to_float_trunc(U)
    Pw = horner32(cc["c"], Uf)
    Kw = cc["lead_wing"] * U * float_to_double_bits(Pw)
    # ---- core
    Xc = np.where(core, np.abs(X), 0.0)
    q = np.rint(Xc * 2.0 ** 25).astype(np.int64)
    idx = np.minimum(q >> 24, t["ni"] - 1)
    frac = (q & 0xFFFFFF).astype(f32)
    tt = fmaf(frac, f32(2.0 ** -23), f32(-1.0))
    rows = t["core"][idx]                                   # (n, 20)
    A = np.empty_like(tt)
    B = np.empty_like(tt)
    Cc = np.empty_like(tt)
    for i in range(t["ni"]):
        m = idx == i
        if m.any():
            A[m] = horner32(t["core"][i, 0:8], tt[m])
            B[m] = horner32(t["core"][i, 8:14], tt[m])
            Cc[m] = horner32(t["core"][i, 14:18], tt[m])
    Hy = fmaf(cc["y2f"], fmaf(cc["y2f"], Cc, B), A)
    T = cc["yy"] - Xc * Xc
    i_ = np.rint(T)
    f = T - i_
    # device: 2^f by MUFU.EX2 on float(f) (<= 2 ulp of fp32; emulated as the correctly rounded
    # value), the integer part added to the exponent field
    Gf = f32(np.ldexp(np.exp2(f.astype(f32).astype(np.float64)).astype(f32).astype(np.float64),
                      i_.astype(np.int64)))
    Xf = mulf(q.astype(f32), f32(2.0 ** -25))
    a = mulf(Xf, cc["ya"])
    a2 = mulf(a, a)
    cm1 = mulf(a2, horner32([-1 / 2, 1 / 24, -1 / 720, 1 / 40320, -1 / 3628800], a2))
    Kc = fmaf(-cc["yf"], Hy, fmaf(Gf, cm1, Gf))
    return np.where(core, float_to_double_bits(np.maximum(Kc, f32(1e-37))), Kw)
