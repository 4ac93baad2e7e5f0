"""The Voigt profile of the channel loop: the coefficients compiled into the kernels
(rajepy_b200/csrc/rjp_weideman.inc) evaluated with the kernel's own real two-term
recurrence (restated in numpy) against scipy.special.wofz, the reference's routine
(maths/rrls.py:353)."""
import os
import re

import numpy as np
from scipy.special import wofz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    txt = open(os.path.join(ROOT, "rajepy_b200", "csrc", "rjp_weideman.inc")).read()
    n = int(re.search(r"#define RJP_WEIDEMAN_N (\d+)", txt).group(1))
    ell = float.fromhex(re.search(r"#define RJP_WEIDEMAN_L (\S+)", txt).group(1))
    body = txt.split("#define RJP_WEIDEMAN_COEFFS")[1]
    coeffs = [float.fromhex(h) for h in re.findall(r"(-?0x[0-9a-f.]+p[+-]\d+)", body)]
    assert len(coeffs) == n
    return ell, np.array(coeffs)


def faddeeva_re(x, y):
    """numpy restatement of rjp::faddeeva_re (rjp_device.cuh)."""
    ell, a = _coefficients()
    dr, nr = ell + y, ell - y
    inv = 1.0 / (dr * dr + x * x)
    zr = (nr * dr - x * x) * inv
    zi = 2.0 * ell * x * inv
    s, q = 2.0 * zr, zr * zr + zi * zi
    b1 = np.zeros_like(x)
    b2 = np.zeros_like(x)
    for c in a[:-1]:
        b1, b2 = c + s * b1 - q * b2, b1
    pr = a[-1] + zr * b1 - q * b2
    pi = zi * b1
    ur, ui = dr * inv, x * inv
    return 2.0 * (pr * (ur * ur - ui * ui) - pi * (2.0 * ur * ui)) + ur / np.sqrt(np.pi)


def test_generator_is_reproducible():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gen_weideman", os.path.join(ROOT, "tools", "gen_weideman.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ell, a = mod.coeffs()
    ell2, a2 = _coefficients()
    assert ell == ell2 and np.array_equal(a, a2)


def test_voigt_accuracy_against_wofz():
    xs = np.linspace(-45.0, 45.0, 3601)
    worst = {}
    for y in (1e-4, 1e-3, 1e-2, 0.1, 0.38, 1.0, 3.0, 7.5, 20.0, 100.0):
        ref = wofz(xs + 1j * y).real
        rel = np.abs(faddeeva_re(xs, y) - ref) / ref
        worst[y] = rel.max()
    assert worst[1e-4] < 5e-8
    assert worst[1e-3] < 5e-9
    assert all(worst[y] < 5e-10 for y in worst if y >= 1e-2), worst


def test_voigt_far_wings_stay_relative():
    for y in (1e-3, 0.1, 5.0):
        xs = np.array([60.0, 200.0, 1e3, 1e4, 1e6])
        ref = wofz(xs + 1j * y).real
        rel = np.abs(faddeeva_re(xs, y) - ref) / ref
        assert rel.max() < 1e-6, (y, rel)


# ---------------------------------------------------------------- mixed-precision split
def test_fast_tables_generator_is_reproducible():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gen_voigt_tables", os.path.join(ROOT, "tools", "gen_voigt_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    committed = open(os.path.join(ROOT, "rajepy_b200", "csrc", "rjp_voigt_tables.inc")).read()
    assert mod.render() == committed


def test_fast_voigt_emulation_against_wofz():
    """The fp64/fp32 split of the channel loop, emulated operation by operation in numpy
    (tests/voigt_emul.py), against scipy.special.wofz over the whole fast class."""
    from tests import voigt_emul as ve
    t = ve.load_tables()
    xs = np.linspace(-60.0, 60.0, 48001)
    for y in (1e-9, 1e-7, 1e-5, 1e-3, 4.5e-3, 0.011, 0.03, 0.06, 0.1):
        ref = wofz(xs + 1j * y).real
        rel = np.abs(ve.voigt_fast(xs, y, t) / ref - 1.0)
        assert rel.max() < 3e-7, (y, rel.max(), xs[rel.argmax()])
        assert np.sqrt(np.mean(rel ** 2)) < 8e-8, y


def test_fast_voigt_far_wings():
    from tests import voigt_emul as ve
    t = ve.load_tables()
    xs = np.array([80.0, 300.0, 1e3, 1e4, 1e6])
    for y in (1e-6, 1e-2, 0.1):
        ref = wofz(xs + 1j * y).real
        assert np.abs(ve.voigt_fast(xs, y, t) / ref - 1.0).max() < 2e-7
