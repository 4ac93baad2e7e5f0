"""Parity helpers shared by the GPU tests and __graft_entry__.smoke()."""
import numpy as np
import scipy.constants as con

RTOL = 1e-6  # BASELINE.json north_star: EM, tau, intensity, flux within 1e-6 relative


def assert_parity(got, ref, what, rtol=RTOL, floor=0.0):
    """|got - ref| <= rtol |ref| + floor with identical NaN masks and identical zero
    masks (EM / tau are exactly 0 on rays that miss the jet, intensity / flux NaN).

    `floor` (same units as the data) exists only for quantities the reference forms as
    1 - exp(-tau): for tau << 1 the reference's own result carries an absolute rounding
    noise of a few ulp(1) times the prefactor, which no implementation can track to 1e-6
    relative."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{what}: NaN masks differ"
    m = ~np.isnan(ref)
    if np.isscalar(floor) and floor == 0.0:
        assert np.array_equal(got[m] == 0, ref[m] == 0), f"{what}: zero masks differ"
    fl = np.broadcast_to(np.asarray(floor, dtype=np.float64), ref.shape)
    err = np.abs(got[m] - ref[m])
    lim = rtol * np.abs(ref[m]) + np.nan_to_num(fl[m])
    bad = err > lim
    if bad.any():
        rel = err[bad] / np.maximum(np.abs(ref[m][bad]), 1e-300)
        raise AssertionError(f"{what}: {bad.sum()} of {m.sum()} values outside tolerance; "
                             f"worst relative error {rel.max():.3e}")
    nz = m & (ref != 0)
    if nz.any():
        return float(np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])))
    return 0.0


def cancellation_floor_ff(oj, freqs):
    """8 ulp(1) x (flux of an optically thick pixel) per frequency, shape (nf, nx, nz)."""
    tm = oj.mean_temperature()
    om = oj.pixel_solid_angle() / 1e-26
    f = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    pref = (2. * f ** 2. * con.k / con.c ** 2.)[:, None, None] * tm[None] * om
    return 8 * np.finfo(np.float64).eps * np.abs(pref)


def cancellation_floor_line(oj, chans):
    tm = oj.mean_temperature()
    om = oj.pixel_solid_angle() / 1e-26
    f = np.atleast_1d(np.asarray(chans, dtype=np.float64))[:, None, None]
    with np.errstate(all='ignore'):
        bnu = 2. * con.h * 1e7 * f ** 3. / (con.c * 1e2) ** 2. / \
            np.expm1(con.h * f / (con.k * tm[None])) * 1e-3 * om
    return 8 * np.finfo(np.float64).eps * np.abs(bnu)


def flux_floors_uniform_t(params, freqs, t_mean):
    """The two cancellation floors above for a jet of uniform temperature (q_T = q^d_T = 0:
    the mean temperature of every jet-crossing ray is T_0 exactly), without an oracle run:
    (continuum floor, line floor), each (nf,) in Jy/pixel."""
    f = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
    om = np.arctan((params["grid"]["c_size"] * con.au) /
                   (params["target"]["dist"] * con.parsec)) ** 2. / 1e-26
    eps8 = 8 * np.finfo(np.float64).eps
    ff = eps8 * 2. * f ** 2. * con.k / con.c ** 2. * t_mean * om
    bnu = 2. * con.h * 1e7 * f ** 3. / (con.c * 1e2) ** 2. / \
        np.expm1(con.h * f / (con.k * t_mean)) * 1e-3 * om
    return ff, eps8 * np.abs(bnu)
