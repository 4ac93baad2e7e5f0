"""BASELINE.json configurations on the GPU.

configs[1] (256^3, 16 continuum frequencies) and configs[3] (burst time series) are checked
against the oracle directly (at sizes it finishes in a minute or two); configs[2] (512^3 x
256 channels) and configs[4] (1024^3 x 512 channels) are far beyond the oracle's reach
(45 / 350 GB of numpy temporaries), so they are checked through size-independent properties:
mirror symmetry of the sky images, the contsub identity S(contsub=False) - S(contsub=True)
= S_ff, independence of the result from how the channels are batched, from empty padding
along the line of sight and from the x-slab decomposition."""
import os
import tempfile

import numpy as np
import pytest
import scipy.constants as con

from tests import cases
from tests.parity import assert_parity, cancellation_floor_ff

pytestmark = pytest.mark.gpu


def _model(params, **kw):
    import rajepy_b200 as rb
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    return rb.JetModel(params, log=log, **kw)


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def test_config2_256cube_16_frequencies():
    """configs[1]: 256^3, continuum at 16 frequencies 1-300 GHz, against the oracle."""
    from oracle import rajepy_oracle as orc
    p = cases.with_grid(cases.base_params(), 256, 256, 256)
    freqs = np.logspace(9, np.log10(3e11), 16)
    jm, oj = _model(p), orc.OracleJet(cases.with_grid(cases.base_params(), 256, 256, 256))
    assert np.array_equal(jm.n_verts_inside(), oj.n_verts_inside().astype(np.uint8))
    assert_parity(jm.emission_measure(), oj.emission_measure(), "EM")
    assert_parity(jm.optical_depth_ff(freqs), oj.optical_depth_ff(freqs), "tau_ff")
    assert_parity(jm.flux_ff(freqs), oj.flux_ff(freqs), "S_ff",
                  floor=cancellation_floor_ff(oj, freqs))


def test_config4_burst_time_series():
    """configs[3]: 4 bursts over epochs 0..5 yr (a subset of the 64), 5 GHz, 96^3."""
    from oracle import rajepy_oracle as orc
    p = cases.with_grid(cases.base_params(), 96, 96, 96)
    jm, oj = _model(p), orc.OracleJet(cases.with_grid(cases.base_params(), 96, 96, 96))
    totals = []
    for yr in np.linspace(0., 5., 64)[::9]:
        jm.time = oj.time = yr * con.year
        s, so = jm.flux_ff(5e9), oj.flux_ff(5e9)
        assert_parity(s, so, f"S_ff t={yr:.2f}", floor=cancellation_floor_ff(oj, 5e9)[0])
        assert_parity(jm.emission_measure(), oj.emission_measure(), f"EM t={yr:.2f}")
        totals.append(np.nansum(s))
    assert max(totals) > 1.05 * totals[0]      # the bursts do brighten the jet


def test_config4_time_series_driver():
    """rajepy_b200.flux_ff_time_series (one fill, one pass per epoch) == per-epoch flux_ff."""
    import rajepy_b200 as rb
    p = cases.with_grid(cases.base_params(), 64, 64, 96)
    epochs = np.linspace(0., 5., 16) * con.year
    series = rb.flux_ff_time_series(p, epochs, 5e9)
    assert series.shape == (16, 64, 96)
    jm = _model(cases.with_grid(cases.base_params(), 64, 64, 96))
    for e in (0, 5, 15):
        jm.time = float(epochs[e])
        ref = jm.flux_ff(5e9)
        assert np.array_equal(np.isnan(series[e]), np.isnan(ref))
        assert np.array_equal(np.nan_to_num(series[e]), np.nan_to_num(ref))


def test_epoch_batch_equals_single_epochs_and_oracle():
    """rjp_integrate_epochs (one ray walk for a batch of model times, 11 epochs = one full and
    one partial block of 8) against the per-epoch pass and against the oracle."""
    from oracle import rajepy_oracle as orc
    p = cases.with_grid(cases.base_params(), 48, 64, 80)
    jm, oj = _model(p), orc.OracleJet(cases.with_grid(cases.base_params(), 48, 64, 80))
    epochs = np.linspace(0., 5., 11) * con.year
    freqs = np.array([5e9, 43e9])
    flux, em = jm._continuum_epochs_device(epochs, freqs, 'flux', with_em=True)
    flux = flux.cpu().numpy().reshape(11, 2, 48, 80)
    em = em.cpu().numpy().reshape(11, 48, 80)
    tau = jm._continuum_epochs_device(epochs, freqs, 'tau').cpu().numpy().reshape(11, 2, 48, 80)
    for e in range(11):
        jm.time = oj.time = float(epochs[e])
        one = jm.flux_ff(freqs)
        assert np.array_equal(np.isnan(flux[e]), np.isnan(one))
        np.testing.assert_allclose(np.nan_to_num(flux[e]), np.nan_to_num(one), rtol=1e-13,
                                   atol=0.0)
        np.testing.assert_allclose(em[e], jm.emission_measure(), rtol=1e-13, atol=0.0)
        np.testing.assert_allclose(tau[e], jm.optical_depth_ff(freqs), rtol=1e-13, atol=0.0)
        if e in (0, 4, 10):
            assert_parity(em[e], oj.emission_measure(), f"EM epoch {e}")
            assert_parity(tau[e], oj.optical_depth_ff(freqs), f"tau_ff epoch {e}")
            assert_parity(flux[e], oj.flux_ff(freqs), f"S_ff epoch {e}",
                          floor=cancellation_floor_ff(oj, freqs))


def test_epoch_batch_matches_reference_time_series(golden_dir):
    """The batched walk (and the public flux_ff_time_series driver) against the time series the
    unmodified reference wrote (tests/golden/series.npz, tools/make_golden_series.py)."""
    import os
    import rajepy_b200 as rb
    from tests.parity import flux_floors_uniform_t
    g = np.load(os.path.join(golden_dir, "series.npz"))
    p = cases.case_series()
    nx, _, nz = (int(v) for v in g["dims"])
    times = g["epochs_yr"] * con.year
    jm = _model(p)
    flux, em = jm._continuum_epochs_device(times, g["freqs"], 'flux', with_em=True)
    tau = jm._continuum_epochs_device(times, g["freqs"], 'tau')
    ne, nf = len(times), len(g["freqs"])
    flux = flux.cpu().numpy().reshape(ne, nf, nx, nz)
    tau = tau.cpu().numpy().reshape(ne, nf, nx, nz)
    em = em.cpu().numpy().reshape(ne, nx, nz)
    floor = flux_floors_uniform_t(p, g["freqs"], p["properties"]["T_0"])[0]
    for e in range(ne):
        assert_parity(em[e], g["em"][e], f"EM epoch {e}")
        assert_parity(tau[e], g["tauff"][e], f"tau_ff epoch {e}")
        assert_parity(flux[e], g["sff"][e], f"S_ff epoch {e}", floor=floor[:, None, None])
    series = rb.flux_ff_time_series(cases.case_series(), times, 4.3e10)
    assert series.shape == (ne, nx, nz)
    for e in range(ne):
        assert_parity(series[e], g["sff"][e, 1], f"series epoch {e}", floor=floor[1])


def test_epoch_batch_user_travel_times():
    """A user-assigned `ts` grid (classes.py:857-859: the setter stores the array the getter
    subtracts from `time`, i.e. a travel time) is honoured by the batched walk."""
    p = cases.with_grid(cases.base_params(), 32, 48, 64)
    jm = _model(p)
    jm.time = 2.0 * con.year
    travel = jm.time - jm.ts
    slower = np.where(np.isnan(travel), np.nan, travel + 0.3 * con.year)
    jm.ts = slower
    epochs = np.array([1.0, 2.0, 3.5]) * con.year
    flux = jm._continuum_epochs_device(epochs, 5e9, 'flux').cpu().numpy().reshape(3, 32, 64)
    plain = _model(cases.with_grid(cases.base_params(), 32, 48, 64))
    for e in range(3):
        jm.time = plain.time = float(epochs[e])
        one = jm.flux_ff(5e9)
        assert np.array_equal(np.isnan(flux[e]), np.isnan(one))
        np.testing.assert_allclose(np.nan_to_num(flux[e]), np.nan_to_num(one), rtol=1e-12,
                                   atol=0.0)
    assert not np.allclose(np.nan_to_num(one), np.nan_to_num(plain.flux_ff(5e9)), rtol=1e-6)


@pytest.mark.parametrize("n,nch", [(512, 256), (1024, 512)])
def test_large_grid_properties(n, nch):
    """configs[2] (512^3, 256 channels) and configs[4] (1024^3, 512 channels)."""
    import rajepy_b200 as rb
    p = cases.with_grid(cases.base_params(), n, n, n)
    jm = _model(p)
    jm.time = 1.0 * con.year
    nu0 = rb.hostmath.rrl_nu_0('H', 58, 1)
    chans = nu0 + (np.arange(nch) - (nch - 1) / 2.) * 1e5
    em = jm.emission_measure()
    # (1) the edge-on jet is mirror symmetric in x (bursts only break the z symmetry)
    assert np.array_equal(em, em[::-1])
    assert np.count_nonzero(em) > 0.02 * em.size
    # (2) contsub identity, NaN pattern == rays that miss the jet
    s_all = jm.flux_rrl('H58a', chans, contsub=False)
    assert np.array_equal(np.isnan(s_all[0]), em == 0)
    sub = slice(nch // 2 - 4, nch // 2 + 4)
    s_cs = jm.flux_rrl('H58a', chans[sub], contsub=True)
    s_ff = jm.flux_ff(chans[sub])
    ok = ~np.isnan(s_ff)
    # both terms are prefactor * (1 - exp(-tau)) formed literally like the reference does:
    # for optically thin edge pixels that carries an absolute noise of ~ulp(1) * prefactor
    omega = np.arctan(0.5 * con.au / (120. * con.parsec)) ** 2. / 1e-26
    pref = (2. * chans[sub] ** 2. * con.k * 1e4 / con.c ** 2. * omega)[:, None, None]
    lim = 2e-7 * np.abs(s_cs) + 1e-13 * s_ff + 32 * np.finfo(float).eps * pref
    assert np.all(np.abs(s_all[sub] - (s_cs + s_ff))[ok] <= lim[ok])
    # (3) how the channels are batched does not matter beyond the fp32 part of the Voigt split:
    # another channel count means another thread layout, the channel offsets are then formed
    # with different roundings (1 ulp of fp64), which can flip an fp32 rounding inside single
    # evaluations (6e-8 each; far less on the sums).  The same call twice is bit-identical.
    tau = jm.optical_depth_rrl('H58a', chans)
    half = jm.optical_depth_rrl('H58a', chans[:nch // 2])
    np.testing.assert_allclose(tau[:nch // 2], half, rtol=2e-7, atol=0)
    one = jm.optical_depth_rrl('H58a', float(chans[7]))
    np.testing.assert_allclose(tau[7], one, rtol=2e-7, atol=0)
    jm._line = None
    assert np.array_equal(jm.optical_depth_rrl('H58a', chans), tau)
    assert tau.min() >= 0.0 and np.array_equal(tau[0] > 0, em > 0)
    del s_all, s_cs, tau, half
    jm.release()


def test_padding_and_slab_invariance():
    """Empty cells appended along the line of sight, or cutting the grid into x-slabs,
    must not change a single bit of the sky images."""
    import torch
    p = cases.with_grid(cases.base_params(), 128, 128, 192)
    jm = _model(p)
    jm.time = 0.8 * con.year
    em, tau = jm.emission_measure(), jm.optical_depth_ff(5e9)
    jp = _model(cases.with_grid(cases.base_params(), 128, 160, 192))
    jp.time = jm.time
    assert np.array_equal(jp.emission_measure(), em)
    assert np.array_equal(jp.optical_depth_ff(5e9), tau)
    rows = []
    for r in range(4):
        js = _model(cases.with_grid(cases.base_params(), 128, 128, 192), shard=(r, 4))
        js.time = jm.time
        c = js._pass()                       # slab-local device images, no gather
        lo, hi = js.slab
        rows.append(c["em"].view(hi - lo, js.nz).cpu().numpy())
    assert np.array_equal(np.concatenate(rows, axis=0), em)
