"""Edge cases of the line-of-sight pass against the numpy oracle: channel lists that are not
equally spaced (the kernels then read the offsets instead of forming them), more channels
than one launch holds (channel blocks), channel counts that do not fill the last group,
a single channel, and a grid that no ray of which crosses the jet."""
import os
import tempfile

import numpy as np
import pytest
import scipy.constants as con

from tests import cases
from tests.parity import assert_parity, cancellation_floor_ff, cancellation_floor_line

pytestmark = pytest.mark.gpu


def _pair(params):
    import rajepy_b200 as rb
    from oracle import rajepy_oracle as orc
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    return rb.JetModel(params, log=log), orc.OracleJet(params)


def _check_cubes(jm, oj, chans):
    tau = assert_parity(jm.optical_depth_rrl('H58a', chans), oj.optical_depth_rrl('H58a', chans),
                        "tau_rrl", floor=1e-290)
    fl = np.nan_to_num(cancellation_floor_line(oj, chans) + cancellation_floor_ff(oj, chans),
                       nan=0.0, posinf=0.0)
    assert_parity(jm.flux_rrl('H58a', chans, contsub=False),
                  oj.flux_rrl('H58a', chans, contsub=False), "S_rrl", floor=fl)
    return tau


def test_unequally_spaced_channels():
    from oracle import rajepy_oracle as orc
    jm, oj = _pair(cases.case_small())
    jm.time = oj.time = 1.0 * con.year
    nu0 = orc.rrl_nu_0('H', 58, 1)
    rng = np.random.default_rng(11)
    chans = np.sort(nu0 + rng.uniform(-1.2e7, 1.2e7, 45))
    ln = jm._line_structs('H58a', chans, jm._device())[0]
    assert ln.chan_step == 0.0                       # takes the table-reading instantiation
    assert _check_cubes(jm, oj, chans) < 2e-7
    even = cases.line_channels(nu0, 45, 3e5)
    assert jm._line_structs('H58a', even, jm._device())[0].chan_step == pytest.approx(3e5)


@pytest.mark.parametrize("nchan", [1, 7, 37, 130])
def test_partial_channel_groups(nchan):
    from oracle import rajepy_oracle as orc
    jm, oj = _pair(cases.case_small())
    jm.time = oj.time = 0.7 * con.year
    chans = cases.line_channels(orc.rrl_nu_0('H', 58, 1), nchan, 4e5)
    _check_cubes(jm, oj, chans)


def test_more_channels_than_one_launch():
    """2100 channels: two channel blocks (8 channels x 256 threads per launch)."""
    from oracle import rajepy_oracle as orc
    p = cases.with_grid(cases.base_params(), 12, 28, 36)
    jm, oj = _pair(p)
    jm.time = oj.time = 0.4 * con.year
    chans = cases.line_channels(orc.rrl_nu_0('H', 58, 1), 2100, 2e4)
    got = jm.optical_depth_rrl('H58a', chans)
    ref = oj.optical_depth_rrl('H58a', chans)
    assert_parity(got, ref, "tau_rrl 2100 channels", floor=1e-290)
    assert_parity(jm.flux_rrl('H58a', chans, contsub=False),
                  oj.flux_rrl('H58a', chans, contsub=False), "S_rrl 2100 channels",
                  floor=np.nan_to_num(cancellation_floor_line(oj, chans) +
                                      cancellation_floor_ff(oj, chans), nan=0.0, posinf=0.0))


def test_grid_that_misses_the_jet():
    """Launch radius beyond the grid: no cell is inside the jet, every product is the
    reference's constant (EM = tau = 0, intensity / flux NaN)."""
    p = cases.with_grid(cases.base_params(), 16, 20, 24)
    p["geometry"]["r_0"] = 500.
    jm, oj = _pair(p)
    assert int(jm.n_verts_inside().sum()) == 0 == int(oj.n_verts_inside().sum())
    assert jm._n_active() == 0
    chans = cases.line_channels(3.285e10, 9, 1e6)
    assert np.array_equal(jm.emission_measure(), np.zeros((16, 24)))
    assert np.array_equal(jm.optical_depth_ff(np.array([5e9, 1e10])), np.zeros((2, 16, 24)))
    assert np.isnan(jm.flux_ff(5e9)).all()
    assert np.array_equal(jm.optical_depth_rrl('H58a', chans), np.zeros((9, 16, 24)))
    assert np.isnan(jm.flux_rrl('H58a', chans, contsub=False)).all()
    lm = jm.los_means()
    assert np.isnan(lm["number_density"]).all() and np.isnan(lm["n_max"])


def test_uncollapsed_optical_depths():
    """collapse=False (classes.py:1173-1177, :1379-1383): per-cell optical depths."""
    from oracle import rajepy_oracle as orc
    jm, oj = _pair(cases.case_small())
    jm.time = oj.time = 1.0 * con.year
    chans = cases.line_channels(orc.rrl_nu_0('H', 58, 1), 3, 2e6)
    got = jm.optical_depth_rrl('H58a', chans, collapse=False)
    ref = oj.optical_depth_rrl('H58a', chans, collapse=False)
    assert got.shape == ref.shape == (3, jm.nx, jm.ny, jm.nz)
    assert_parity(got, ref, "tau_rrl cells", rtol=1e-9)
    assert_parity(jm.optical_depth_rrl('H58a', float(chans[1]), collapse=False), ref[1],
                  "tau_rrl cells scalar", rtol=1e-9)
    assert_parity(np.nansum(got, axis=2), jm.optical_depth_rrl('H58a', chans), "sum == collapsed")
    f = np.array([5e9, 4.3e10])
    assert_parity(jm.optical_depth_ff(f, collapse=False), oj.optical_depth_ff(f, collapse=False),
                  "tau_ff cells", rtol=1e-9)


def test_user_assigned_temperature_and_ion_fraction():
    """The `temperature` / `ion_fraction` setters (classes.py:936-940, :994-1000): assigning
    0.5 T and 2 x everywhere must give the products of a jet with T_0 / 2 and 2 x_0."""
    import copy
    import rajepy_b200 as rb
    from oracle import rajepy_oracle as orc
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    p = cases.case_small()
    jm = rb.JetModel(copy.deepcopy(p), log=log)
    t3d, x3d = jm.temperature, jm.ion_fraction
    jm.temperature = 0.5 * t3d
    jm.ion_fraction = 2.0 * x3d
    q = copy.deepcopy(p)
    q["properties"]["T_0"] *= 0.5
    q["properties"]["x_0"] *= 2.0
    oj = orc.OracleJet(q)
    jm.time = oj.time = 1.0 * con.year
    assert_parity(jm.emission_measure(), oj.emission_measure(), "EM", rtol=1e-12)
    f = np.array([5e9, 4.3e10])
    # the Gaunt factor of the q_T = 0 branch is a scalar of params T_0 (classes.py:1421-1425),
    # which a setter does not touch: cache the oracle's halved temperature grid, then give it
    # the T_0 the GPU model still carries in its parameter dict
    oj.temperature()
    oj.p["properties"]["T_0"] = p["properties"]["T_0"]
    assert_parity(jm.optical_depth_ff(f), oj.optical_depth_ff(f), "tau_ff", rtol=1e-12)
    chans = cases.line_channels(orc.rrl_nu_0('H', 58, 1), 16, 4e5)
    assert_parity(jm.optical_depth_rrl('H58a', chans), oj.optical_depth_rrl('H58a', chans),
                  "tau_rrl", floor=1e-290)
    assert np.array_equal(np.nan_to_num(jm.temperature), np.nan_to_num(0.5 * t3d))


def test_user_assigned_launch_times_and_velocities():
    """The `ts` / `vel` setters (classes.py:857-859, :1097-1099): the integrators then read the
    travel time / line-of-sight velocity of every cell from the assigned grids."""
    import copy
    import rajepy_b200 as rb
    from oracle import rajepy_oracle as orc
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    p = cases.case_inclined()
    jm = rb.JetModel(copy.deepcopy(p), log=log)
    oj = orc.OracleJet(copy.deepcopy(p))
    jm.time = oj.time = 1.1 * con.year
    base = (jm.emission_measure(), jm.flux_ff(5e9))
    # slower material (travel times stretched) and a sheared velocity field
    travel = oj.travel_time() * 1.7 + 3e5
    vx, vlos, vz = oj.vel()
    shear = 4.0 * np.sin(np.arange(vlos.shape[1]) / 5.0)[None, :, None]
    new_v = (vx, vlos + shear, vz)
    jm.ts = travel              # the reference's setter takes the travel-time grid
    jm.vel = new_v
    oj._c["tt"] = travel
    oj._c["vel"] = new_v
    assert np.array_equal(np.nan_to_num(jm.ts), np.nan_to_num(jm.time - travel))
    assert jm.vel is new_v
    assert_parity(jm.chi_xyz, oj.chi_xyz(), "chi", rtol=1e-12)
    em = jm.emission_measure()
    assert_parity(em, oj.emission_measure(), "EM")
    assert not np.array_equal(em, base[0])          # the bursts sit elsewhere now
    f = np.array([5e9, 4.3e10])
    assert_parity(jm.optical_depth_ff(f), oj.optical_depth_ff(f), "tau_ff")
    assert_parity(jm.flux_ff(f), oj.flux_ff(f), "S_ff", floor=cancellation_floor_ff(oj, f))
    chans = cases.line_channels(orc.rrl_nu_0('H', 58, 1), 12, 5e5)
    assert_parity(jm.optical_depth_rrl('H58a', chans), oj.optical_depth_rrl('H58a', chans),
                  "tau_rrl", floor=1e-290)
    fl = np.nan_to_num(cancellation_floor_line(oj, chans) + cancellation_floor_ff(oj, chans))
    assert_parity(jm.flux_rrl('H58a', chans, contsub=False),
                  oj.flux_rrl('H58a', chans, contsub=False), "S_rrl", floor=fl)
    jm.release()
