"""Line cubes at BASELINE scale against the UNMODIFIED reference.

tests/golden/big_*.npz hold outputs of the reference itself (tools/make_golden_big.py) for
  c2rrl  256^3 / 0.5 au -- 16 channels picked from the 512-channel H58a grid of configs[4]
         (core, shoulders, far wings; read from the channel array) + 16 equally spaced ones,
  r256   128 x 128 x 512 / 1.0 au -- the jet out to |r| = 256 au, the range of Lorentz/Gauss
         ratios and cells per ray that the 1024^3 grid of configs[4] reaches, where the
         fast / fp64 class split of the Voigt routine is exercised,
as the columns of the jet-crossing rays (everything else is the constant 0 / NaN, asserted).
Bar: 1e-6 relative with identical masks (north_star)."""
import os
import tempfile

import numpy as np
import pytest

from tests import cases
from tests.parity import assert_parity, flux_floors_uniform_t

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(params):
    import rajepy_b200 as rb
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    return rb.JetModel(params, log=log)


@pytest.mark.parametrize("name", ["r256", "c2rrl"])
def test_line_cube_vs_reference_fixture(name):
    z = np.load(os.path.join(GOLD, f"big_{name}.npz"))
    factory, epoch_yr, line, sets, extras, _ = cases.BIG_CASES[name]
    p = factory()
    jm = _model(p)
    jm.time = epoch_yr * cases.YEAR
    assert (jm.nx, jm.ny, jm.nz) == tuple(int(v) for v in z["dims"])
    rays = z["rays"]
    cols = rays[z["sel"]]
    em = jm.emission_measure().ravel()
    # identical masks: exactly the reference's rays cross the jet
    assert np.array_equal(np.flatnonzero(em != 0), rays)
    assert_parity(em[cols], z["em"], "EM")
    t0 = p["properties"]["T_0"]
    worst = {}
    for tag in sets:
        chans = z[f"chans_{tag}"]
        nch = chans.size
        ff_floor, l_floor = flux_floors_uniform_t(p, chans, t0)
        tau = jm.optical_depth_rrl(line, chans).reshape(nch, -1)
        assert (np.delete(tau, rays, axis=1) == 0).all()
        worst[f"tau_{tag}"] = assert_parity(tau[:, cols], z[f"taurrl_{tag}"], f"tau_rrl {tag}")
        s = jm.flux_rrl(line, chans, contsub=False).reshape(nch, -1)
        assert np.isnan(np.delete(s, rays, axis=1)).all()
        worst[f"s_{tag}"] = assert_parity(s[:, cols], z[f"srrl_{tag}"], f"S_rrl {tag}",
                                          floor=(ff_floor + l_floor)[:, None])
        if tag in extras:
            s = jm.flux_rrl(line, chans, contsub=True).reshape(nch, -1)
            assert np.isnan(np.delete(s, rays, axis=1)).all()
            assert_parity(s[:, cols], z[f"srrl_cs_{tag}"], f"S_rrl contsub {tag}",
                          floor=l_floor[:, None])
            om = jm._pixel_solid_angle() / 1e-26
            i_l = jm.intensity_rrl(line, chans).reshape(nch, -1)
            assert np.isnan(np.delete(i_l, rays, axis=1)).all()
            assert_parity(i_l[:, cols], z[f"irrl_{tag}"], f"I_rrl {tag}",
                          floor=(l_floor / om)[:, None])
            # scalar-frequency branch (what flux_rrl calls, classes.py:1321-1322)
            one = jm.intensity_rrl(line, float(chans[3])).ravel()
            assert_parity(one[cols], z[f"irrl_{tag}"][3], f"I_rrl scalar {tag}",
                          floor=float(l_floor[3] / om))
    print(f"{name}: worst relative errors {worst}")
    jm.release()
