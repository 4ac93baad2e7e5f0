"""The numpy oracle (oracle/rajepy_oracle.py) is pinned against outputs of the UNMODIFIED
reference stored in tests/golden/*.npz (generator: tools/make_golden.py)."""
import os

import numpy as np
import pytest
import scipy.constants as con

from oracle import rajepy_oracle as orc
from tests import cases

SMALL_CASES = ["small", "inclined", "powerlaws", "nobursts", "tgrad"]


def rel_close(a, b, rtol):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN masks differ"
    m = ~np.isnan(b)
    assert np.array_equal(a[m] == 0, b[m] == 0), "zero masks differ"
    nz = m & (b != 0)
    if nz.any():
        err = np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))
        assert err <= rtol, f"max rel err {err:.3e} > {rtol}"


@pytest.mark.parametrize("name", SMALL_CASES + ["c1"])
def test_oracle_matches_reference_fixture(name, golden_dir):
    path = os.path.join(golden_dir, f"{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing")
    g = np.load(path)
    factory, epochs, freqs, line, nch, chanw = cases.CASES[name]
    oj = orc.OracleJet(factory())
    assert (oj.nx, oj.ny, oj.nz) == tuple(int(v) for v in g["dims"])
    nv = oj.n_verts_inside()
    assert np.array_equal(nv.astype(np.uint8), g["nverts"]), "vertex counts not bit-exact"
    idx = g["jet_idx"]
    assert np.array_equal(np.flatnonzero(nv.ravel() > 0), idx)
    rel_close(oj.p["properties"]["n_0"], g["n_0"], 1e-15)
    rel_close(-oj.travel_time().ravel()[idx], g["ts0"], 1e-12)
    rel_close(oj.nd_base().ravel()[idx], g["nd_base"], 1e-13)
    rel_close(oj.ion_fraction().ravel()[idx], g["xi"], 1e-13)
    rel_close(oj.temperature().ravel()[idx], g["temp"], 1e-13)
    vx, vl, vz = oj.vel()
    rel_close(vl.ravel()[idx], g["vlos"], 1e-12)
    rel_close(vx.ravel()[idx], g["vx"], 1e-12)
    rel_close(vz.ravel()[idx], g["vz"], 1e-12)
    rel_close(oj.rreff().ravel()[idx], g["reff"], 1e-13)
    el, n, dn = orc.rrl_parser(line)
    nu0 = orc.rrl_nu_0(el, n, dn)
    assert nu0 == float(g["nu0"])
    chans = cases.line_channels(nu0, nch, chanw)
    assert np.array_equal(chans, g["chans"])
    if "gff" in g:
        for f, gv in zip(freqs, g["gff"]):
            rel_close(orc.gff(f, oj.p["properties"]["T_0"]), gv, 1e-13)
    for e, yr in enumerate(epochs):
        oj.time = yr * con.year
        rel_close(oj.chi_xyz().ravel()[idx], g[f"chi_{e}"], 1e-12)
        rel_close(oj.emission_measure(), g[f"em_{e}"], 1e-12)
        rel_close(oj.optical_depth_ff(np.array(freqs)), g[f"tauff_{e}"], 1e-12)
        rel_close(oj.intensity_ff(np.array(freqs)), g[f"iff_{e}"], 1e-12)
        rel_close(oj.flux_ff(np.array(freqs)), g[f"sff_{e}"], 1e-12)
        rel_close(oj.optical_depth_rrl(line, chans), g[f"taurrl_{e}"], 1e-11)
        rel_close(oj.flux_rrl(line, chans, contsub=False), g[f"srrl_{e}"], 1e-11)
        rel_close(oj.flux_rrl(line, chans, contsub=True), g[f"srrl_cs_{e}"], 1e-9)


def test_oracle_matches_reference_time_series(golden_dir):
    """tests/golden/series.npz: 11 model times of a bursting jet, written by the unmodified
    reference (tools/make_golden_series.py)."""
    g = np.load(os.path.join(golden_dir, "series.npz"))
    oj = orc.OracleJet(cases.case_series())
    assert (oj.nx, oj.ny, oj.nz) == tuple(int(v) for v in g["dims"])
    assert np.array_equal(g["epochs_yr"], cases.SERIES_EPOCHS_YR)
    for e, yr in enumerate(g["epochs_yr"]):
        oj.time = yr * con.year
        rel_close(oj.emission_measure(), g["em"][e], 1e-12)
        rel_close(oj.optical_depth_ff(g["freqs"]), g["tauff"][e], 1e-12)
        rel_close(oj.flux_ff(g["freqs"]), g["sff"][e], 1e-12)


def test_oracle_fuzz_against_the_reference_itself():
    """Random jets (geometry, all power-law indices, bursts, lines) through the UNMODIFIED
    reference and the oracle side by side (tools/fuzz_oracle_vs_reference.py; 150 cases are
    logged in profiles/r2_oracle_fuzz_vs_reference.txt).  Needs /root/reference: build
    container only."""
    import subprocess
    import sys
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools",
                                                       "fuzz_oracle_vs_reference.py"), "6", "3"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "worst relative deviations" in res.stdout


def test_survey_smoke_values():
    """Survey-time probe values of the reference."""
    assert orc.gff(5e9, 1e4) == pytest.approx(5.083477778218337, rel=1e-12)
    assert orc.rrl_nu_0('H', 58, 1) == pytest.approx(32852207385.82994, rel=1e-15)


def test_lz_to_grid_dims_example():
    """files/example-model-params.py as shipped (l_z = 2") -> 108 x 110 x 588
    (SURVEY.md section 6)."""
    p = cases.base_params()
    p["grid"]["l_z"] = 2.
    oj = orc.OracleJet(p)
    assert (oj.nx, oj.ny, oj.nz) == (108, 110, 588)
