"""SURVEY 8 rows f1-f3 through the CUDA JetModel: the RT block of Pipeline.execute
(`rajepy_b200.pipeline.run_rt`), FITS products written through `savefits=`, and the
save / load_model checkpoint -- against fixtures written by the unmodified reference
(tools/make_golden_pipeline.py)."""
import copy
import json
import os
import tempfile

import numpy as np
import pytest
import scipy.constants as con

from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _log(tmp):
    import rajepy_b200 as rb
    return rb.logger.Log(os.path.join(tmp, "m.log"), verbose=False)


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def test_run_rt_products_names_and_totals():
    import rajepy_b200 as rb
    from rajepy_b200 import pipeline as pl
    from rajepy_b200.fitsio import read_fits
    with open(os.path.join(GOLD, "pipeline_runs.json")) as f:
        gold = json.load(f)
    model_p, params = cases.pipeline_case()
    with tempfile.TemporaryDirectory() as tmp:
        dcy = os.path.join(tmp, "pl")
        os.mkdir(dcy)
        params["dcys"]["model_dcy"] = dcy
        jm = rb.JetModel(model_p, log=_log(tmp))
        runs = pl.build_runs(dcy, params)
        save_file = os.path.join(dcy, gold["save_file"])
        pl.run_rt(jm, runs, params=params, save_file=save_file)
        for run, g in zip(runs, gold["runs"]):
            for key in ("fits_em", "fits_tau", "fits_flux"):
                assert os.path.exists(os.path.join(dcy, g[key])), g[key]
            # totals as the reference reduces its own arrays (classes.py:2461-2472)
            assert np.allclose(run.results["flux"], g["flux"], rtol=1e-6, atol=0)
            assert run.completed
            hdr, data = read_fits(run.fits_flux)
            assert data.shape == (run.nchan, jm.nz, jm.nx)          # (freq, Dec, RA)
            assert hdr["BUNIT"] == "Jy pixel^-1" and hdr["CTYPE3"] == "FREQ"
            assert hdr["CDELT3"] == (run.chanwidth if run.nchan > 1 else 1.0)
        assert os.path.exists(os.path.join(dcy, gold["model_file"]))
        assert os.path.exists(save_file)
        # second execution: nothing is rewritten, the totals come from the files
        stamp = {r.fits_flux: os.path.getmtime(r.fits_flux) for r in runs}
        runs2 = pl.build_runs(dcy, params)
        pl.run_rt(jm, runs2)
        for r2, r1 in zip(runs2, runs):
            assert os.path.getmtime(r2.fits_flux) == stamp[r2.fits_flux]
            # (the file holds (freq, Dec, RA): another summation order, a few ulp)
            assert np.allclose(r2.results["flux"], r1.results["flux"], rtol=1e-12, atol=0)
        # resume: completed runs are skipped entirely
        runs3, params3, model_file, _ = pl.load_pipeline(save_file)
        assert [r.completed for r in runs3] == [True] * 4 + [False]   # saved before the flag
        for r in runs3:
            r.results.pop("flux", None)
        pl.run_rt(rb.JetModel.load_model(model_file), runs3, resume=True)
        assert all("flux" not in r.results for r in runs3[:4]) and "flux" in runs3[4].results
        # clobber rewrites
        os.utime(runs[0].fits_flux, (1, 1))
        pl.run_rt(jm, runs[:1], clobber=True)
        assert os.path.getmtime(runs[0].fits_flux) > 1
        jm.release()


def test_savefits_through_every_rt_method():
    """`savefits=` of the six RT methods writes what the method returns, in FITS axis order
    (classes.py:1543-1652; miscellaneous/functions.py:236-301)."""
    import rajepy_b200 as rb
    from rajepy_b200.fitsio import read_fits
    p = cases.case_small()
    chans = cases.line_channels(rb.hostmath.rrl_nu_0('H', 58, 1), 4, 1e6)
    freqs = np.array([5e9, 6e9, 7e9])
    with tempfile.TemporaryDirectory() as tmp:
        jm = rb.JetModel(p, log=_log(tmp))
        jm.time = 1.0 * con.year
        f = lambda n: os.path.join(tmp, n + ".fits")  # noqa: E731
        calls = {"em": (jm.emission_measure(savefits=f("em")), 'pc cm^-6'),
                 "tau": (jm.optical_depth_ff(freqs, savefits=f("tau")), 'dimensionless'),
                 "tau1": (jm.optical_depth_ff(5e9, savefits=f("tau1")), 'dimensionless'),
                 "int": (jm.intensity_ff(freqs, savefits=f("int")), 'W m^-2 Hz^-1 sr^-1'),
                 "flux": (jm.flux_ff(freqs, savefits=f("flux")), 'Jy pixel^-1'),
                 "taul": (jm.optical_depth_rrl('H58a', chans, savefits=f("taul")),
                          'dimensionless'),
                 "intl": (jm.intensity_rrl('H58a', chans, savefits=f("intl")),
                          'W m^-2 Hz^-1 sr^-1'),
                 "fluxl": (jm.flux_rrl('H58a', chans, contsub=False, savefits=f("fluxl")),
                           'Jy pixel^-1')}
        for name, (arr, bunit) in calls.items():
            hdr, data = read_fits(f(name))
            want = np.swapaxes(arr, -1, -2)         # (.., nx, nz) -> (.., Dec = z, RA = x)
            assert _same(data, want), name
            assert hdr["BUNIT"] == bunit and hdr["NAXIS"] == arr.ndim
            assert hdr["NAXIS1"] == jm.nx and hdr["NAXIS2"] == jm.nz
            assert hdr["CRPIX1"] == jm.nx / 2 + 0.5 and hdr["CRPIX2"] == jm.nz / 2 + 0.5
        hdr, _ = read_fits(f("fluxl"))
        assert hdr["CRPIX3"] == 4 / 2. + 0.5 and hdr["CDELT3"] == 1e6
        assert hdr["CRVAL3"] == chans[4 // 2 - 1] + 1e6 / 2
        # the header cannot drift: card-by-card text of one cube
        with open(f("fluxl"), "rb") as fh:
            raw = fh.read(2880 * 4).decode("ascii", errors="replace")
        cards = [raw[i:i + 80].rstrip() for i in range(0, len(raw), 80)]
        cards = cards[:cards.index("END") + 1]
        with open(os.path.join(GOLD, "fits_header_flux_rrl.txt")) as fh:
            want_cards = fh.read().split("\n")
        assert cards == want_cards
        jm.release()


def test_save_load_round_trip_and_reference_checkpoint():
    """save -> load_model -> identical products; and a checkpoint written by the reference's
    own JetModel.save is adopted (classes.py:48-88, :1704-1713)."""
    import rajepy_b200 as rb
    from rajepy_b200.compat import load_pickle
    chans = cases.line_channels(rb.hostmath.rrl_nu_0('H', 58, 1), 5, 1e6)
    with tempfile.TemporaryDirectory() as tmp:
        jm = rb.JetModel(cases.case_inclined(), log=_log(tmp))
        jm.time = 0.8 * con.year
        before = (jm.emission_measure(), jm.flux_ff(np.array([5e9, 2.2e10])),
                  jm.optical_depth_rrl('H58a', chans), jm.flux_rrl('H58a', chans, contsub=False))
        ff, ar = jm.fill_factor, jm.areas
        path = os.path.join(tmp, "jetmodel.save")
        jm.save(path)
        jm.release()
        saved = load_pickle(path)
        assert set(saved) == {"params", "areas", "ffs", "time", "log"}
        assert saved["ffs"].dtype == np.float64 and saved["ffs"].shape == ff.shape
        jm2 = rb.JetModel.load_model(path)
        assert jm2.time == 0.8 * con.year
        assert _same(jm2.fill_factor, ff) and _same(jm2.areas, ar)
        after = (jm2.emission_measure(), jm2.flux_ff(np.array([5e9, 2.2e10])),
                 jm2.optical_depth_rrl('H58a', chans),
                 jm2.flux_rrl('H58a', chans, contsub=False))
        for a, b in zip(before, after):
            assert _same(a, b)
        jm2.release()
    # the reference's own save file
    ref = load_pickle(os.path.join(GOLD, "ref_jetmodel.save"))
    jm3 = rb.JetModel.load_model(os.path.join(GOLD, "ref_jetmodel.save"),
                                 log=_log(tempfile.mkdtemp()))
    assert jm3.time == ref["time"] == 0.75 * con.year
    assert _same(jm3.fill_factor, ref["ffs"]) and _same(jm3.areas, ref["areas"])
    model_p, _ = cases.pipeline_case()
    fresh = rb.JetModel(model_p, log=_log(tempfile.mkdtemp()))
    fresh.time = jm3.time
    assert _same(jm3.flux_ff(5e9), fresh.flux_ff(5e9))
    # a fill-factor grid that differs from the model's own is adopted, not recomputed
    ffs = ref["ffs"].copy()
    idx = np.argwhere(ffs == 1.0)[0]
    ffs[tuple(idx)] = 0.5
    jm3._adopt_fill_factor(ffs)
    assert jm3.fill_factor[tuple(idx)] == 0.5
    assert not _same(jm3.emission_measure(), fresh.emission_measure())
    jm3.release()
    fresh.release()


def test_derived_property_grids():
    """mass_density / pressure (classes.py:901-908, :1002-1007) against the oracle."""
    import rajepy_b200 as rb
    from oracle import rajepy_oracle as orc
    p = cases.case_powerlaws()
    jm = rb.JetModel(copy.deepcopy(p), log=_log(tempfile.mkdtemp()))
    oj = orc.OracleJet(copy.deepcopy(p))
    jm.time = oj.time = 1.5 * con.year
    n_o = oj.number_density()
    mu_mh = p["properties"]["mu"] * orc.atomic_mass("H")
    for got, want in ((jm.mass_density, mu_mh * 1e3 * n_o),
                      (jm.pressure, n_o * oj.temperature() * con.k * 1e7)):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        m = ~np.isnan(want)
        assert np.max(np.abs(got[m] / want[m] - 1.0)) < 1e-9
    jm.release()
