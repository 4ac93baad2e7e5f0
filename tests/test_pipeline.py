"""Pipeline-side glue against fixtures the UNMODIFIED reference wrote
(tools/make_golden_pipeline.py): run list, product names, channel grids, and loading of the
reference's own save files.  CPU only."""
import json
import os
import tempfile

import numpy as np

from rajepy_b200 import pipeline as pl
from rajepy_b200.compat import load_pickle
from tests import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    with open(os.path.join(GOLD, "pipeline_runs.json")) as f:
        return json.load(f)


def test_run_list_matches_reference_pipeline():
    """Order, names and channel grids of classes.py:2116-2172 / :1875-1900 / :1954-1967."""
    gold = _gold()["runs"]
    _, params = cases.pipeline_case()
    runs = pl.build_runs("pl" + os.sep, params)
    assert len(runs) == len(gold)
    for run, g in zip(runs, gold):
        assert run.obs_type == g["obs_type"]
        assert run.year == g["year"] and run.day == g["day"]
        assert run.freq == g["freq"]                       # RRL: rrl_nu_0 bit for bit
        assert getattr(run, "line", None) == g["line"]
        assert run.bandwidth == g["bandwidth"] and run.chanwidth == g["chanwidth"]
        assert run.nchan == g["nchan"]
        assert np.array_equal(run.chan_freqs, np.array(g["chan_freqs"]))
        rel = lambda p: os.path.relpath(p, "pl")  # noqa: E731
        assert rel(run.rt_dcy) == g["rt_dcy"]
        assert rel(run.fits_flux) == g["fits_flux"]
        assert rel(run.fits_tau) == g["fits_tau"]
        assert rel(run.fits_em) == g["fits_em"]
        assert run.radiative_transfer == g["radiative_transfer"]
        assert run.simobserve == g["simobserve"]
        assert not run.completed and run.results == {} and run.products == {}


def test_freq_str():
    assert pl.freq_str(5e9) == "5GHz"
    assert pl.freq_str(4.3e10) == "43GHz"
    assert pl.freq_str(1.5e6, ".1f") == "1.5MHz"
    assert pl.freq_str([1e3, 2e12]) == ["1kHz", "2THz"]


def test_reference_model_save_file_loads():
    """A pickle written by the reference's JetModel.save (classes.py:1704-1713) names
    RaJePy.logger classes; compat maps them."""
    loaded = load_pickle(os.path.join(GOLD, "ref_jetmodel.save"))
    model, _ = cases.pipeline_case()
    assert set(loaded) == {"params", "areas", "ffs", "time", "log"}
    assert loaded["ffs"].shape == (12, 16, 28) and loaded["areas"].shape == (12, 16, 28)
    assert loaded["time"] == 0.75 * cases.YEAR
    assert loaded["params"]["grid"]["n_x"] == model["grid"]["n_x"]
    log = loaded["log"]
    assert len(log.entries) > 0 and "INFO" in str(log.entries[0])


def test_reference_pipeline_save_file_loads():
    """Pipeline.save of the reference (classes.py:2215-2258) -> our run classes."""
    with tempfile.TemporaryDirectory() as tmp:
        runs, params, model_file, log = pl.load_pipeline(os.path.join(GOLD, "ref_pipeline.save"))
    gold = _gold()["runs"]
    assert [type(r).__name__ for r in runs] == ["ContinuumRun"] * 4 + ["RRLRun"]
    for run, g in zip(runs, gold):
        assert run.year == g["year"] and run.freq == g["freq"]
        assert np.array_equal(run.chan_freqs, np.array(g["chan_freqs"]))
        assert os.path.basename(run.fits_flux) == os.path.basename(g["fits_flux"])
        # the reference had stored its totals in the runs before saving
        assert np.allclose(run.results["flux"], g["flux"], rtol=0, atol=0)
    assert model_file.endswith("jetmodel.save")
    assert set(params) >= {"continuum", "rrls", "dcys"}


def test_pipeline_state_round_trip():
    _, params = cases.pipeline_case()
    with tempfile.TemporaryDirectory() as tmp:
        params["dcys"]["model_dcy"] = os.path.join(tmp, "pl")
        runs = pl.build_runs(params["dcys"]["model_dcy"], params)
        runs[1].results["flux"] = 1.25
        runs[1].completed = True
        f = os.path.join(tmp, "pipeline.save")
        pl.save_pipeline(f, runs, params, os.path.join(tmp, "pl", "jetmodel.save"))
        runs2, params2, model_file, log = pl.load_pipeline(f)
        assert [r.completed for r in runs2] == [False, True, False, False, False]
        assert runs2[1].results["flux"] == 1.25
        assert runs2[4].line == "H58a" and runs2[4].fits_tau == runs[4].fits_tau
        assert model_file == os.path.join(tmp, "pl", "jetmodel.save")


def test_total_flux_reductions():
    """classes.py:2461-2472"""
    _, params = cases.pipeline_case()
    runs = pl.build_runs("pl", params)
    rng = np.random.default_rng(1)
    cube = rng.random((3, 4, 5))
    cube[:, 0, 0] = np.nan
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)      # mean of the all-NaN pixel
        want = np.nansum(np.nanmean(cube, axis=0))
    assert pl.total_flux(runs[0], cube) == want
    assert np.array_equal(pl.total_flux(runs[4], cube), np.nansum(np.nansum(cube, axis=1), axis=1))
