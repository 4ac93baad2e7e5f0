"""
Parity-test cases (jet parameter dicts) shared by the golden-fixture generator
(`tools/make_golden.py`), the oracle tests and the GPU parity tests.

The base dict restates files/example-model-params.py:11-56 of the reference (same
values; `n_0` is derived at construction, classes.py:234-242).
"""
import copy

import numpy as np

YEAR = 31536000.0  # scipy.constants.year (365 d)


def base_params():
    return {
        "target": {"name": "test2", "ra": "04:31:34.07736", "dec": "+18:08:04.9020",
                   "epoch": "J2000", "dist": 120., "v_lsr": 6.2, "M_star": 0.55,
                   "R_1": .25, "R_2": 2.5},
        "grid": {"n_x": 50, "n_y": 400, "n_z": 50, "l_z": None, "c_size": 0.5},
        "geometry": {"epsilon": 7. / 9., "opang": 25., "w_0": 1., "r_0": 1.,
                     "inc": 90., "pa": 0., "rotation": "CCW"},
        "power_laws": {"q_v": 0., "q_T": 0., "q_x": 0., "q^d_n": 0., "q^d_T": 0.,
                       "q^d_v": 0., "q^d_x": 0.},
        "properties": {"v_0": 150., "x_0": 0.1, "T_0": 1E4, "mu": 1.3,
                       "mlr_bj": 1e-7, "mlr_rj": 5e-8},
        "ejection": {"t_0": np.array([0.5, 0.75, 1., 2.]),
                     "hl": np.array([0.15, 0.15, 0.45, 0.5]),
                     "chi": np.array([5., 5., 2.5, 10.]),
                     "which": np.array(["R", "B", "B", "RB"])},
    }


def with_grid(p, nx, ny, nz, cs=None):
    p = copy.deepcopy(p)
    p["grid"].update({"n_x": nx, "n_y": ny, "n_z": nz, "l_z": None})
    if cs is not None:
        p["grid"]["c_size"] = cs
    return p


def case_c1():
    """BASELINE config 1: 50x400x50, c_size 0.5 au, edge-on, 5 GHz."""
    return base_params()


def case_small():
    """Small edge-on grid used by the quick CPU/GPU tests."""
    return with_grid(base_params(), 20, 40, 60)


def case_inclined():
    """Inclined + rotated jet (inc=60, pa=30), clockwise rotation."""
    p = with_grid(base_params(), 40, 48, 56)
    p["geometry"].update({"inc": 60., "pa": 30., "rotation": "CW"})
    return p


def case_powerlaws():
    """Non-dyadic cell size, power-law indices switched on, cross-sectional velocity
    law (hyp2f1 travel time), cross-sectional temperature law."""
    p = with_grid(base_params(), 32, 36, 44, cs=0.3)
    p["geometry"].update({"inc": 75., "pa": -20., "epsilon": 0.9, "opang": 30.,
                          "w_0": 0.8, "r_0": 1.2})
    p["power_laws"].update({"q_v": -0.15, "q_T": 0., "q_x": -0.2, "q^d_n": -0.4,
                            "q^d_T": 0.25, "q^d_v": -0.5, "q^d_x": 0.3})
    p["ejection"] = {"t_0": np.array([0.3, 1.2]), "hl": np.array([0.2, 0.35]),
                     "chi": np.array([4., 0.25]), "which": np.array(["RB", "B"])}
    return p


def case_tgrad():
    """q_T != 0: Reynolds (1986) Gaunt branch (classes.py:1426) and the cm-vs-au
    temperature quirk (classes.py:957-962), which makes T ~ 1 K."""
    p = with_grid(base_params(), 16, 24, 40)
    p["power_laws"].update({"q_T": -0.3, "q_v": 0.2})
    return p


def case_nobursts():
    p = with_grid(base_params(), 24, 28, 36)
    p["ejection"] = {"t_0": np.array([]), "hl": np.array([]), "chi": np.array([]),
                     "which": np.array([])}
    p["geometry"].update({"inc": 35., "pa": 110.})
    return p


# name -> (params factory, epochs [yr], continuum freqs [Hz], rrl line, n channels,
#          channel width [Hz])
CASES = {
    "small": (case_small, [0.0, 1.0], [5e9, 4.3e10], "H58a", 8, 1e6),
    "inclined": (case_inclined, [0.8], [1e9, 2.2e10, 3e11], "H58a", 6, 2e6),
    "powerlaws": (case_powerlaws, [1.5], [5e9, 1e11], "He42b", 5, 3e6),
    "nobursts": (case_nobursts, [0.0], [8e9], "H30a", 4, 5e6),
    "tgrad": (case_tgrad, [0.6], [5e9], "H58a", 3, 1e6),
    "c1": (case_c1, [0.0, 1.0], [5e9], "H58a", 8, 1e6),
}


def line_channels(nu0, nchan, chanw):
    """Channel centres nu0 + (k - (nchan-1)/2) * chanw, i.e. ContinuumRun.chan_freqs
    (classes.py:1897-1900) for bandwidth = nchan * chanw centred on nu0."""
    return nu0 - nchan * chanw / 2. + chanw / 2. + np.arange(nchan) * chanw


# ------------------------------------------------------------------ large line-cube fixtures
# (tools/make_golden_big.py runs the unmodified reference; tests/test_gpu_big_parity.py)
_BENCH_PICK = (0, 40, 100, 160, 200, 230, 245, 252, 255, 256, 260, 275, 300, 350, 420, 511)


def case_c2rrl():
    """256^3, c_size 0.5 au: BASELINE configs[1] geometry with the H58a cube of configs[4]."""
    return with_grid(base_params(), 256, 256, 256)


def case_r256():
    """The jet out to |r| = 256 au (what the 1024^3 / 0.5 au grid of configs[4] reaches) on a
    128 x 128 x 512 grid with c_size 1.0 au."""
    return with_grid(base_params(), 128, 128, 512, cs=1.0)


# name -> (params factory, epoch [yr], line, {set: () -> channel offsets from nu0 [Hz]},
#          sets that also store the contsub=True flux and intensity_rrl, ray stride)
BIG_CASES = {
    "c2rrl": (case_c2rrl, 1.0, "H58a",
              {"pick": lambda: (np.array(_BENCH_PICK) - 255.5) * 1e5,
               "uni": lambda: (np.arange(16) - 7.5) * 3.2e6}, (), 1),
    "r256": (case_r256, 1.0, "H58a",
             {"pick": lambda: (np.array(_BENCH_PICK) - 255.5) * 1e5,
              "uni": lambda: (np.arange(8) - 3.5) * 4e5}, ("uni",), 3),
}


def case_series():
    """Variable-ejection time series (BASELINE configs[3] in small): the example jet, slightly
    inclined, 11 model times over 0-5 yr (`SERIES_EPOCHS_YR`), continuum at 5 and 43 GHz.
    tools/make_golden_series.py runs the unmodified reference once per epoch."""
    p = with_grid(base_params(), 24, 40, 72)
    p["geometry"].update({"inc": 80., "pa": 10.})
    return p


SERIES_EPOCHS_YR = np.linspace(0., 5., 11)
SERIES_FREQS = np.array([5e9, 4.3e10])


def pipeline_case():
    """(model params, pipeline params) of the Pipeline-glue fixtures
    (tools/make_golden_pipeline.py): a tiny grid, two epochs x two continuum bands (one with
    3 channels), one epoch x one RRL with 6 channels."""
    model = with_grid(base_params(), 12, 16, 28)
    pline = {'min_el': 20.,
             'dcys': {"model_dcy": "pl"},
             'continuum': {'times': np.array([1.0, 0.25]),
                           'freqs': np.array([5., 43.]) * 1e9,
                           't_obs': np.array([1200, 1200]),
                           'tscps': np.array([('VLA', 'A'), ('VLA', 'B')]),
                           't_ints': np.array([5, 5]),
                           'bws': np.array([.6e9, 1e9]),
                           'chanws': np.array([2.e8, 1.e9])},
             'rrls': {'times': np.array([1.0]),
                      'lines': np.array(['H58a']),
                      't_obs': np.array([3000]),
                      'tscps': np.array([('VLA', 'A')]),
                      't_ints': np.array([60]),
                      'bws': np.array([6e6]),
                      'chanws': np.array([1e6])}}
    return model, pline
