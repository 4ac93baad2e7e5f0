"""
Fuzz the numpy oracle against the UNMODIFIED reference (build container only; the reference is
imported through oracle/ref_shim.py): random jet parameters on small grids -- geometry, all
seven power-law indices (incl. the 2F1 travel time and the q_T != 0 branch), bursts, lines --
and every product of the hot path compared.  Prints the worst relative deviation per product
and exits non-zero if the oracle leaves its pin tolerance.
    python tools/fuzz_oracle_vs_reference.py [n_cases] [seed]
"""
import os
import sys
import time

import numpy as np
import scipy.constants as con

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rajepy_oracle as orc, ref_shim  # noqa: E402
from tests import cases  # noqa: E402


def random_params(rng):
    p = cases.with_grid(cases.base_params(), int(rng.integers(5, 11)) * 2,
                        int(rng.integers(6, 14)) * 2, int(rng.integers(8, 20)) * 2,
                        cs=float(rng.choice([0.25, 0.3, 0.5, 0.7, 1.0])))
    p["geometry"].update({
        "inc": float(rng.uniform(20., 90.)), "pa": float(rng.uniform(-180., 180.)),
        "epsilon": float(rng.uniform(0.5, 1.2)), "opang": float(rng.uniform(10., 50.)),
        "w_0": float(rng.uniform(0.5, 2.0)), "r_0": float(rng.uniform(0.5, 2.0)),
        "rotation": str(rng.choice(["CW", "CCW"]))})
    if rng.random() < 0.5:       # plain jet: no cross-sectional laws, maybe isothermal
        pl = {"q_v": float(rng.choice([0., rng.uniform(-0.5, 0.5)])),
              "q_T": float(rng.choice([0., rng.uniform(-0.5, 0.2)])),
              "q_x": float(rng.choice([0., rng.uniform(-0.5, 0.5)])),
              "q^d_n": 0., "q^d_T": 0., "q^d_v": 0., "q^d_x": 0.}
    else:
        pl = {"q_v": float(rng.uniform(-0.5, 0.5)), "q_T": 0.,
              "q_x": float(rng.uniform(-0.5, 0.5)),
              "q^d_n": float(rng.uniform(-1., 1.)), "q^d_T": float(rng.uniform(-0.5, 0.5)),
              "q^d_v": float(rng.uniform(-1., 1.)), "q^d_x": float(rng.uniform(-0.5, 0.5))}
    p["power_laws"].update(pl)
    p["properties"].update({"v_0": float(rng.uniform(100., 600.)),
                            "x_0": float(rng.uniform(0.02, 0.9)),
                            "T_0": float(rng.uniform(5e3, 2e4)),
                            "mlr_bj": float(10 ** rng.uniform(-8, -6)),
                            "mlr_rj": float(10 ** rng.uniform(-8, -6))})
    p["target"].update({"dist": float(rng.uniform(100., 3000.)),
                        "v_lsr": float(rng.uniform(-30., 30.)),
                        "M_star": float(rng.uniform(0.3, 15.)),
                        "R_1": float(rng.uniform(0.1, 0.5)), "R_2": float(rng.uniform(1., 5.))})
    nb = int(rng.integers(0, 5))
    p["ejection"] = {"t_0": rng.uniform(0., 3., nb), "hl": rng.uniform(0.05, 0.6, nb),
                     "chi": rng.uniform(0.2, 10., nb),
                     "which": np.array([str(rng.choice(["R", "B", "RB"])) for _ in range(nb)])}
    return p


def worst(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.shape != b.shape or not np.array_equal(np.isnan(a), np.isnan(b)):
        return np.inf
    m = ~np.isnan(b)
    if not np.array_equal(a[m] == 0, b[m] == 0):
        return np.inf
    nz = m & (b != 0)
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))) if nz.any() else 0.0


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    ref_shim.load_reference()
    tol = {"nverts": 0.0, "em": 1e-11, "tau_ff": 1e-11, "I_ff": 1e-11, "S_ff": 1e-11,
           "tau_rrl": 1e-10, "S_rrl": 1e-10, "S_rrl_cs": 1e-8, "I_rrl": 1e-8}
    over = {k: 0.0 for k in tol}
    bad = 0
    t0 = time.time()
    for c in range(n_cases):
        p = random_params(rng)
        import copy
        jm = ref_shim.make_reference_model(copy.deepcopy(p))
        oj = orc.OracleJet(copy.deepcopy(p))
        t = float(rng.uniform(0., 4.)) * con.year
        jm.time = oj.time = t
        line = str(rng.choice(["H58a", "H30a", "He42b", "H92a", "H110g"]))
        el, n, dn = orc.rrl_parser(line)
        nu0 = orc.rrl_nu_0(el, n, dn)
        chans = cases.line_channels(nu0, 5, float(rng.choice([2e5, 1e6, 5e6])))
        freqs = np.array([1.4e9, 2.2e10, 2.3e11])
        ff = jm.fill_factor
        nv = oj.n_verts_inside()
        ffo = np.where(nv == 8, 1.0, np.where(nv > 0, 0.5, np.nan))
        res = {"nverts": 0.0 if (np.array_equal(np.isnan(ff), np.isnan(ffo)) and
                                  np.array_equal(np.nan_to_num(ff), np.nan_to_num(ffo)))
               else np.inf,
               "em": worst(oj.emission_measure(), jm.emission_measure()),
               "tau_ff": worst(oj.optical_depth_ff(freqs), jm.optical_depth_ff(freqs)),
               "I_ff": worst(oj.intensity_ff(freqs), jm.intensity_ff(freqs)),
               "S_ff": worst(oj.flux_ff(freqs), jm.flux_ff(freqs)),
               "tau_rrl": worst(oj.optical_depth_rrl(line, chans),
                                jm.optical_depth_rrl(line, chans)),
               "S_rrl": worst(oj.flux_rrl(line, chans, contsub=False),
                              jm.flux_rrl(line, chans, contsub=False)),
               "S_rrl_cs": worst(oj.flux_rrl(line, chans, contsub=True),
                                 jm.flux_rrl(line, chans, contsub=True)),
               # (the reference's intensity_rrl only takes one channel: classes.py:1270 does not
               # broadcast an array of frequencies against the temperature map)
               "I_rrl": worst(oj.intensity_rrl(line, float(chans[1])),
                              jm.intensity_rrl(line, float(chans[1])))}
        flag = [k for k, v in res.items() if v > tol[k]]
        for k, v in res.items():
            over[k] = max(over[k], v)
        in_jet = int((nv > 0).sum())
        print(f"case {c:3d} grid {oj.nx}x{oj.ny}x{oj.nz} in-jet {in_jet:6d} "
              f"inc {p['geometry']['inc']:5.1f} pa {p['geometry']['pa']:7.1f} "
              f"q^d_v {p['power_laws']['q^d_v']:+.2f} q_T {p['power_laws']['q_T']:+.2f} "
              f"bursts {len(p['ejection']['t_0'])} {line:6s} "
              + ("OK" if not flag else "DEVIATES: " + ", ".join(f"{k}={res[k]:.2e}"
                                                                for k in flag)), flush=True)
        bad += bool(flag)
    print("worst relative deviations:", {k: f"{v:.2e}" for k, v in over.items()},
          f"({time.time() - t0:.0f} s)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
