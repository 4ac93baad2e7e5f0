"""
Generate tests/golden/series.npz: the variable-ejection time series of tests/cases.case_series
computed by the UNMODIFIED reference (/root/reference through oracle/ref_shim.py), one
`jm.time = t; emission_measure(); optical_depth_ff(freqs); flux_ff(freqs)` per model time --
what Pipeline does per run year (classes.py:2347-2453).  Build container only:
    python tools/make_golden_series.py

Stored: dims (3,), epochs_yr (ne,), freqs (nf,), em (ne, nx, nz), tauff / sff (ne, nf, nx, nz).
"""
import os
import sys
import time

import numpy as np
import scipy.constants as con

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import cases  # noqa: E402


def main():
    ref_shim.load_reference()
    jm = ref_shim.make_reference_model(cases.case_series())
    t0 = time.time()
    em, tau, flux = [], [], []
    for yr in cases.SERIES_EPOCHS_YR:
        jm.time = yr * con.year
        em.append(jm.emission_measure())
        tau.append(jm.optical_depth_ff(cases.SERIES_FREQS))
        flux.append(jm.flux_ff(cases.SERIES_FREQS))
    out = {"dims": np.array([jm.nx, jm.ny, jm.nz]), "epochs_yr": cases.SERIES_EPOCHS_YR,
           "freqs": cases.SERIES_FREQS, "em": np.array(em), "tauff": np.array(tau),
           "sff": np.array(flux)}
    tot = np.nansum(out["sff"][:, 0], axis=(1, 2))
    assert tot.max() > 1.05 * tot[0], "the bursts should brighten the jet"
    path = os.path.join(ROOT, "tests", "golden", "series.npz")
    np.savez_compressed(path, **out)
    print(f"series: dims={tuple(out['dims'])} epochs={len(em)} ({time.time() - t0:.1f}s) -> "
          f"{os.path.getsize(path) / 1e3:.0f} kB; 5 GHz totals [Jy]: "
          + " ".join(f"{t:.4e}" for t in tot))


if __name__ == "__main__":
    main()
