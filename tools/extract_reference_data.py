"""
One-off data extraction (run in the build container, where /root/reference is mounted).

Writes the two published data tables the hot path consumes into rajepy_b200/data/ in a
re-packed (ALTERED, see NOTICE there) form:

  * gaunt_vanhoof2014.npz  -- thermally averaged free-free Gaunt factors of
    van Hoof et al. (2014, MNRAS 444, 420), the 146 x 81 table the reference reads from
    files/vanHoofetal2014.data (maths/physics.py:626-663).  Only the g_ff block and
    the grid description are kept (uncertainty block dropped).
  * atomic_masses.json -- AME2003 atomic masses [micro-u] of the twelve isotopes the
    reference can look up (_constants.py:7-10 NZ table via maths/physics.py:607-624).
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("RAJEPY_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "rajepy_b200", "data")


def main():
    with open(os.path.join(REF, "files", "vanHoofetal2014.data"), "rt") as f:
        lines = f.readlines()
    magic = int(lines[28].split("#")[0])
    n_g2, n_u = (int(_) for _ in lines[29].split("#")[0].split())
    lg2_start = float(lines[30].split("#")[0])
    lu_start = float(lines[31].split("#")[0])
    step = float(lines[32].split("#")[0])
    rows = [[float(v) for v in ln.split()] for ln in lines[42:42 + n_u]]
    gff = np.array(rows, dtype=np.float64)
    assert gff.shape == (n_u, n_g2) == (146, 81), gff.shape
    np.savez(os.path.join(OUT, "gaunt_vanhoof2014.npz"), gff=gff,
             log_gamma2_start=lg2_start, log_u_start=lu_start, step=step,
             magic=magic)

    import pandas as pd
    ams = pd.read_pickle(os.path.join(REF, "files", "atomic_masses.pkl"))
    nz = {"H": (1, 0), "He": (2, 2), "Li": (3, 4), "Be": (4, 5), "B": (5, 6),
          "C": (6, 6), "N": (7, 7), "O": (8, 8), "F": (9, 10), "Ne": (10, 10),
          "Na": (11, 12), "Mg": (12, 12)}
    out = {}
    for el, (z, n) in nz.items():
        m = ams[(ams["N"] == n) & (ams["Z"] == z)]["mass[micro-u]"].values[0]
        out[el] = {"Z": z, "N": n, "mass_micro_u": float(m)}
    with open(os.path.join(OUT, "atomic_masses.json"), "wt") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
