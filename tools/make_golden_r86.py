"""
Generate tests/golden/r86.npz: values of the reference's Reynolds (1986) analytic flux,
maths/physics.py:297-374 `flux_expected_r86`, from the UNMODIFIED reference (imported through
oracle/ref_shim.py) for the parity cases.  Run in the build container only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import cases  # noqa: E402

CASES = ("small", "inclined", "nobursts", "tgrad")
FREQS = (1e9, 5e9, 4.3e10, 3e11)
YMAX = (0.3, 1.0)
YMIN = (None, 0.05)


def main():
    rjp = ref_shim.load_reference()
    out = {}
    for name in CASES:
        jm = ref_shim.make_reference_model(cases.CASES[name][0]())
        vals = np.empty((len(FREQS), 2, len(YMAX), len(YMIN)))
        for i, f in enumerate(FREQS):
            for j, which in enumerate("RB"):
                for k, ymax in enumerate(YMAX):
                    for m, ymin in enumerate(YMIN):
                        vals[i, j, k, m] = float(rjp.maths.physics.flux_expected_r86(
                            jm, f, which, ymax, ymin))
        out[name] = vals
        print(name, vals[1, :, 1, 0])
    np.savez(os.path.join(ROOT, "tests", "golden", "r86.npz"), freqs=np.array(FREQS),
             ymax=np.array(YMAX), ymin=np.array([np.nan, 0.05]), **out)


if __name__ == "__main__":
    main()
