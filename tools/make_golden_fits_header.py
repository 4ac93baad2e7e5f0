"""tests/golden/fits_header_flux_rrl.txt: the header cards (one per line, trailing blanks
stripped, up to END) that `JetModel.flux_rrl(..., savefits=...)` writes for tests/cases.py:
case_small() at model time 1 yr with 4 H58a channels of 1 MHz -- the card-by-card restatement of
the reference's save_fits (classes.py:1588-1648) that tests/test_gpu_pipeline.py pins.  The
header depends only on the parameters, so no GPU is needed to regenerate it."""
import os
import sys
import tempfile

import numpy as np
import scipy.constants as con

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from rajepy_b200.fitsio import build_header  # noqa: E402
from tests import cases  # noqa: E402

jm = rb.JetModel(cases.case_small(),
                 log=rb.logger.Log(os.path.join(tempfile.mkdtemp(), "h.log"), verbose=False))
jm.time = 1.0 * con.year
chans = cases.line_channels(rb.hostmath.rrl_nu_0('H', 58, 1), 4, 1e6)
cards = build_header(jm, np.zeros((4, jm.nz, jm.nx)), 'flux', chans)
out = os.path.join(ROOT, "tests", "golden", "fits_header_flux_rrl.txt")
with open(out, "wt") as f:
    f.write("\n".join(c.rstrip() for c in cards))
print(open(out).read())
