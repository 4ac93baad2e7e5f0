"""Measured parity of the CUDA path against the numpy oracle / reference fixtures, per
product and case (max relative error over pixels where the reference is finite and
non-zero).  Prints a table; run on a GPU box:  python tools/parity_report.py"""
import os
import sys
import tempfile

import numpy as np
import scipy.constants as con

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from oracle import rajepy_oracle as orc  # noqa: E402
from tests import cases  # noqa: E402
from tests.parity import assert_parity, cancellation_floor_ff, cancellation_floor_line  # noqa: E402


def main():
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)
    print(f"{'case':10s} {'EM':>9s} {'tau_ff':>9s} {'S_ff':>9s} {'tau_rrl':>9s} {'S_rrl':>9s} "
          f"{'S_rrl cs':>9s}  nverts")
    for name, (factory, epochs, freqs, line, nch, chanw) in cases.CASES.items():
        jm, oj = rb.JetModel(factory(), log=log), orc.OracleJet(factory())
        exact = np.array_equal(jm.n_verts_inside(), oj.n_verts_inside().astype(np.uint8))
        el, n, dn = orc.rrl_parser(line)
        chans = cases.line_channels(orc.rrl_nu_0(el, n, dn), max(nch, 32), chanw / 4)
        worst = np.zeros(6)
        for yr in epochs:
            jm.time = oj.time = yr * con.year
            f = np.array(freqs)
            fl_c = cancellation_floor_ff(oj, f)
            fl_l = np.nan_to_num(cancellation_floor_line(oj, chans) +
                                 cancellation_floor_ff(oj, chans), nan=0.0, posinf=0.0)
            tiny = 1e-290
            errs = [
                assert_parity(jm.emission_measure(), oj.emission_measure(), "EM"),
                assert_parity(jm.optical_depth_ff(f), oj.optical_depth_ff(f), "tau_ff"),
                assert_parity(jm.flux_ff(f), oj.flux_ff(f), "S_ff", floor=fl_c),
                assert_parity(jm.optical_depth_rrl(line, chans),
                              oj.optical_depth_rrl(line, chans), "tau_rrl", floor=tiny),
                assert_parity(jm.flux_rrl(line, chans, contsub=False),
                              oj.flux_rrl(line, chans, contsub=False), "S_rrl", floor=fl_l),
                assert_parity(jm.flux_rrl(line, chans, contsub=True),
                              oj.flux_rrl(line, chans, contsub=True), "S_rrl cs", floor=fl_l),
            ]
            worst = np.maximum(worst, [e if e is not None else 0.0 for e in errs])
        print(f"{name:10s} " + " ".join(f"{w:9.2e}" for w in worst) +
              f"  {'bit-exact' if exact else 'DIFFERENT'}", flush=True)
        jm.release()


if __name__ == "__main__":
    main()
