"""
Generate the LARGE line-cube fixtures tests/golden/big_<case>.npz by executing the UNMODIFIED
reference (/root/reference through oracle/ref_shim.py).  Build container only, minutes of CPU:

    python tools/make_golden_big.py [c2rrl] [r256]

Cases (tests/cases.py:BIG_CASES):
  c2rrl  256^3, c_size 0.5 au, epoch 1 yr: 16 channels picked from the bench's 512-channel
         H58a grid (nu0 + (k - 255.5) * 100 kHz: line core, shoulders and far wings; NOT equally
         spaced, so the kernels read the offsets) + 16 equally spaced channels (3.2 MHz apart)
  r256   128 x 128 x 512, c_size 1.0 au: the jet reaches |r| = 256 au like the 1024^3 / 0.5 au
         grid of BASELINE configs[4], i.e. the same range of Lorentz/Gauss ratios y and of cells
         per ray, at 1/128 of the cells

Only columns of rays that cross the jet are stored (EM != 0); everywhere else the
reference has tau = 0 and flux = NaN exactly (asserted here), which the tests check from `rays`.
Stored: dims, rays (int64 flat x*nz+z of EVERY jet-crossing ray), sel (indices into `rays` of
the rays whose columns are kept: every `stride`-th, to bound the file size), em (nsel,),
chans_<set>, and per channel set taurrl_<set>, srrl_<set> (contsub=False), and for the sets
named in `extras` srrl_cs_<set> (contsub=True) and irrl_<set> (intensity_rrl), each (nch, nsel).
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


def run_case(name):
    factory, epoch_yr, line, sets, extras, stride = cases.BIG_CASES[name]
    rjp = ref_shim.load_reference()
    jm = ref_shim.make_reference_model(factory())
    jm.time = epoch_yr * con.year
    t0 = time.time()
    el, n, dn = rjp.maths.rrls.rrl_parser(line)
    nu0 = rjp.maths.rrls.rrl_nu_0(el, n, dn)
    em = jm.emission_measure()
    rays = np.flatnonzero(em.ravel() != 0)
    sel = np.arange(0, rays.size, stride)
    out = {"dims": np.array([jm.nx, jm.ny, jm.nz]), "rays": rays, "sel": sel,
           "em": em.ravel()[rays[sel]], "nu0": nu0, "epoch_yr": epoch_yr}
    print(f"{name}: fill + EM {time.time() - t0:.0f}s, {rays.size} jet-crossing rays",
          flush=True)

    def cols(a, const_nan):
        a = a.reshape(a.shape[0], -1)
        rest = np.delete(a, rays, axis=1)
        if const_nan:
            assert np.isnan(rest).all()
        else:
            assert (rest == 0).all()
        return a[:, rays[sel]]

    for tag, offs in sets.items():
        chans = nu0 + np.asarray(offs(), dtype=np.float64)
        out[f"chans_{tag}"] = chans
        out[f"taurrl_{tag}"] = cols(jm.optical_depth_rrl(line, chans), False)
        print(f"  {tag}: tau {time.time() - t0:.0f}s", flush=True)
        out[f"srrl_{tag}"] = cols(jm.flux_rrl(line, chans, contsub=False), True)
        print(f"  {tag}: flux {time.time() - t0:.0f}s", flush=True)
        if tag in extras:
            out[f"srrl_cs_{tag}"] = cols(jm.flux_rrl(line, chans, contsub=True), True)
            # intensity_rrl's array branch is broken in the reference: per channel
            out[f"irrl_{tag}"] = cols(np.stack([jm.intensity_rrl(line, float(f))
                                               for f in chans]), True)
    path = os.path.join(ROOT, "tests", "golden", f"big_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {time.time() - t0:.0f}s -> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(cases.BIG_CASES)):
        run_case(nm)
