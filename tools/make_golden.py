"""
Generate tests/golden/<case>.npz by executing the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on the parity cases of
tests/cases.py.  Run in the build container only:  python tools/make_golden.py [case ...]

Stored per case (float64 unless noted):
  dims (3,), nverts uint8 (nx,ny,nz)           -- classes.py:657-666 count
  jet_idx int64 flat indices of cells with nverts>0, and for those cells:
     ts0 (launch time at model time 0 = -travel time), nd_base, xi, temp, vx, vlos, vz, reff
  per epoch e: em_e, tauff_e (nf,nx,nz), iff_e, sff_e, taurrl_e (nch,nx,nz),
               srrl_e (nch,nx,nz; contsub=False), srrl_cs_e (contsub=True)
  scalars: n_0, mod_r_0, nu0, gff per continuum freq
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
    factory, epochs, freqs, line, nch, chanw = cases.CASES[name]
    rjp = ref_shim.load_reference()
    jm = ref_shim.make_reference_model(factory())
    out = {"dims": np.array([jm.nx, jm.ny, jm.nz])}
    t0 = time.time()
    ff = jm.fill_factor
    # recompute the integer count exactly as classes.py:657-666 does (not kept by the ref)
    g = jm.params["geometry"]
    cs = jm.csize
    nv = np.zeros(ff.shape, dtype=int)
    for dx, dy, dz in ((0., 0., 0.), (cs, 0., 0.), (0., cs, 0.), (cs, cs, 0.),
                       (0., 0., cs), (cs, 0., cs), (0., cs, cs), (cs, cs, cs)):
        rv, wv = rjp.maths.geometry.xyz_to_rwp(jm.xx + dx, jm.yy + dy, jm.zz + dz,
                                               g["inc"], g["pa"])[:2]
        wrv = rjp.maths.geometry.w_r(rv, g["w_0"], g["mod_r_0"], g["r_0"], g["epsilon"])
        nv = np.where((wrv >= wv) & (np.abs(rv) >= g["r_0"]), nv + 1, nv)
    ff_from_nv = np.where(nv == 8, 1.0, np.where(nv > 0, 0.5, np.nan))
    assert np.array_equal(np.isnan(ff), np.isnan(ff_from_nv))
    assert np.array_equal(np.nan_to_num(ff), np.nan_to_num(ff_from_nv))
    out["nverts"] = nv.astype(np.uint8)
    idx = np.flatnonzero(nv.ravel() > 0)
    out["jet_idx"] = idx
    jm.time = 0.
    out["ts0"] = jm.ts.ravel()[idx]
    out["nd_base"] = jm.number_density.ravel()[idx] / jm.chi_xyz.ravel()[idx]
    out["xi"] = jm.ion_fraction.ravel()[idx]
    out["temp"] = jm.temperature.ravel()[idx]
    vx, vy, vz = jm.vel
    out["vx"], out["vlos"], out["vz"] = (v.ravel()[idx] for v in (vx, vy, vz))
    out["reff"] = jm.rreff.ravel()[idx]
    out["n_0"] = jm.params["properties"]["n_0"]
    out["mod_r_0"] = jm.params["geometry"]["mod_r_0"]
    el, n, dn = rjp.maths.rrls.rrl_parser(line)
    nu0 = rjp.maths.rrls.rrl_nu_0(el, n, dn)
    out["nu0"] = nu0
    chans = cases.line_channels(nu0, nch, chanw)
    out["chans"] = chans
    out["freqs"] = np.array(freqs)
    if jm.params["power_laws"]["q_T"] == 0.:
        out["gff"] = np.array([float(rjp.maths.physics.gff(f, jm.params["properties"]["T_0"]))
                               for f in freqs])
    for e, yr in enumerate(epochs):
        jm.time = yr * con.year
        out[f"chi_{e}"] = jm.chi_xyz.ravel()[idx]
        out[f"em_{e}"] = jm.emission_measure()
        out[f"tauff_{e}"] = jm.optical_depth_ff(np.array(freqs))
        out[f"iff_{e}"] = jm.intensity_ff(np.array(freqs))
        out[f"sff_{e}"] = jm.flux_ff(np.array(freqs))
        out[f"taurrl_{e}"] = jm.optical_depth_rrl(line, chans)
        out[f"srrl_{e}"] = jm.flux_rrl(line, chans, contsub=False)
        out[f"srrl_cs_{e}"] = jm.flux_rrl(line, chans, contsub=True)
    out["epochs_yr"] = np.array(epochs)
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: dims={tuple(out['dims'])} in-jet={idx.size} "
          f"({time.time() - t0:.1f}s) -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(cases.CASES)):
        run_case(nm)
