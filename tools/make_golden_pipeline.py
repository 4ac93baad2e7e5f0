"""
Pin the Pipeline-side glue (run list, product names, channel grids, flux totals, save files)
to the UNMODIFIED reference.  Build container only:

    python tools/make_golden_pipeline.py

Drives the reference's own `Pipeline.__init__` (classes.py:2045-2172) and `JetModel.save` /
`Pipeline.save` under oracle/ref_shim.py on a tiny model and writes
  tests/golden/pipeline_runs.json    per run: type, year, day, freq, line, bandwidth, chanwidth,
                                     nchan, chan_freqs, rt_dcy / fits_* relative to the pipeline
                                     directory, and results['flux'] as classes.py:2461-2472
                                     reduces the reference's own flux arrays
  tests/golden/ref_jetmodel.save     pickle written by the reference's JetModel.save (:1704-1713)
  tests/golden/ref_pipeline.save     pickle written by the reference's Pipeline.save (:2215-2258)
The pipeline parameters and the model are tests/cases.py:pipeline_case().
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import scipy.constants as con

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from tests import cases  # noqa: E402


def main():
    rjp = ref_shim.load_reference()
    model_params, pl_params = cases.pipeline_case()
    work = tempfile.mkdtemp(prefix="rjp_pl_")
    dcy = os.path.join(work, "pl")
    pl_params["dcys"]["model_dcy"] = dcy
    jm = ref_shim.make_reference_model(model_params, logfile=os.path.join(work, "jm.log"))
    pl = rjp.Pipeline(jm, pl_params, log=jm.log)
    runs = []
    for run in pl.runs:
        jm.time = run.year * con.year
        if run.obs_type == 'continuum':
            fluxes = jm.flux_ff(run.chan_freqs)
            flux = float(np.nansum(np.nanmean(fluxes, axis=0)))
        else:
            fluxes = jm.flux_rrl(run.line, run.chan_freqs, contsub=False)
            flux = [float(v) for v in np.nansum(np.nansum(fluxes, axis=1), axis=1)]
        run.results['flux'] = flux
        rel = lambda p: os.path.relpath(p, dcy)  # noqa: E731
        runs.append({"obs_type": run.obs_type, "year": float(run.year), "day": int(run.day),
                     "freq": float(run.freq), "line": getattr(run, "line", None),
                     "bandwidth": float(run.bandwidth), "chanwidth": float(run.chanwidth),
                     "nchan": int(run.nchan), "chan_freqs": [float(f) for f in run.chan_freqs],
                     "rt_dcy": rel(run.rt_dcy), "fits_flux": rel(run.fits_flux),
                     "fits_tau": rel(run.fits_tau), "fits_em": rel(run.fits_em),
                     "radiative_transfer": bool(run.radiative_transfer),
                     "simobserve": bool(run.simobserve), "flux": flux})
    gold = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gold, "pipeline_runs.json"), "wt") as f:
        json.dump({"runs": runs, "model_file": os.path.relpath(pl.model_file, dcy),
                   "save_file": os.path.relpath(pl.save_file, dcy)}, f, indent=1)
    jm.fill_factor  # noqa: B018 -- so that the save file carries ffs / areas
    jm.time = 0.75 * con.year
    jm.save(os.path.join(work, "jetmodel.save"))
    shutil.copy(os.path.join(work, "jetmodel.save"), os.path.join(gold, "ref_jetmodel.save"))
    pl.save(os.path.join(work, "pipeline.save"))
    shutil.copy(os.path.join(work, "pipeline.save"), os.path.join(gold, "ref_pipeline.save"))
    for nm in ("pipeline_runs.json", "ref_jetmodel.save", "ref_pipeline.save"):
        print(nm, os.path.getsize(os.path.join(gold, nm)), "bytes")


if __name__ == "__main__":
    main()
