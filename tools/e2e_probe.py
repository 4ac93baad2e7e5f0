"""Where does the end-to-end (host result) time go?  python tools/e2e_probe.py [grid] [nchan]"""
import os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb
from bench import workload
import scipy.constants as con

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nchan = int(sys.argv[2]) if len(sys.argv) > 2 else 512
params, cont, line, chans = workload(grid, nchan)
log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)

def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label:40s} {1e3*(time.perf_counter()-t0):9.1f} ms", flush=True); return r

for rep in range(2):
    print("--- rep", rep)
    import copy
    jm = t("JetModel()", lambda: rb.JetModel(copy.deepcopy(params), log=log))
    jm.time = con.year
    t("fill (incl. ties, ray list)", jm._ensure_filled)
    t("flux_ff(16)", lambda: jm.flux_ff(cont))
    res = t("line pass (device)", lambda: jm._pass(line, chans, contsub=False))
    x = t("pinned alloc 4.3GB", lambda: torch.empty(res["tau"].shape, dtype=torch.float64, pin_memory=True))
    t("D2H into pinned", lambda: x.copy_(res["tau"]))
    t("numpy copy of pinned", lambda: x.numpy().copy())
    pg = t("pageable alloc+D2H (.cpu())", lambda: res["tau"].cpu())
    del x, pg
    t("optical_depth_rrl (API)", lambda: jm.optical_depth_rrl(line, chans))
    t("flux_rrl (API)", lambda: jm.flux_rrl(line, chans, contsub=False))
    jm.release()
