"""Hand-over of one dense cube to the host at the bench size: wall time of JetModel._host_cube
for fixed splits between host threads (constants + packed columns) and the copy engine
(whole planes), and for several thread counts.  python tools/e2e_probe.py"""
import copy
import os
import sys
import tempfile
import time

import numpy as np
import scipy.constants as con
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from rajepy_b200 import jetmodel as jmod  # noqa: E402
from bench import workload  # noqa: E402


def main():
    params, cont, line, chans = workload(1024, 512)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)
    jm = rb.JetModel(copy.deepcopy(params), log=log)
    jm.time = 1.0 * con.year
    res = jm._pass(line, chans, contsub=False)
    torch.cuda.synchronize()
    nch = len(chans)
    ref = jm._host_image(res["flux"], lead=nch)
    # one e2e step, phase by phase, like bench.py's e2e_step
    os.environ.pop("RAJEPY_B200_HOST_SPLIT", None)
    for rep in range(4):
        t = [time.perf_counter()]
        j2 = rb.JetModel(copy.deepcopy(params), log=log)
        j2.time = 1.0 * con.year
        s_ff = j2.flux_ff(cont)
        t.append(time.perf_counter())
        r2 = j2._pass(line, chans, contsub=False)
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        pin = torch.empty((nch, 1024 * 1024), dtype=torch.float64, pin_memory=True)
        t.append(time.perf_counter())
        del pin
        tau = j2._host_cube(r2["tau"], 0.0, nch)
        t.append(time.perf_counter())
        fl = j2._host_cube(r2["flux"], float("nan"), nch)
        t.append(time.perf_counter())
        j2.release()
        del tau, fl, s_ff, r2
        t.append(time.perf_counter())
        print("step phases [ms]: model+flux_ff %.1f, pass %.1f, pinned alloc %.1f, tau cube %.1f, "
              "flux cube %.1f, release %.1f  rates %s" % (*[1e3 * (b - a) for a, b in
                                                           zip(t[:-1], t[1:])], jmod._HANDOVER),
              flush=True)
    if "--steps-only" in sys.argv:
        return
    for threads in (16, 8, 32):
        os.environ["RAJEPY_B200_HOST_THREADS"] = str(threads)
        for split in (0.0, 0.3, 0.5, 0.7, 0.85, 1.0):
            os.environ["RAJEPY_B200_HOST_SPLIT"] = str(split)
            ts = []
            for _ in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = jm._host_cube(res["flux"], float("nan"), nch)
                ts.append(time.perf_counter() - t0)
                ok = out.shape == ref.shape
                del out
            out = jm._host_cube(res["flux"], float("nan"), nch)
            same = np.array_equal(np.isnan(out), np.isnan(ref)) and \
                np.array_equal(np.nan_to_num(out), np.nan_to_num(ref))
            del out
            print(f"threads {threads:2d} split {split:4.2f}: best {min(ts[1:]) * 1e3:7.1f} ms "
                  f"(first {ts[0] * 1e3:7.1f})  {4.295 / min(ts[1:]):6.1f} GB/s  identical={same} "
                  f"rates {jmod._HANDOVER}", flush=True)
    jm.release()


if __name__ == "__main__":
    main()
