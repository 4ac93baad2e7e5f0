"""Where the time of the batched time series (BASELINE configs[3]: 512^3, 64 epochs, 5 GHz) goes:
device time by CUDA events and host wall clock of `flux_ff_time_series`, best of n.
    python tools/epochs_probe.py [grid] [epochs]"""
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
from tests import cases  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    ne = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "e.log"), verbose=False)
    p = cases.with_grid(cases.base_params(), n, n, n)
    epochs = np.linspace(0., 5., ne) * con.year
    for it in range(6):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        s = rb.flux_ff_time_series(copy.deepcopy(p), epochs, 5e9, log=log, host=False)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"run {it}: device {a.elapsed_time(b):7.3f} ms   host (launch side) "
              f"{(t1 - t0) * 1e3:7.3f} ms   checksum {float(torch.nansum(s[-1])):.12e}",
              flush=True)
    # the same epochs one pass at a time (round 1's driver)
    jm = rb.JetModel(copy.deepcopy(p), log=log)
    for it in range(2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in epochs:
            jm.time = float(t)
            last = jm._continuum_images_device(5e9, 'flux')
        b.record()
        torch.cuda.synchronize()
        print(f"per-epoch passes {it}: device {a.elapsed_time(b):7.3f} ms   checksum "
              f"{float(torch.nansum(last)):.12e}", flush=True)


if __name__ == "__main__":
    main()
