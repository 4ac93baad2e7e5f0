"""Device times of the five BASELINE.json configurations on one GPU (CUDA events, best of 3,
results left on the device):  python tools/config_times.py"""
import copy
import os
import sys
import tempfile

import numpy as np
import scipy.constants as con
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from tests import cases  # noqa: E402


def timed(fn, n=3):
    best = 1e30
    for _ in range(n):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "c.log"), verbose=False)
    nu0 = rb.hostmath.rrl_nu_0('H', 58, 1)
    freqs16 = np.logspace(9, np.log10(3e11), 16)

    def model(n, shape=None):
        p = cases.base_params() if shape is None else \
            cases.with_grid(cases.base_params(), *shape)
        jm = rb.JetModel(copy.deepcopy(p), log=log)
        return jm

    def c1():
        jm = model(0)
        jm.rt_products(np.array([5e9]), host=False)
        jm.release()

    def c2():
        jm = model(0, (256, 256, 256))
        jm.rt_products(freqs16, host=False)
        jm.release()

    def c3():
        jm = model(0, (512, 512, 512))
        jm.rt_products(None, 'H58a', cases.line_channels(nu0, 256, 1e5), contsub=False,
                       host=False)
        jm.release()

    series = {}

    def c4():
        jm = model(0, (512, 512, 512))
        for t in np.linspace(0., 5., 64):
            jm.time = t * con.year
            series[t] = jm.rt_products(np.array([5e9]), host=False)["flux_ff"]
        jm.release()

    def c4b():
        p = cases.with_grid(cases.base_params(), 512, 512, 512)
        series["batched"] = rb.flux_ff_time_series(p, np.linspace(0., 5., 64) * con.year, 5e9,
                                                   log=log, host=False)

    def c5():
        jm = model(0, (1024, 1024, 1024))
        jm.time = con.year
        jm.rt_products(freqs16, 'H58a', cases.line_channels(nu0, 512, 1e5), contsub=False,
                       host=False)
        jm.release()

    rows = (("C1 50x400x50, 5 GHz continuum (fill + pass)", c1, 50 * 400 * 50 * 1),
            ("C2 256^3, 16 continuum frequencies", c2, 256 ** 3 * 16),
            ("C3 512^3, 256-channel H58a cube", c3, 512 ** 3 * 256),
            ("C4 512^3, 64 epochs x 5 GHz continuum (1 fill + 64 passes)", c4, 512 ** 3 * 64),
            ("C4 batched: one ray walk for the 64 epochs (flux_ff_time_series)", c4b,
             512 ** 3 * 64),
            ("C5 1024^3, 16 continuum + 512-channel cube", c5, 1024 ** 3 * 528))
    for name, fn, units in rows:
        fn()
        ms = timed(fn)
        print(f"{name:62s} {ms:9.3f} ms   {units / ms / 1e6:12.1f} Gcell.channel/s", flush=True)


if __name__ == "__main__":
    main()
