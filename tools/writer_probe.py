"""Bandwidth of the constant writer (rjp_fill_missed) alone, heavy and light grids.
python tools/writer_probe.py"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from rajepy_b200 import _cabi  # noqa: E402
from tests import cases  # noqa: E402


def main():
    lib = _cabi.load()
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "w.log"), verbose=False)
    jm = rb.JetModel(cases.with_grid(cases.base_params(), 1024, 1024, 1024), log=log)
    d = jm._ensure_filled()
    nray, nch = 1024 * 1024, 512
    tau = torch.empty((nch, nray), dtype=torch.float64, device="cuda")
    flux = torch.empty((nch, nray), dtype=torch.float64, device="cuda")
    missed = int((d["extents"][:, 0] >= d["extents"][:, 1]).sum())
    gb = missed * nch * 16 / 1e9
    for light in (0, 1):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            st = lib.rjp_fill_missed(d["extents"].data_ptr(), nray, nch, nray, 0, 0, 0,
                                     tau.data_ptr(), flux.data_ptr(), light,
                                     torch.cuda.current_stream().cuda_stream)
            b.record()
            torch.cuda.synchronize()
            _cabi.check(st, "fill_missed")
            best = min(best, a.elapsed_time(b))
        print(f"light={light}: {best:.3f} ms for {gb:.2f} GB -> {gb / best * 1e3:.0f} GB/s")
    # reference point: torch fill of the same bytes
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tau.zero_()
        flux.fill_(float("nan"))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"torch zero_/fill_ of both cubes: {best:.3f} ms -> {nch * nray * 16 / 1e9 / best * 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
