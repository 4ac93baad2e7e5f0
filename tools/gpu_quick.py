"""Quick on-GPU timing probe (not the bench): fill + continuum pass + line pass at a
few grid sizes, CUDA-event timed.  python tools/gpu_quick.py 256 512 [--nchan 256]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from tests import cases  # noqa: E402


def timed(fn, n=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


SPARSE = "--dense-fill" not in sys.argv
TWO_LEVEL = "--one-level-fill" not in sys.argv


def main():
    argv = sys.argv[1:]
    nchan = 256
    if "--nchan" in argv:
        i = argv.index("--nchan")
        nchan = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    sizes = [int(a) for a in args] or [256]
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "q.log"), verbose=False)
    for n in sizes:
        p = cases.with_grid(cases.base_params(), n, n, n)
        jm = rb.JetModel(p, log=log)
        t0 = time.time()
        d = jm._ensure_filled()
        torch.cuda.synchronize()
        wall_fill = time.time() - t0
        ncell = n ** 3
        injet = int((d["nverts"] > 0).sum())
        lib = rb._cabi.load()

        work = torch.empty(d["bricks"].numel() + 4, dtype=torch.int32, device="cuda")

        def refill():
            cnt = torch.zeros(8, dtype=torch.int32, device="cuda")
            ties = torch.empty((1 << 16, 4), dtype=torch.int32, device="cuda")
            lib.rjp_fill_grid(d["model"], d["nverts"].data_ptr(), d["cells"].data_ptr(),
                              d["bricks"].data_ptr() if SPARSE else None,
                              work.data_ptr() if (SPARSE and TWO_LEVEL) else None,
                              ties.data_ptr(), 1 << 16, cnt.data_ptr(),
                              d["extents"].data_ptr(), jm._stream())
        t_fill = timed(refill)

        def cont():
            jm._cont = None
            jm._pass()
        t_cont = timed(cont, 5)
        nu0 = rb.hostmath.rrl_nu_0('H', 58, 1)
        chans = cases.line_channels(nu0, nchan, 1e5)

        def line():
            jm._line = None
            jm._pass('H58a', chans, contsub=False)
        t_line = timed(line, 3)
        if hasattr(lib, "rjp_debug_stamps"):
            import ctypes
            buf = (ctypes.c_ulonglong * 4)()
            lib.rjp_debug_stamps(buf)
            line()
            lib.rjp_debug_stamps(buf)
            t0 = min(buf[0], buf[2])
            print(f"  stamps [us]: line kernel {(buf[0] - t0) / 1e3:.0f}..{(buf[1] - t0) / 1e3:.0f}, "
                  f"missed kernel {(buf[2] - t0) / 1e3:.0f}..{(buf[3] - t0) / 1e3:.0f}")
        gb = ncell * 16 / 1e9
        print(f"n={n} cells={ncell:.3e} in-jet={injet} ({100 * injet / ncell:.2f}%) "
              f"ties={d['n_ties']} patched={d['n_patched']} first-fill wall {wall_fill:.2f}s\n"
              f"  fill      {t_fill:8.3f} ms  ({ncell * 17 / 1e6 / t_fill:7.1f} GB/s written)\n"
              f"  continuum {t_cont:8.3f} ms  ({gb / t_cont * 1e3:7.1f} GB/s)\n"
              f"  line x{nchan} {t_line:8.3f} ms  ({gb / t_line * 1e3:7.1f} GB/s, "
              f"{ncell * nchan / t_line / 1e6:.1f} Gcell.ch/s, "
              f"{injet * nchan / t_line / 1e6:.2f} G in-jet evals/s)", flush=True)
        jm.release()
        del jm, d
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
