"""Host-side profile of one bench step (fill + line pass + images, HBM-resident results):
where the ~1 ms between kernels goes.  python tools/step_profile.py [steps]"""
import cProfile
import copy
import os
import pstats
import sys
import tempfile

import scipy.constants as con
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from bench import workload  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    params, cont, line, chans = workload(1024, 512)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)

    def step():
        jm = rb.JetModel(copy.deepcopy(params), log=log)
        jm.time = 1.0 * con.year
        jm._ensure_filled()
        jm._pass(line, chans, contsub=False)
        out = jm.rt_products(cont, line, chans, contsub=False, host=False)
        jm.release()
        return out

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pr = cProfile.Profile()
    a.record()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    b.record()
    torch.cuda.synchronize()
    print(f"{a.elapsed_time(b) / steps:.3f} ms/step (with profiler overhead)")
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
