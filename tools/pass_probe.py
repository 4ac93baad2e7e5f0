"""Kernel-level probe of the integration pass at the bench size (not the bench): the pass
with different constant-writer configurations, the ray kernels alone, the writer alone.
python tools/pass_probe.py [grid] [nchan]   (knobs are env vars read by the launcher)"""
import copy
import os
import sys
import tempfile

import scipy.constants as con
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from rajepy_b200 import _cabi  # noqa: E402
from bench import workload  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    grid = int(pos[0]) if len(pos) > 0 else 1024
    nchan = int(pos[1]) if len(pos) > 1 else 512
    params, cont, line, chans = workload(grid, nchan)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)
    jm = rb.JetModel(copy.deepcopy(params), log=log)
    jm.time = 1.0 * con.year
    d = jm._ensure_filled()
    n_act = jm._n_active()
    ext = d["extents"]
    missed = int((ext[:, 0] >= ext[:, 1]).sum())
    print(f"grid {grid}^3, {nchan} channels, {n_act} jet-crossing rays, {missed} missed", flush=True)

    def run_pass():
        jm._line = None
        jm._cont = None
        jm._pass(line, chans, contsub=False)

    def setenv(**kw):
        for k in ("RJP_PIPE_CHUNKS", "RJP_CHAN_BLOCK", "RJP_GRID_FACTOR", "RJP_WRITER_CTAS", "RJP_WRITER_WARPS", "RJP_NO_BULK", "RJP_WRITER_PER_SM", "RJP_SKIP_WRITER", "RJP_SKIP_LINES",
                  "RJP_LINE_THREADS", "RJP_FUSE_WRITER"):
            os.environ.pop(k, None)
        for k, v in kw.items():
            os.environ[k] = str(v)

    configs = [("pass (line kernel || writer 1/SM)", {}),
               ("ray kernels alone", {"RJP_SKIP_WRITER": 1}),
               ("writer alone (1/SM, bulk)", {"RJP_SKIP_LINES": 1}),
               ("writer alone (2/SM, bulk)", {"RJP_SKIP_LINES": 1, "RJP_WRITER_PER_SM": 2}),
               ("writer alone (4/SM, bulk)", {"RJP_SKIP_LINES": 1, "RJP_WRITER_PER_SM": 4}),
               ("writer alone (4/SM, scalar)", {"RJP_SKIP_LINES": 1, "RJP_WRITER_PER_SM": 4,
                                                "RJP_NO_BULK": 1}),
               ("pass, scalar writer", {"RJP_NO_BULK": 1}),
               ("pass, writer 2/SM", {"RJP_WRITER_PER_SM": 2})]
    if "--quick" in sys.argv:
        configs = configs[:2]
    if "--blocks" in sys.argv:
        configs = [("alone, 512 ch per launch", {"RJP_SKIP_WRITER": 1}),
                   ("alone, 256 ch per launch (1-warp CTAs)", {"RJP_SKIP_WRITER": 1,
                                                               "RJP_CHAN_BLOCK": 256}),
                   ("pass,  256 ch per launch", {"RJP_CHAN_BLOCK": 256})]
    if "--writer" in sys.argv:
        configs = [("ray kernels alone", {"RJP_SKIP_WRITER": 1})]
        for ctas, warps in ((148, 1), (74, 2), (37, 4), (37, 8), (24, 8),
                            (148, 4)):
            e = {"RJP_WRITER_CTAS": ctas, "RJP_WRITER_WARPS": warps}
            configs.append((f"writer alone {ctas}x{warps}", dict(e, RJP_SKIP_LINES=1)))
            configs.append((f"pass, writer {ctas}x{warps}", e))
    if "--factors" in sys.argv:
        configs = []
        for f in (1, 4, 16, 64):
            configs.append((f"alone, grid factor {f}", {"RJP_SKIP_WRITER": 1,
                                                        "RJP_GRID_FACTOR": f}))
            configs.append((f"pass,  grid factor {f}", {"RJP_GRID_FACTOR": f}))
    gb = missed * nchan * 16 / 1e9
    for name, env in configs:
        setenv(**env)
        best, med = timed(run_pass)
        note = f"  ({gb / best * 1e3:.0f} GB/s of constants)" if "alone (" in name else ""
        print(f"{name:32s} best {best:7.3f} ms  median {med:7.3f} ms{note}", flush=True)
    setenv()
    jm.release()


if __name__ == "__main__":
    main()
