"""Per-slab cost of a tile-sharded step, measured on ONE GPU: for every rank r of a world of N the
slab model (shard=(r, N), shard_axis='tile') is filled and its line pass timed with CUDA events
(no collectives are involved in either).  Shows what the busiest rank of an N-GPU run pays.
    python tools/slab_probe.py [N ...]          (default 2 4 8)
RJP_WRITER_CTAS / RJP_SKIP_WRITER / RJP_SKIP_LINES apply (debug knobs of rjp_integrate)."""
import copy
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from bench import workload  # noqa: E402


def main():
    worlds = [int(a) for a in sys.argv[1:]] or [2, 4, 8]
    dev = torch.device("cuda", 0)
    params, cont, line, chans = workload(1024, 512)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "s.log"), verbose=False)
    for world in worlds:
        worst = 0.0
        for r in range(world):
            best_f = best_p = 1e30
            for it in range(4):
                jm = rb.JetModel(copy.deepcopy(params), log=log, device=dev, shard=(r, world),
                                 shard_axis="tile")
                jm.time = 31536000.0
                torch.cuda.synchronize()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                jm._ensure_filled(sync=False)
                e[1].record()
                jm._pass(line, chans, contsub=False)
                e[2].record()
                torch.cuda.synchronize()
                if it > 0:
                    best_f = min(best_f, e[0].elapsed_time(e[1]))
                    best_p = min(best_p, e[1].elapsed_time(e[2]))
                n_act, cells = jm._ray_counts()
                slab = jm.slab
                jm.release()
            worst = max(worst, best_f + best_p)
            print(f"N={world} rank {r}: slab {slab} ({slab[1] - slab[0]:4d} planes)  rays "
                  f"{n_act:6d}  in-extent cells {cells:8d}  fill {best_f:6.3f} ms  pass "
                  f"{best_p:6.3f} ms", flush=True)
        print(f"N={world}: busiest slab fill + pass = {worst:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
