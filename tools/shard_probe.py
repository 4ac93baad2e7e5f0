"""Where a sharded line pass spends its time (run under torchrun on >= 2 GPUs):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_probe.py
CUDA-event timing of the phases of JetModel._pass on every rank."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rajepy_b200 as rb  # noqa: E402
from rajepy_b200 import jetmodel as jmod, sharding  # noqa: E402
from bench import workload  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    params, cont, line, chans = workload(1024, 512)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "p.log"), verbose=False)
    jm = rb.JetModel(params, log=log, device=dev, shard=(rank, world))
    jm.time = 31536000.0
    jm._ensure_filled()
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    # wrap the phases
    orig_fill, orig_exch = jm._fill_remote_constants, jm._exchange_cubes
    orig_pack = sharding.exchange_ray_columns

    side_marks = []

    def fill(*a):
        mark("start")
        r = orig_fill(*a)
        e = torch.cuda.Event(enable_timing=True)
        e.record(r)                      # end of the remote constant fills on the side stream
        side_marks.append(e)
        return r

    def exch(tau, flux, side):
        mark("integrate launched+done(stream)")
        orig_exch(tau, flux, side)
        mark("exchange done")

    jm._fill_remote_constants, jm._exchange_cubes = fill, exch
    for it in range(4):
        marks.clear()
        side_marks.clear()
        jm._line = None
        dist.barrier()
        torch.cuda.synchronize()
        jm._pass(line, chans, contsub=False)
        torch.cuda.synchronize()
    t0 = marks[0][1]
    txt = f"rank {rank}: n_active={jm._n_active()} " + ", ".join(
        f"{n} @ {t0.elapsed_time(e):.2f} ms" for n, e in marks)
    txt += f", side stream fills done @ {t0.elapsed_time(side_marks[0]):.2f} ms, slab {jm.slab}"
    print(txt, file=sys.stderr, flush=True)
    # exchange alone
    tau, flux = jm._line["tau"], jm._line["flux"]
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig_exch(tau, flux, jm._dev["stream3"])
        b.record()
        torch.cuda.synchronize()
    print(f"rank {rank}: exchange alone {a.elapsed_time(b):.2f} ms", file=sys.stderr, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
