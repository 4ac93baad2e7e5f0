#!/usr/bin/env python
"""
bench.py -- headline benchmark of the RaJePy hot path on B200.

Workload (BASELINE.json configs[4], the configuration the metric's target is quoted on;
it fits one GPU): the example jet on a 1024^3 grid, c_size 0.5 au, epoch 1 yr (bursts
active) -> continuum images at 16 frequencies (1-300 GHz) + the 512-channel H58a cube.
One "step" = the whole hot path for one epoch: grid fill (K1+K2), ONE fused line-of-sight
pass (K3+K4+K5: EM, tau_ff kernel sums, mean T, tau_L and flux cubes) and the continuum
image epilogue.  Metric = dense cells x channels / step time (Gcell.channel/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 1024] [--nchan 512]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
  python bench.py --impl reference ...     the numpy restatement of the reference's own
                                           CPU path, timed on a bounded sample

N > 1: the grid is sharded by x-slabs (strong scaling: total work fixed); every rank
fills and integrates its slab and the sky tiles are all-gathered over NCCL inside the
timed region.  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  L2: every step writes 8.6 GB of cube output (and the grid
state lives in an 18 GB buffer), far more than the 126 MB L2, so nothing a step reads can
still be cached from the previous one; no explicit flush is needed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gcell.channel integrations/s (grid fill + continuum + RRL cube, dense cells x channels)"
UNIT = "Gcell.channel/s"


def workload(grid, nchan):
    from tests import cases
    import rajepy_b200.hostmath as hm
    params = cases.with_grid(cases.base_params(), grid, grid, grid)
    cont = np.logspace(9, np.log10(3e11), 16)
    nu0 = hm.rrl_nu_0('H', 58, 1)
    chans = nu0 + (np.arange(nchan) - (nchan - 1) / 2.) * 1e5
    return params, cont, 'H58a', chans


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler is
    started before the warm-up (nvidia-smi needs ~100 ms to come up) and only the samples whose
    timestamps fall between mark_start() and mark_stop() are kept."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = os.path.join(tempfile.mkdtemp(), "clocks.csv")
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.f = open(self.path, "wt")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "10"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    @staticmethod
    def _stamp(txt):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, smax, reasons, sm_all = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for ln in f:
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    clk, cmax = float(parts[1]), float(parts[2])
                except ValueError:
                    continue
                sm_all.append(clk)
                smax.append(cmax)
                ts = self._stamp(parts[0])
                inside = ts is not None and self.t0 is not None and self.t1 is not None and \
                    self.t0 - 0.005 <= ts <= self.t1 + 0.005
                if not inside:
                    continue
                sm.append(clk)
                for nm, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "samples_whole_run": len(sm_all),
                "sm_mhz_whole_run": statistics.median(sm_all) if sm_all else None,
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arm
def oracle_step(grid, n_cont, n_line):
    """One pass of the reference algorithm (numpy restatement, oracle/) on a bounded
    sample of the workload: same jet, same cell size, grid^3 cells, n_cont continuum
    frequencies and n_line line channels.  Returns (seconds, cell.channel units)."""
    from oracle import rajepy_oracle as orc
    import scipy.constants as con
    params, cont, line, chans = workload(grid, 512)
    cont = cont[:: max(1, len(cont) // n_cont)][:n_cont]
    mid = len(chans) // 2
    chans = chans[mid - n_line // 2: mid - n_line // 2 + n_line]
    t0 = time.perf_counter()
    oj = orc.OracleJet(params, time_s=1.0 * con.year)
    oj.fill_factor()
    oj.emission_measure()
    oj.optical_depth_ff(cont)
    oj.flux_ff(cont)
    oj.optical_depth_rrl(line, chans)
    oj.flux_rrl(line, chans, contsub=False)
    dt = time.perf_counter() - t0
    return dt, grid ** 3 * (len(cont) + len(chans))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    grid, n_cont, n_line = 64, 16, 8
    for _ in range(min(args.warmup, 1)):
        oracle_step(grid, n_cont, n_line)
    times, units = [], 0
    for _ in range(args.steps):
        dt, units = oracle_step(grid, n_cont, n_line)
        times.append(dt)
    tot = sum(times)
    value = units * len(times) / tot / 1e9
    sample = (f"{grid}^3 cells of the same jet (c_size 0.5 au), {n_cont} continuum "
              f"frequencies + {n_line} H58a channels per step, numpy oracle "
              f"(oracle/rajepy_oracle.py), single-threaded like the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def bench_config(args):
    return {"workload": f"BASELINE configs[4]: example jet, {args.grid}^3 grid, c_size 0.5 au, "
                        f"epoch 1 yr, 16 continuum freqs 1-300 GHz + {args.nchan}-channel "
                        f"H58a cube (chan 100 kHz), contsub=False",
            "grid": [args.grid] * 3, "n_continuum": 16, "n_channels": args.nchan,
            "sharding": "work-balanced x-slabs, sparse cube exchange" if args.gpus > 1 else "none",
            "fill": "sparse: state buffers are recycled between models together with their "
                    "per-brick occupancy map, so a fill rewrites only the bricks around the jet "
                    "(a dense fill of fresh memory takes 2.75 ms at 1024^3, the first of a "
                    "process also a one-off zero fill; both are in the warm-up)",
            "l2": "each step streams 8.6 GB of cube output through L2 (126 MB) between reuses "
                  "of any input; no explicit flush needed"}


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import scipy.constants as con
    import rajepy_b200 as rb
    from rajepy_b200 import jetmodel as jmod

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_fd = None
    if world > 1:
        # stdout carries ONE JSON line: NCCL prints its version banner / debug lines to fd 1
        # from C, so fd 1 is pointed at stderr for the run and the line goes to the saved fd
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    params, cont, line, chans = workload(args.grid, args.nchan)
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), f"bench{rank}.log"), verbose=False)
    ncell = args.grid ** 3
    nchan_total = len(cont) + len(chans)
    units = ncell * nchan_total

    def make_model(host_ranks=None):
        import copy
        jm = rb.JetModel(copy.deepcopy(params), log=log, device=dev, shard=(rank, world),
                         host_ranks=host_ranks)
        jm.time = 1.0 * con.year
        return jm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kernel_ms = []

    def device_step():
        """HBM-resident step: fill + fused sweep + epilogue (+ all-gather of tiles)."""
        jm = make_model()
        jm._ensure_filled(sync=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        jm._pass(line, chans, contsub=False)
        e1.record()
        out = jm.rt_products(cont, line, chans, contsub=False, host=False)
        kernel_ms.append((e0, e1))
        jm.release()
        return out

    def e2e_step():
        """Through the public JetModel API with host (numpy) results."""
        # N > 1: the products land on rank 0's host (the rank that writes the FITS files);
        # the other ranks take part in the exchange only
        jm = make_model(host_ranks=(0,) if world > 1 else None)
        s_ff = jm.flux_ff(cont)
        t_l = jm.optical_depth_rrl(line, chans)
        s_l = jm.flux_rrl(line, chans, contsub=False)
        jm.release()
        if s_l is None:
            return 0, 0.0
        return s_ff.nbytes + t_l.nbytes + s_l.nbytes, float(np.nansum(s_l[len(chans) // 2]))

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        out = device_step()
        del out
    kernel_ms.clear()
    launches0 = jmod.LAUNCHES["count"]
    barrier()
    if sampler:
        sampler.mark_start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = device_step()
        del out
    ev1.record()
    barrier()
    if sampler:
        sampler.mark_stop()
    clocks = sampler.stop() if sampler else None
    launches = jmod.LAUNCHES["count"] - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kernel_ms) / len(kernel_ms)],
                       device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    value = units / (ms_per_step * 1e-3) / 1e9

    # end-to-end through the public API (host results), fewer repetitions: it is slow
    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d2h, checksum = e2e_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev,
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units / float(e2e_s) / 1e9

    if rank == 0:
        hbm, peak_src = peaks()
        nxs = args.grid // world
        alg_bytes = (ncell // world) * 16 + 2 * len(chans) * nxs * args.grid * 8 + \
            nxs * args.grid * 28
        kernel_s = float(kms) * 1e-3
        achieved = alg_bytes / kernel_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"pass@{args.grid}x{args.nchan}")
        line_out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(args),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": "JetModel(params) -> flux_ff(16 freqs), optical_depth_rrl, "
                            "flux_rrl(contsub=False) returned as numpy arrays (N > 1: on rank "
                            "0, host_ranks=(0,)); inputs are the parameter dict (no bulk H2D "
                            "exists on this path); bound by the device->host copy of the cubes",
                    "checksum_jy": checksum},
            "roofline": {"bound": "hbm",
                         "kernel": "integration pass: integrate_line_kernel (K3+K4+K5 ray walk) "
                                   "|| missed_rays_kernel (constant cube planes), two streams",
                         "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic,
                         "peak_source": peak_src, "kernel_ms": float(kms),
                         "algorithmic_bytes": alg_bytes,
                         "note": "algorithmic bytes per SURVEY 8(d) = 16 B per DENSE cell + tau and "
                                 "flux cubes + 4 sky images; the pass itself walks only the "
                                 "per-ray in-jet extents recorded by the fill (0.4 % of the "
                                 "cells), so its real DRAM traffic (`traffic`) is the 8.6 GB of "
                                 "cube output and its time is set by instruction issue in the "
                                 "channel loop (2.2e9 Voigt evaluations); see DESIGN.md "
                                 "section 4 and profiles/README.md"},
        }
        if world == 1 and not args.no_cpu_baseline:
            dt, u = oracle_step(128, 16, 8)
            line_out["cpu_baseline"] = {
                "value": u / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "128^3 cells of the same jet, 16 continuum freqs + 8 H58a "
                          "channels, numpy oracle (single-threaded like the reference), "
                          f"{dt:.1f} s"}
        if json_fd is not None:
            os.write(json_fd, (json.dumps(line_out) + "\n").encode())
        else:
            print(json.dumps(line_out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--nchan", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
