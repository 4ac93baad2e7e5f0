"""Sharded (x-slab) runs on >= 2 GPUs reproduce the single-GPU products bit for bit
(each ray is integrated by exactly one rank with the same arithmetic; the exchange is a
pure all-gather).  Skipped on boxes with one GPU; the gather logic itself is covered on
CPU by tests/test_sharding_gloo.py."""
import os
import socket
import tempfile

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import scipy.constants as con
    import torch
    import torch.distributed as dist
    import rajepy_b200 as rb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), f"m{rank}.log"), verbose=False)
    p = cases.with_grid(cases.base_params(), 64, 96, 128)
    nu0 = rb.hostmath.rrl_nu_0('H', 58, 1)
    chans = cases.line_channels(nu0, 96, 2e5)
    freqs = np.array([5e9, 4.3e10])
    res = {}
    for tag, shard in (("one", (0, 1)), ("sharded", (rank, world))):
        jm = rb.JetModel(p if tag == "one" else cases.with_grid(cases.base_params(), 64, 96, 128),
                         log=log, device=f"cuda:{rank}", shard=shard)
        jm.time = 0.9 * con.year
        res[tag] = (jm.n_verts_inside(), jm.emission_measure(), jm.flux_ff(freqs),
                    jm.optical_depth_rrl('H58a', chans),
                    jm.flux_rrl('H58a', chans, contsub=False),
                    jm.flux_rrl('H58a', chans[:8], contsub=True), jm.temperature)
    ok = all(np.array_equal(np.nan_to_num(a), np.nan_to_num(b)) and
             np.array_equal(np.isnan(a), np.isnan(b))
             for a, b in zip(res["one"], res["sharded"]))
    # products on rank 0's host only: the other ranks take part in the exchange and get None
    jm = rb.JetModel(cases.with_grid(cases.base_params(), 64, 96, 128), log=log,
                     device=f"cuda:{rank}", shard=(rank, world), host_ranks=(0,))
    jm.time = 0.9 * con.year
    em, cube = jm.emission_measure(), jm.flux_rrl('H58a', chans, contsub=False)
    if rank == 0:
        ok = ok and np.array_equal(em, res["one"][1]) and \
            np.array_equal(np.nan_to_num(cube), np.nan_to_num(res["one"][4]))
    else:
        ok = ok and em is None and cube is None
    # equal-width slabs give the same products as the work-balanced ones
    jm = rb.JetModel(cases.with_grid(cases.base_params(), 64, 96, 128), log=log,
                     device=f"cuda:{rank}", shard=(rank, world), balance=False)
    jm.time = 0.9 * con.year
    ok = ok and np.array_equal(jm.optical_depth_rrl('H58a', chans), res["one"][3])
    # epoch-sharded time series (BASELINE config 4): every rank gets the whole series
    epochs = np.linspace(0., 5., 7) * con.year
    series = rb.flux_ff_time_series(cases.with_grid(cases.base_params(), 64, 96, 128), epochs,
                                    5e9, rank=rank, world=world, device=f"cuda:{rank}", log=log)
    one = rb.JetModel(cases.with_grid(cases.base_params(), 64, 96, 128), log=log,
                      device=f"cuda:{rank}")
    for e in range(len(epochs)):
        one.time = float(epochs[e])
        ok = ok and np.array_equal(np.nan_to_num(series[e]), np.nan_to_num(one.flux_ff(5e9)))
    # channel sharding: every rank integrates a block of channels on the whole grid; the cube
    # is handed over through shared host memory (rank 0 gets it), the per-channel totals are
    # all-gathered.  Another thread layout than the unsharded kernel: fp32-level differences
    # in single evaluations (<= 2e-7), continuum products bit-identical.
    jm = rb.JetModel(cases.with_grid(cases.base_params(), 64, 96, 128), log=log,
                     device=f"cuda:{rank}", shard=(rank, world), shard_axis='channel',
                     host_ranks=(0,))
    jm.time = 0.9 * con.year
    em = jm.emission_measure()
    s_ff = jm.flux_ff(freqs)
    tau = jm.optical_depth_rrl('H58a', chans)
    cube = jm.flux_rrl('H58a', chans, contsub=False)
    tot = jm.rrl_flux_totals('H58a', chans, contsub=False)
    with np.errstate(all="ignore"):
        want_tot = np.nansum(res["one"][4].reshape(len(chans), -1), axis=1)
    ok = ok and np.allclose(tot, want_tot, rtol=1e-7, atol=0)
    if rank == 0:
        ok = ok and np.array_equal(em, res["one"][1])
        ok = ok and np.array_equal(np.nan_to_num(s_ff), np.nan_to_num(res["one"][2]))
        for got, want in ((tau, res["one"][3]), (cube, res["one"][4])):
            ok = ok and got.shape == want.shape
            ok = ok and np.array_equal(np.isnan(got), np.isnan(want))
            ok = ok and np.array_equal(got == 0, want == 0)
            m = ~np.isnan(want) & (want != 0)
            ok = ok and float(np.max(np.abs(got[m] / want[m] - 1.0))) < 5e-7
    else:
        ok = ok and tau is None and cube is None
    del tau, cube
    jm.release()
    # sky tiles: the cubes stay tiles on the device (bit-identical rows of the single-GPU cube);
    # the host cube is assembled by channel blocks after an all-to-all of the packed columns
    jm = rb.JetModel(cases.with_grid(cases.base_params(), 64, 96, 128), log=log,
                     device=f"cuda:{rank}", shard=(rank, world), shard_axis='tile',
                     host_ranks=(0,))
    jm.time = 0.9 * con.year
    lo, hi = jm.slab
    dev_res = jm.rt_products(freqs, 'H58a', chans, contsub=False, host=False)
    ok = ok and dev_res["rows"] == (lo, hi)
    tile = dev_res["flux_rrl"].view(len(chans), hi - lo, jm.nz).cpu().numpy()
    ok = ok and np.array_equal(np.nan_to_num(tile), np.nan_to_num(res["one"][4][:, lo:hi]))
    ok = ok and tuple(dev_res["flux_ff"].shape) == (2, jm.nx, jm.nz)     # complete sky images
    ok = ok and np.array_equal(np.nan_to_num(dev_res["flux_ff"].cpu().numpy()),
                               np.nan_to_num(res["one"][2]))
    ok = ok and np.array_equal(dev_res["em"].cpu().numpy(), res["one"][1])
    tot = jm.rrl_flux_totals('H58a', chans, contsub=False)
    ok = ok and np.allclose(tot, want_tot, rtol=1e-12, atol=0)
    tau = jm.optical_depth_rrl('H58a', chans)
    cube = jm.flux_rrl('H58a', chans, contsub=False)
    if rank == 0:
        ok = ok and np.array_equal(tau, res["one"][3])
        ok = ok and np.array_equal(np.nan_to_num(cube), np.nan_to_num(res["one"][4]))
        ok = ok and np.array_equal(np.isnan(cube), np.isnan(res["one"][4]))
    else:
        ok = ok and tau is None and cube is None
    del tau, cube
    jm.release()
    open(os.path.join(out_dir, f"r{rank}.txt"), "w").write("ok" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    tmp = tempfile.mkdtemp()
    mp.spawn(_worker, args=(world, _free_port(), tmp), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(tmp, f"r{r}.txt")).read() == "ok"
