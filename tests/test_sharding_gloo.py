"""The N > 1 path on CPU: two gloo ranks each hold an x-slab tile of sky images / cubes
(cut from an oracle result) and must reassemble the full product with the all-gather
the GPU path uses (rajepy_b200.sharding.gather_x)."""
import os
import socket
import tempfile

import numpy as np
import pytest

from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, img_path, out_dir):
    import torch
    import torch.distributed as dist
    from rajepy_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(img_path)
    lo, hi = sharding.slab_bounds(nx, rank, world)
    img = torch.from_numpy(d["img"][lo:hi].copy())
    cube = torch.from_numpy(d["cube"][:, lo:hi].copy())
    cnt = torch.from_numpy(d["cnt"][lo:hi].copy())
    full_img = sharding.gather_x(img, nx, rank, world, dim=0)
    full_cube = sharding.gather_x(cube, nx, rank, world, dim=1)
    full_cnt = sharding.gather_x(cnt, nx, rank, world, dim=0)
    host_cube = full_cube.to_host() if hasattr(full_cube, "to_host") else full_cube
    assert tuple(full_cube.shape) == (cube.shape[0], nx, cube.shape[2])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), img=full_img.numpy(),
             cube=host_cube.numpy(), cnt=full_cnt.numpy(),
             cube0=full_cube[0].numpy() if hasattr(full_cube, "to_host") else full_cube[0].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nx_case", ["even", "uneven"])
def test_gather_x_world2(nx_case):
    import torch.multiprocessing as mp
    from oracle import rajepy_oracle as orc
    p = cases.case_small() if nx_case == "even" else cases.with_grid(cases.base_params(),
                                                                   18, 24, 30)
    if nx_case == "uneven":
        p["grid"]["n_x"] = 18
    oj = orc.OracleJet(p)
    world = 2 if nx_case == "even" else 4   # 18 planes over 4 ranks: 5,5,4,4
    em = oj.emission_measure()
    tau = oj.optical_depth_ff(np.array([5e9, 2e10, 1e11]))
    cnt = (oj.n_verts_inside() > 0).sum(axis=1).astype(np.int32)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "in.npz")
    np.savez(path, img=em, cube=tau, cnt=cnt)
    mp.spawn(_worker, args=(world, _free_port(), oj.nx, path, tmp), nprocs=world, join=True)
    for r in range(world):
        out = np.load(os.path.join(tmp, f"r{r}.npz"))
        assert np.array_equal(out["img"], em)
        assert np.array_equal(out["cube"], tau)
        assert np.array_equal(out["cube0"], tau[0])
        assert np.array_equal(out["cnt"], cnt)


def test_sharded_model_slabs_are_consistent():
    """Every rank's JetModel agrees on the slab decomposition and the slabs tile the grid."""
    import rajepy_b200 as rb
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    slabs = [rb.JetModel(cases.case_c1(), log=log, shard=(r, 8)).slab for r in range(8)]
    assert slabs[0][0] == 0 and slabs[-1][1] == 50
    assert all(slabs[i][1] == slabs[i + 1][0] for i in range(7))


# ------------------------------------------------------------------ sparse cube exchange
def _exchange_worker(rank, world, port, nx, nz, path, out_dir, bounds):
    import torch
    import torch.distributed as dist
    from rajepy_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(path)
    lo, hi = bounds[rank] if bounds else sharding.slab_bounds(nx, rank, world)
    nch = d["tau"].shape[0]
    # what a rank holds after its own pass: full-size cubes whose rows of the OWN slab are
    # final, everything else uninitialised (poisoned here)
    cubes = []
    for name in ("tau", "flux"):
        c = np.full((nch, nx, nz), 123.456)
        c[:, lo:hi] = d[name][:, lo:hi]
        cubes.append(torch.from_numpy(c).view(nch, nx * nz))
    ext = torch.from_numpy(d["extents"][lo * nz: hi * nz].copy())
    local_rays = torch.from_numpy(np.flatnonzero(d["extents"][lo * nz: hi * nz, 0] <
                                                 d["extents"][lo * nz: hi * nz, 1])
                                  .astype(np.int32))
    local_rays = local_rays[torch.randperm(local_rays.numel())]   # the GPU list is unordered
    meta = sharding.build_ray_meta(ext, local_rays, lo, nx, nz, rank, world, bounds=bounds)
    assert sum(meta["counts"]) == int((d["extents"][:, 0] < d["extents"][:, 1]).sum())
    sharding.exchange_ray_columns(cubes, [0.0, float("nan")], meta, nx, nz, rank, world,
                                  bounds=bounds)
    # tau only (flux not requested)
    only = torch.from_numpy(np.where(np.arange(nx)[None, :, None] // 1 >= 0, 7.0, 7.0) *
                            np.ones((nch, nx, nz))).view(nch, nx * nz)
    only.view(nch, nx, nz)[:, lo:hi] = torch.from_numpy(d["tau"][:, lo:hi])
    sharding.exchange_ray_columns([only, None], [0.0, float("nan")], meta, nx, nz, rank, world,
                                  bounds=bounds)
    np.savez(os.path.join(out_dir, f"x{rank}.npz"), tau=cubes[0].view(nch, nx, nz).numpy(),
             flux=cubes[1].view(nch, nx, nz).numpy(), only=only.view(nch, nx, nz).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bounds", [(2, None), (4, None),
                                          (4, [(0, 7), (7, 9), (9, 10), (10, 18)])])
def test_sparse_cube_exchange(world, bounds):
    """Only the cube columns of jet-crossing rays travel; every rank rebuilds the constants
    of the other slabs from the all-gathered extents.  The result must be the full oracle
    cube on every rank (18 planes over 4 ranks: uneven slabs; work-balanced slabs of very
    different widths, one of them a single plane)."""
    import torch.multiprocessing as mp
    from oracle import rajepy_oracle as orc
    p = cases.with_grid(cases.base_params(), 18, 24, 30)
    oj = orc.OracleJet(p)
    nu0 = orc.rrl_nu_0('H', 58, 1)
    chans = cases.line_channels(nu0, 5, 1e6)
    tau = oj.optical_depth_rrl('H58a', chans)
    flux = oj.flux_rrl('H58a', chans, contsub=False)
    inside = oj.n_verts_inside() > 0                       # (nx, ny, nz)
    any_in = inside.any(axis=1)
    first = np.where(any_in, inside.argmax(axis=1), 2147483647)
    last = np.where(any_in, inside.shape[1] - inside[:, ::-1].argmax(axis=1), 0)
    extents = np.stack([first, last], axis=-1).reshape(-1, 2).astype(np.int32)
    assert np.array_equal(np.isnan(flux[0]).reshape(-1), extents[:, 0] >= extents[:, 1])
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "in.npz")
    np.savez(path, tau=tau, flux=flux, extents=extents)
    mp.spawn(_exchange_worker, args=(world, _free_port(), oj.nx, oj.nz, path, tmp, bounds),
             nprocs=world, join=True)
    for r in range(world):
        out = np.load(os.path.join(tmp, f"x{r}.npz"))
        assert np.array_equal(out["tau"], tau)
        assert np.array_equal(np.isnan(out["flux"]), np.isnan(flux))
        assert np.array_equal(np.nan_to_num(out["flux"]), np.nan_to_num(flux))
        assert np.array_equal(out["only"], tau)


# ------------------------------------------------------------------ epoch sharding
def _epoch_worker(rank, world, port, n_epochs, out_dir):
    import torch
    import torch.distributed as dist
    from rajepy_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.epoch_shares(n_epochs, rank, world)
    # "image" of epoch e: constant e + pixel ramp
    local = torch.stack([torch.arange(6, dtype=torch.float64) + 100.0 * e for e in mine]) \
        if mine else torch.empty((0, 6), dtype=torch.float64)
    full = sharding.gather_epochs(local, n_epochs, rank, world)
    np.save(os.path.join(out_dir, f"e{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_epochs", [(2, 7), (4, 64), (4, 3)])
def test_epoch_sharded_series_is_reassembled_in_order(world, n_epochs):
    import torch.multiprocessing as mp
    tmp = tempfile.mkdtemp()
    mp.spawn(_epoch_worker, args=(world, _free_port(), n_epochs, tmp), nprocs=world, join=True)
    want = np.arange(6)[None, :] + 100.0 * np.arange(n_epochs)[:, None]
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(tmp, f"e{r}.npy")), want)


# ------------------------------------------------------------------ channel sharding
def _chan_worker(rank, world, port, nchan, out_dir):
    """Every rank owns chan_bounds(nchan, rank, world) of a synthetic cube: the all-gather of
    the per-channel totals and the shared-memory hand-over (hostshare, without CUDA: the
    segment is only page-locked on a GPU box) must reassemble the whole product."""
    import torch
    import torch.distributed as dist
    from rajepy_b200 import hostshare, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plane = 6 * 10
    cube = np.arange(nchan * plane, dtype=np.float64).reshape(nchan, plane) * 0.5
    cube[:, ::7] = np.nan
    lo, hi = sharding.chan_bounds(nchan, rank, world)
    local = torch.from_numpy(cube[lo:hi].copy())
    tot = sharding.gather_channel_totals(torch.nansum(local, dim=1), nchan, rank, world)
    ok = np.array_equal(tot.numpy(), np.nansum(cube, axis=1))
    for _ in range(2):                       # second round reuses the pooled segment
        seg = hostshare.segment(nchan * plane * 8, rank)
        seg._register = lambda off, nb: None  # no CUDA here
        mine = seg.tensor(lo * plane * 8, (hi - lo, plane))
        mine.copy_(local)
        dist.barrier()
        if rank == 0:
            arr = seg.array((nchan, 6, 10))
            ok = ok and np.array_equal(np.nan_to_num(arr.reshape(nchan, plane)),
                                       np.nan_to_num(cube))
            ok = ok and seg.busy
            del arr
            ok = ok and not seg.busy
        dist.barrier()
    open(os.path.join(out_dir, f"r{rank}.txt"), "w").write("ok" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nchan", [(2, 8), (3, 10), (4, 3)])
def test_channel_sharding_world(world, nchan):
    import torch.multiprocessing as mp
    from rajepy_b200 import sharding
    # the blocks tile [0, nchan) in order, sizes differ by at most one
    bounds = [sharding.chan_bounds(nchan, r, world) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == nchan
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in bounds]
    assert max(sizes) - min(sizes) <= 1
    tmp = tempfile.mkdtemp()
    mp.spawn(_chan_worker, args=(world, _free_port(), nchan, tmp), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(tmp, f"r{r}.txt")).read() == "ok"
