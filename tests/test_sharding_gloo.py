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
