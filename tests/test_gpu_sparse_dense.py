"""The ray walk over the fill's per-ray extents (what JetModel uses) must give the same
continuum sums as the dense sweep that reads every cell of the state (rjp_integrate with
extents = NULL): this checks the extents / ray list the fill records against the state."""
import os
import tempfile

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small", "inclined", "powerlaws", "c1"])
def test_sparse_walk_equals_dense_sweep(name):
    import torch
    import rajepy_b200 as rb
    from rajepy_b200 import _cabi
    lib = _cabi.load()
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    jm = rb.JetModel(cases.CASES[name][0](), log=log)
    jm.time = 1.1 * cases.YEAR
    sparse = jm._pass()
    d = jm._ensure_filled()
    npix = sparse["em"].numel()
    em, kff, tsum = (torch.empty(npix, dtype=torch.float64, device="cuda") for _ in range(3))
    cnt = torch.empty(npix, dtype=torch.int32, device="cuda")
    st = lib.rjp_integrate(d["model"], jm._epoch_struct(), jm._continuum_struct(),
                           d["cells"].data_ptr(), None, None, None, 0, em.data_ptr(),
                           kff.data_ptr(),
                           tsum.data_ptr(), cnt.data_ptr(), None, None, 0, 1, None, None,
                           0, 0, None, None, None, 0, jm._stream(), None)
    _cabi.check(st, "rjp_integrate(dense)")
    torch.cuda.synchronize()
    assert torch.equal(cnt, sparse["cnt"])
    for a, b, what in ((em, sparse["em"], "em"), (kff, sparse["kff"], "kff"),
                       (tsum, sparse["tsum"], "tsum")):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        assert np.array_equal(a == 0, b == 0), what
        nz = a != 0
        assert np.abs(a[nz] / b[nz] - 1.0).max() < 1e-13, what
    # a line pass must leave the same continuum images behind as the continuum-only walk
    nu0 = rb.hostmath.rrl_nu_0('H', 58, 1)
    jm._pass('H58a', cases.line_channels(nu0, 16, 2e5), contsub=False)
    for k in ("em", "kff", "tsum"):
        a, b = jm._cont[k].cpu().numpy(), sparse[k].cpu().numpy()
        assert np.array_equal(a == 0, b == 0), k
        nz = a != 0
        assert np.abs(a[nz] / b[nz] - 1.0).max() < 1e-13, k
    assert torch.equal(jm._cont["cnt"], sparse["cnt"])


def test_recycled_state_equals_dense_fill():
    """The sparse fill on a recycled buffer (occupancy map of a DIFFERENT previous jet) must
    leave exactly the state a dense fill writes into fresh memory."""
    import copy
    import torch
    import rajepy_b200 as rb
    from rajepy_b200 import _cabi, jetmodel
    lib = _cabi.load()
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    jetmodel.clear_state_pool()
    pa = cases.with_grid(cases.base_params(), 72, 80, 96)
    pb = copy.deepcopy(pa)
    pb["geometry"].update({"inc": 55., "pa": 40., "opang": 40., "w_0": 2.0})
    first = rb.JetModel(pa, log=log)
    first._ensure_filled()
    first.release()                       # buffers + occupancy map go to the pool
    assert len(jetmodel._STATE_POOL) == 1
    second = rb.JetModel(pb, log=log)
    d = second._ensure_filled()           # recycled: zeroes stale bricks, writes new ones
    assert len(jetmodel._STATE_POOL) == 0
    ncell = d["nverts"].numel()
    nv = torch.full((ncell,), 0xAB, dtype=torch.uint8, device="cuda")
    cl = torch.full((ncell, 2), float("nan"), dtype=torch.float64, device="cuda")
    ext = torch.empty_like(d["extents"])
    ties = torch.empty((1 << 16, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(8, dtype=torch.int32, device="cuda")
    st = lib.rjp_fill_grid(d["model"], nv.data_ptr(), cl.data_ptr(), None, None, ties.data_ptr(),
                           1 << 16, cnt.data_ptr(), ext.data_ptr(), second._stream())
    _cabi.check(st, "rjp_fill_grid(dense)")
    torch.cuda.synchronize()
    if d["n_patched"] == 0:
        assert torch.equal(nv, d["nverts"])
        assert torch.equal(cl.view(torch.int64), d["cells"].view(torch.int64))
        assert torch.equal(ext, d["extents"])
    # occupancy map is consistent with the data: bricks marked empty hold only zeros
    bricks = d["bricks"].cpu().numpy()
    nxs, ny, nz = second._x_hi - second._x_lo, second._ny, second._nz
    occ = (d["nverts"].view(nxs, ny, nz) != 0).cpu().numpy()
    tx, ty, tz = 4, 8, 32
    bx, by, bz = -(-nxs // tx), -(-ny // ty), -(-nz // tz)
    pad = np.zeros((bx * tx, by * ty, bz * tz), dtype=bool)
    pad[:nxs, :ny, :nz] = occ
    per_brick = pad.reshape(bx, tx, by, ty, bz, tz).any(axis=(1, 3, 5)).reshape(-1)
    assert bricks.size == per_brick.size
    assert not np.any(per_brick & (bricks == 0))
    second.release()
    jetmodel.clear_state_pool()


def test_state_pool_is_keyed_on_the_layout():
    """Two models with the same number of cells and bricks but different (ny, nz): the flat
    cell index and the brick ids depend on the layout, so the second model must not reuse the
    first one's occupancy map."""
    import copy
    import rajepy_b200 as rb
    from rajepy_b200 import jetmodel
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    jetmodel.clear_state_pool()
    pa = cases.with_grid(cases.base_params(), 32, 64, 128)
    pb = cases.with_grid(cases.base_params(), 32, 128, 64)
    first = rb.JetModel(pa, log=log)
    first._ensure_filled()
    first.release()
    second = rb.JetModel(copy.deepcopy(pb), log=log)
    got = second.n_verts_inside()
    second.release()
    jetmodel.clear_state_pool()
    fresh = rb.JetModel(copy.deepcopy(pb), log=log)
    want = fresh.n_verts_inside()
    assert np.array_equal(got, want)
    assert np.array_equal(np.nan_to_num(fresh.emission_measure()),
                          np.nan_to_num(rb.JetModel(copy.deepcopy(pb), log=log).emission_measure()))
    jetmodel.clear_state_pool()


def test_ray_list_is_ordered_and_complete():
    """rjp_ray_list: ascending slab-local ids of exactly the rays with a non-empty extent."""
    import torch
    import rajepy_b200 as rb
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    for name in ("small", "inclined", "c1"):
        jm = rb.JetModel(cases.CASES[name][0](), log=log)
        d = jm._ensure_filled()
        n = jm._n_active()
        ext = d["extents"].cpu().numpy()
        want = np.flatnonzero(ext[:, 0] < ext[:, 1])
        got = d["rays"][:n].cpu().numpy()
        assert n == want.size
        assert np.array_equal(got, want)
        jm.release()
