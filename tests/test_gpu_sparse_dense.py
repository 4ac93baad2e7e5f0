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
                           d["cells"].data_ptr(), None, None, 0, em.data_ptr(), kff.data_ptr(),
                           tsum.data_ptr(), cnt.data_ptr(), None, None, 0, 1, None, None,
                           jm._stream(), None)
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
