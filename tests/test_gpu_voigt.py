"""The Voigt profile function of the channel loop evaluated ON THE DEVICE
(rjp_voigt_profile, the kernels' own routines) against scipy.special.wofz, the routine the
reference calls (maths/rrls.py:353).  Tolerances: 3e-7 per evaluation for the mixed
fp64/fp32 class (y <= 0.1), 5e-8 for the fp64 class; the bar on line-of-sight sums is 1e-6."""
import numpy as np
import pytest
from scipy.special import wofz

pytestmark = pytest.mark.gpu


def _device_voigt(x, y):
    import torch
    from rajepy_b200 import _cabi
    lib = _cabi.load()
    xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    yd = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64), device="cuda")
    out = torch.empty_like(xd)
    _cabi.check(lib.rjp_voigt_profile(xd.data_ptr(), yd.data_ptr(), xd.numel(), out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "voigt")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_device_voigt_fast_class():
    xs = np.linspace(-60.0, 60.0, 60001)
    for y in (1e-9, 1e-6, 1e-4, 4.5e-3, 0.011, 0.03, 0.06, 0.1):
        got = _device_voigt(xs, np.full_like(xs, y))
        rel = np.abs(got / wofz(xs + 1j * y).real - 1.0)
        assert rel.max() < 3e-7, (y, rel.max(), xs[rel.argmax()])
        assert np.sqrt(np.mean(rel ** 2)) < 8e-8, y


def test_device_voigt_matches_numpy_emulation():
    """The numpy emulation used by the CPU tests is the arithmetic the device performs."""
    from tests import voigt_emul as ve
    xs = np.linspace(-30.0, 30.0, 20001)
    for y in (1e-4, 0.011, 0.1):
        got = _device_voigt(xs, np.full_like(xs, y))
        emu = ve.voigt_fast(xs, y)
        assert np.abs(got / emu - 1.0).max() < 2.5e-7


def test_device_voigt_fp64_class():
    xs = np.linspace(-45.0, 45.0, 9001)
    for y in (0.11, 0.38, 1.0, 7.5, 40.0):
        got = _device_voigt(xs, np.full_like(xs, y))
        rel = np.abs(got / wofz(xs + 1j * y).real - 1.0)
        assert rel.max() < 5e-8, (y, rel.max())


def test_device_voigt_random_points():
    rng = np.random.default_rng(7)
    x = rng.uniform(-50, 50, 200000)
    y = 10.0 ** rng.uniform(-6, 1.5, 200000)
    got = _device_voigt(x, y)
    rel = np.abs(got / wofz(x + 1j * y).real - 1.0)
    assert rel.max() < 3e-7, (rel.max(), x[rel.argmax()], y[rel.argmax()])
