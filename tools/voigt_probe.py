"""Error of the device Voigt routine against wofz (max / rms per y), for the library selected
by RAJEPY_B200_LIB."""
import os
import sys

import numpy as np
from scipy.special import wofz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_voigt import _device_voigt  # noqa: E402

xs = np.linspace(-60.0, 60.0, 240001)
for y in (1e-9, 1e-6, 1e-4, 4.5e-3, 0.011, 0.03, 0.06, 0.1):
    got = _device_voigt(xs, np.full_like(xs, y))
    ref = wofz(xs + 1j * y).real
    rel = np.abs(got / ref - 1.0)
    core = np.abs(xs) < 5.0
    print(f"y={y:8.1e} max {rel.max():.2e} at x={xs[rel.argmax()]:+.3f}  rms {np.sqrt(np.mean(rel**2)):.2e} "
          f" core: max {rel[core].max():.2e} rms {np.sqrt(np.mean(rel[core]**2)):.2e} "
          f"mean signed {np.mean((got / ref - 1.0)[core]):+.2e}")
