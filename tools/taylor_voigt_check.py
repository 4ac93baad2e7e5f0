"""Accuracy check (numpy + scipy.special.wofz) of the Taylor-propagated Voigt profile used by
rjp::voigt_taylor (rjp_device.cuh): worst relative error of Re w(xc + h + iy) for offsets |h| <= H
and expansion order M, with and without the acceptance rule of the kernel."""
import numpy as np
from scipy.special import wofz
def taylor_re(xc, y, h, M):
    z = xc + 1j*y
    c0 = wofz(z)
    cs = [c0, -2*z*c0 + 2j/np.sqrt(np.pi)]
    for n in range(1, M):
        cs.append(-2*(z*cs[n] + cs[n-1])/(n+1))
    a = np.array([c.real for c in cs])   # shape (M+1, ...)
    s = np.zeros_like(h*xc)
    for n in range(M, -1, -1):
        s = s*h + a[n]
    t_last = np.abs(np.array(cs[M]))*np.abs(h)**M
    t_prev = np.abs(np.array(cs[M-1]))*np.abs(h)**(M-1)
    return s, np.maximum(t_last, t_prev), c0
xc = np.linspace(0, 30, 1201)
for M in (12, 16, 20):
  for H in (0.1, 0.25, 0.5):
    print(f"M={M} H={H}")
    for y in (1e-3, 1e-2, 0.1, 1.0, 5.0):
        worst = 0; worst_acc = 0; nacc = 0
        for h in (H, -H, 0.6*H):
            s, tail, c0 = taylor_re(xc, y, h, M)
            ref = wofz(xc + h + 1j*y).real
            rel = np.abs(s-ref)/ref
            # acceptance
            acc = (tail <= 1e-10*np.abs(c0.real)) & (2*np.abs(xc)*H <= 12)
            worst = max(worst, rel.max())
            if acc.any(): worst_acc = max(worst_acc, rel[acc].max())
            nacc = acc.mean()
        print(f"   y={y:6.3f} worst rel (all) {worst:.2e}  worst among accepted {worst_acc:.2e}  accepted frac {nacc:.2f}")
