"""
Host-side (fp64, numpy/scipy) scalar physics that feeds the CUDA kernels: everything in
the reference's hot path that is O(1) or O(n_freq) rather than O(cells).  Constants come
from `scipy.constants` at run time, like the reference (never hard-coded CODATA values).

Names and argument meaning follow the reference modules so that reference-side callers
can switch imports: maths/geometry.py (mod_r_0, rho, w_r, r_eff, xyz_rotate, xyz_to_rwp),
maths/physics.py (q_n, q_tau, n_0_from_mlr, atomic_mass, z_number, rydberg_constant,
doppler_shift, blackbody_nu, gff) and maths/rrls.py (rrl_nu_0, energy_n, f_n1n2,
ni_from_ne, deltanu_l, deltanu_g, rrl_parser).
"""
import functools
import json
import os

import numpy as np
import scipy.constants as con

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

MSOL = 1.98847E30  # kg (_constants.py:5)
AU2CM = con.au * 1e2
NZ = {"H": (1, 0), "He": (2, 2), "Li": (3, 4), "Be": (4, 5), "B": (5, 6), "C": (6, 6),
      "N": (7, 7), "O": (8, 8), "F": (9, 10), "Ne": (10, 10), "Na": (11, 12),
      "Mg": (12, 12)}

c_cgs = con.c * 1e2
h_cgs = con.h * 1e7
k_cgs = con.k * 1e7


# ----------------------------------------------------------------- geometry (scalars)
def mod_r_0(opang, epsilon, w_0):
    """geometry.py:12-31"""
    return epsilon * w_0 / np.tan(np.radians(opang) / 2.)


def rho(r, r_0, mr0=None):
    """geometry.py:34-61"""
    if mr0:
        return (np.abs(r) + mr0 - r_0) / mr0
    return np.abs(r) / r_0


def w_r(r, w_0, mr0, r_0, eps):
    """geometry.py:96-118"""
    return w_0 * rho(r, r_0, mr0) ** eps


def r_eff(w, r_1, r_2, w_0, r, mr0, r_0, eps):
    """geometry.py:305-336"""
    return r_1 + ((r_2 - r_1) * w) / w_r(r, w_0, mr0, r_0, eps)


def rotation_trig(alpha_deg, beta_deg):
    """cos/sin of the two rotation angles exactly as geometry.py:243-247 forms them."""
    a = np.radians(alpha_deg)
    b = np.radians(beta_deg)
    return float(np.cos(a)), float(np.sin(a)), float(np.cos(b)), float(np.sin(b))


def xyz_rotate(x, y, z, alpha, beta, order='xy'):
    """geometry.py:212-263"""
    ca, sa, cb, sb = rotation_trig(alpha, beta)

    def x_rot(x_, y_, z_):
        return x_, ca * y_ - sa * z_, sa * y_ + ca * z_

    def y_rot(x_, y_, z_):
        return cb * x_ + sb * z_, y_, cb * z_ - sb * x_

    if order.lower() == 'xy':
        return y_rot(*x_rot(x, y, z))
    if order.lower() == 'yx':
        return x_rot(*y_rot(x, y, z))
    raise ValueError(f"Order of rotation, {order.__repr__()}, not recognised")


def xyz_to_rwp(x, y, z, inc, pa):
    """geometry.py:181-209 + :266-302 -> (r, w, phi)"""
    x1, y2, r = xyz_rotate(x, y, z, inc - 90., pa, order='yx')
    w = np.sqrt(x1 ** 2. + y2 ** 2.)
    with np.errstate(all='ignore'):
        p = np.arcsin(y2 / w)
    if not np.isscalar(x1):
        p = np.where(x1 < 0, -p + np.pi, p)
    elif x1 < 0:
        p = -p + np.pi
    return r, w, p


def lz_to_grid_dims(params):
    """classes.py:90-122"""
    cs_au = params["grid"]["c_size"]
    g = params["geometry"]
    i_rads = np.radians(g["inc"])
    pa_rads = np.radians(g["pa"])
    l_xz_au = params['grid']['l_z'] * params['target']['dist']
    xmax_au = l_xz_au * np.sin(pa_rads)
    ymax_au = l_xz_au * np.tan(1.571 - i_rads)
    zmax_au = l_xz_au * np.cos(pa_rads)
    rmax_au = xyz_to_rwp(xmax_au, ymax_au, zmax_au, g["inc"], g["pa"])[0]
    wmax_au = w_r(rmax_au, g["w_0"], g["mod_r_0"], g["r_0"], g["epsilon"])
    wmax_cells = int(np.ceil(np.abs(wmax_au / cs_au)))
    dims = [int(np.ceil(np.abs(v / cs_au))) + 2 * wmax_cells
            for v in (xmax_au, ymax_au, zmax_au)]
    return tuple(d if d % 2 == 0 else d + 1 for d in dims)


# ----------------------------------------------------------------- physics scalars
def q_n(epsilon, q_v):
    """physics.py:17-35"""
    return -q_v - (2.0 * epsilon)


def q_tau(epsilon, q_x, q_n_, q_T):
    """physics.py:38-63"""
    return epsilon + 2.0 * q_x + 2.0 * q_n_ - 1.35 * q_T


@functools.lru_cache(maxsize=None)
def _mass_table():
    with open(os.path.join(_DATA, "atomic_masses.json"), "rt") as f:
        return json.load(f)


def atomic_mass(atom):
    """kg (physics.py:607-624)"""
    m = _mass_table()[atom]["mass_micro_u"]
    m *= 1e-6 * con.u
    return m


def z_number(atom):
    """physics.py:523-532"""
    return {'H': 1, 'He': 2, 'Li': 3, 'Be': 4, 'B': 5, 'C': 6, 'N': 7, 'O': 8}[atom]


def rydberg_constant(atom):
    """m^-1 (physics.py:535-544)"""
    m_atom = atomic_mass(atom)
    return con.Rydberg * (m_atom / (m_atom + con.m_e))


def doppler_shift(nu_0, v_lsr):
    """physics.py:547-558"""
    return nu_0 * (1. - v_lsr * 1000. / con.c)


def blackbody_nu(freq, temp):
    """erg s^-1 cm^-2 Hz^-1 sr^-1 (physics.py:561-574)"""
    p1 = 2. * con.h * 1e7 * freq ** 3. / (con.c * 1e2) ** 2.
    p2 = np.exp(con.h * 1e7 * freq / (con.k * 1e7 * temp)) - 1.
    return p1 * p2 ** -1.


def n_0_from_mlr(mlr, v_0, w_0, mu, q_nd, q_nv, R_1, R_2):
    """cm^-3 (physics.py:474-517)"""
    a = q_nd + q_nv
    if a == -1. or a == -2.:
        a *= 1. + 1e-12
    r2 = R_2 * con.au
    r1 = R_1 * con.au
    mlr_si = mlr * MSOL / con.year
    constant = 2. * con.pi * (mu * atomic_mass('H')) * (v_0 * 1e3) * (w_0 * con.au) ** 2.
    return mlr_si / constant / \
        ((r1 ** 2. + r2 * (r2 * (a + 1.) - r1 * (a + 2.)) * (r2 / r1) ** a) /
         ((r2 - r1) ** 2. * (a + 1.) * (a + 2.))) / 1e6


# ----------------------------------------------------------------- Gaunt factor
@functools.lru_cache(maxsize=None)
def _gaunt_table():
    """van Hoof et al. (2014) table and its axes as physics.py:626-663 builds them."""
    d = np.load(os.path.join(_DATA, "gaunt_vanhoof2014.npz"))
    g = np.array(d["gff"])
    n_u, n_g = g.shape
    step = float(d["step"])
    u0, g0 = float(d["log_u_start"]), float(d["log_gamma2_start"])
    lus = np.linspace(np.round(u0, decimals=1), np.round(u0 + step * (n_u - 1), decimals=1),
                      n_u)
    lgs = np.linspace(np.round(g0, decimals=1), np.round(g0 + step * (n_g - 1), decimals=1),
                      n_g)
    lg2, lu2 = np.meshgrid(lgs, lus)
    return lg2, lu2, g


def gff(freq, temp, z=1.):
    """Free-free Gaunt factor(s) of van Hoof et al. (2014) at frequency(ies) `freq` and a
    scalar temperature (physics.py:666-698): nearest table node, 5x5 patch around it
    (row clamp uses the column count, physics.py:687-690), bicubic FITPACK surface
    (`interp2d(kind='cubic')` on scattered input == bisplrep(kx=ky=3, s=0)).  Channels
    that share a patch share one spline fit."""
    from scipy.interpolate import bisplev, bisplrep
    scalar = np.isscalar(freq)
    freqs = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    key = (freqs.tobytes(), float(temp), float(z))
    hit = _GFF_CACHE.get(key)
    if hit is not None:
        return float(hit[0]) if scalar else hit.copy()
    ry = con.m_e * con.e ** 4. / (8 * con.epsilon_0 ** 2. * con.h ** 2.)
    logg2 = float(np.log10(z ** 2. * ry / (con.k * temp)))
    logus = np.log10(con.h * freqs / (con.k * temp))
    lg2s, lus, g = _gaunt_table()
    ncol = len(lg2s[0])
    col = int(np.argmin(np.abs(lg2s[0] - logg2)))
    col = min(max(col, 2), ncol - 3)
    rows = np.argmin(np.abs(lus[:, 0][None, :] - logus[:, None]), axis=1)
    rows = np.clip(rows, 2, ncol - 3)
    out = np.empty(freqs.shape)
    for row in np.unique(rows):
        sl = (slice(row - 2, row + 3), slice(col - 2, col + 3))
        tck = bisplrep(lg2s[sl].ravel(), lus[sl].ravel(), g[sl].ravel(), kx=3, ky=3, s=0.0)
        for i in np.flatnonzero(rows == row):
            out[i] = np.ravel(bisplev(np.atleast_1d(logg2), np.atleast_1d(logus[i]),
                                      tck))[0]
    if len(_GFF_CACHE) > 256:
        _GFF_CACHE.clear()
    _GFF_CACHE[key] = out.copy()
    return float(out[0]) if scalar else out


_GFF_CACHE = {}


# ----------------------------------------------------------------- recombination lines
def rrl_parser(rrl_str):
    """'H58a' -> ('H', 58, 1)  (rrls.py:605-624)"""
    dn = {'a': 1, 'b': 2, 'g': 3, 'd': 4}[rrl_str[-1].lower()]
    element, n = '', ''
    for char in rrl_str[:-1]:
        if char.isalpha():
            element += char
        else:
            n += char
    return element, int(n), dn


def rrl_nu_0(atom, n, delta_n=1):
    """Hz (rrls.py:14-29)"""
    return rydberg_constant(atom) * con.c * z_number(atom) ** 2. * \
        (1. / n ** 2. - 1. / (n + delta_n) ** 2.)


def energy_n(n, atom):
    """erg (rrls.py:32-41)"""
    return -2.17989724e-11 * z_number(atom) ** 2. / n ** 2.


def f_n1n2(n_1, delta_n):
    """rrls.py:44-59"""
    m_deltan = {1: 0.190775, 2: 0.026332, 3: 0.0081056, 4: 0.0034918}[delta_n]
    return n_1 * m_deltan * (1. + 1.5 * delta_n / n_1)


def ni_from_ne(n_e, atom='H'):
    """rrls.py:62-83"""
    xyz = {'H': 0.710, 'He': 0.276, 'CNO': 0.014}
    mu = (xyz['H'] / atomic_mass("H") * con.u + xyz['He'] / atomic_mass("He") * con.u +
          xyz['CNO'] / 14.24) ** -1.
    m_atom = atomic_mass(atom) / con.u
    return xyz[atom] * n_e * mu / m_atom


def deltanu_l(n_e, n, delta_n, gamma=4.5):
    """Hz (rrls.py:86-101)"""
    return 8.2 * n_e * (n / 100.) ** gamma * (1. + gamma / 2. * delta_n / n)


def deltanu_g(nu_0, temp, atom):
    """Hz (rrls.py:104-118)"""
    m = atomic_mass(atom)
    return np.sqrt(4. * np.log(2.) * 2. * con.k * temp / (m * con.c ** 2.)) * nu_0


def chan_freqs(freq, bandwidth, chanwidth):
    """ContinuumRun.chan_freqs (classes.py:1893-1900)"""
    nchan = int(bandwidth / chanwidth)
    chan1 = freq - bandwidth / 2. + chanwidth / 2.
    return chan1 + np.arange(nchan) * chanwidth


# ----------------------------------------------------------------- Reynolds (1986) cross-check
A_K = 0.212      # _constants.py:13
A_J = 6.5e-38    # _constants.py:14


def flux_expected_r86(jm, freq, which, y_max, y_min=None):
    """Exact flux [Jy] of one lobe of the jet from equation 8 of Reynolds (1986), between the
    angular distances y_min and y_max [arcsec] from the jet base (maths/physics.py:297-374).
    `jm` is a JetModel (only its parameters and steady-state mass-loss rates are read): the
    analytic sanity check the reference overlays on its SED plots
    (plotting/functions.py:1198-1199)."""
    from mpmath import gammainc
    p = jm.params
    inc = p['geometry']['inc']
    w_0 = p['geometry']['w_0'] * con.au * 1e2
    t_0 = p['properties']['T_0']
    n_0 = p['properties']['n_0']
    if which == 'R':
        n_0 *= jm.ss_jml('R') / jm.ss_jml('B')
    x_0 = p['properties']['x_0']
    q_tau_ = p["power_laws"]["q_tau"]
    q_t = p["power_laws"]["q_T"]
    eps = p["geometry"]["epsilon"]
    mod_r_0 = p['geometry']['mod_r_0'] * con.au * 1e2
    mod_y_0 = mod_r_0 * np.sin(np.radians(inc))
    r_0 = p['geometry']['r_0'] * con.au * 1e2
    y_0 = r_0 * np.sin(np.radians(inc))
    d = p['target']['dist'] * con.parsec * 1e2
    if p["power_laws"]["q^d_n"] != 0.:
        mlr = p["properties"]["mlr"] * 1.989e30 / con.year
        n_0 = mlr / (np.pi * p['properties']['mu'] * atomic_mass("H") * w_0 ** 2. *
                     p['properties']["v_0"] * 1e5)
    y_max = np.tan(y_max * con.arcsec) * d + mod_y_0 - y_0
    if y_min is not None:
        y_min = np.tan(y_min * con.arcsec) * d + mod_y_0 - y_0
    else:
        y_min = mod_y_0
    tau_0 = 2. * A_K * w_0 * (n_0 * x_0) ** 2. * t_0 ** -1.35 * freq ** -2.1 * \
        np.sin(np.radians(inc)) ** -1.
    c = 1. + eps + q_t

    def indef_integral(yval):
        const = 2. * w_0 * d ** -2. * A_J * A_K ** -1. * t_0 * freq ** 2.
        rho_ = yval / mod_y_0
        tau = tau_0 * rho_ ** q_tau_
        p1 = yval / (q_tau_ * c) * rho_ ** (c - 1.) * tau ** (-c / q_tau_)
        p2 = q_tau_ * tau ** (c / q_tau_) + c * gammainc(c / q_tau_, tau)
        return const * (float(p1) * float(p2))

    flux = indef_integral(y_max) - indef_integral(y_min)
    flux *= 1e-7 * 1e2 ** 2.
    return flux / 1e-26
