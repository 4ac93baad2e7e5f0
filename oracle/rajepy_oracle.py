"""
TEST INFRASTRUCTURE ONLY -- the CPU oracle for RaJePy's hot path.

A plain numpy/scipy restatement of the reference algorithm (grid fill + line-of-sight
radiative transfer), written from the reference's *behaviour*; every function cites the
reference file:line it follows (paths relative to the reference root).  It exists to
CHECK the CUDA product (`rajepy_b200`), never to serve it: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.  The product package never imports anything from `oracle/`.

Pinning: `tests/golden/*.npz` hold outputs of the UNMODIFIED reference executed in the
build container through `oracle/ref_shim.py` (generator: `tools/make_golden.py`);
`tests/test_oracle_golden.py` checks this restatement against them (vertex counts
bit-exact, everything else <= 1e-12 relative), and -- when /root/reference is present --
`tests/test_oracle_vs_reference.py` re-runs the reference live.

Third-party arithmetic on the path (not under the reference tree): numpy ufuncs,
`scipy.special.hyp2f1`, `scipy.special.wofz`, FITPACK `bisplrep/bisplev` (what the
removed `scipy.interpolate.interp2d` used for scattered input), `scipy.constants`
(CODATA values are taken from scipy at run time, never hard-coded).

Conventions (classes.py:46, :465-474): arrays are (nx, ny, nz), C order, y = line of sight.
"""
import json
import os

import numpy as np
import scipy.constants as con
from scipy.special import hyp2f1, wofz

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "rajepy_b200", "data")

MSOL = 1.98847e30  # _constants.py:5
NZ = {"H": (1, 0), "He": (2, 2), "Li": (3, 4), "Be": (4, 5), "B": (5, 6), "C": (6, 6),
      "N": (7, 7), "O": (8, 8), "F": (9, 10), "Ne": (10, 10), "Na": (11, 12),
      "Mg": (12, 12)}  # _constants.py:7-10

c_cgs = con.c * 1e2  # maths/rrls.py:7-11
h_cgs = con.h * 1e7
k_cgs = con.k * 1e7


# ------------------------------------------------------------------ scalars / tables
def atomic_mass(atom):
    """kg.  maths/physics.py:607-624 (table lookup of the isotope in NZ)."""
    with open(os.path.join(_DATA, "atomic_masses.json"), "rt") as f:
        tab = json.load(f)
    m = tab[atom]["mass_micro_u"]
    m *= 1e-6 * con.u
    return m


def mod_r_0(opang, epsilon, w_0):
    """maths/geometry.py:12-31"""
    return epsilon * w_0 / np.tan(np.radians(opang) / 2.)


def rho(r, r_0, mr0=None):
    """maths/geometry.py:34-61"""
    if mr0:
        return (np.abs(r) + mr0 - r_0) / mr0
    return np.abs(r) / r_0


def w_r(r, w_0, mr0, r_0, eps):
    """maths/geometry.py:96-118"""
    return w_0 * rho(r, r_0, mr0) ** eps


def r_eff(w, r_1, r_2, w_0, r, mr0, r_0, eps):
    """maths/geometry.py:305-336"""
    return r_1 + ((r_2 - r_1) * w) / w_r(r, w_0, mr0, r_0, eps)


def n_0_from_mlr(mlr, v_0, w_0, mu, q_nd, q_nv, R_1, R_2):
    """cm^-3.  maths/physics.py:474-517"""
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


def xyz_rotate(x, y, z, alpha, beta, order='xy'):
    """maths/geometry.py:212-263 (degrees; individual roundings of every mul/add)."""
    a = np.radians(alpha)
    b = np.radians(beta)
    ca, sa = np.cos(a), np.sin(a)
    cb, sb = np.cos(b), np.sin(b)

    def xr(x_, y_, z_):
        return x_, ca * y_ - sa * z_, sa * y_ + ca * z_

    def yr(x_, y_, z_):
        return cb * x_ + sb * z_, y_, cb * z_ - sb * x_

    if order == 'xy':
        return yr(*xr(x, y, z))
    if order == 'yx':
        return xr(*yr(x, y, z))
    raise ValueError(order)


def xyz_to_rwp(x, y, z, inc, pa):
    """maths/geometry.py:181-209 and :266-302 -> (r, w, phi)."""
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
    i_rads = np.radians(params["geometry"]["inc"])
    pa_rads = np.radians(params["geometry"]["pa"])
    l_xz_au = params['grid']['l_z'] * params['target']['dist']
    xmax_au = l_xz_au * np.sin(pa_rads)
    ymax_au = l_xz_au * np.tan(1.571 - i_rads)
    zmax_au = l_xz_au * np.cos(pa_rads)
    rmax_au, _, __ = xyz_to_rwp(xmax_au, ymax_au, zmax_au,
                                params["geometry"]["inc"], params["geometry"]["pa"])
    wmax_au = w_r(rmax_au, params["geometry"]["w_0"], params["geometry"]["mod_r_0"],
                  params["geometry"]["r_0"], params["geometry"]["epsilon"])
    wmax_cells = int(np.ceil(np.abs(wmax_au / cs_au)))
    dims = [int(np.ceil(np.abs(v / cs_au))) + 2 * wmax_cells
            for v in (xmax_au, ymax_au, zmax_au)]
    return tuple(d if d % 2 == 0 else d + 1 for d in dims)


def rrl_parser(rrl_str):
    """maths/rrls.py:605-624"""
    dn = {'a': 1, 'b': 2, 'g': 3, 'd': 4}[rrl_str[-1].lower()]
    el = ''.join(ch for ch in rrl_str[:-1] if ch.isalpha())
    n = ''.join(ch for ch in rrl_str[:-1] if not ch.isalpha())
    return el, int(n), dn


def z_number(atom):
    """maths/physics.py:523-532"""
    return {'H': 1, 'He': 2, 'Li': 3, 'Be': 4, 'B': 5, 'C': 6, 'N': 7, 'O': 8}[atom]


def rrl_nu_0(atom, n, delta_n=1):
    """Hz.  maths/rrls.py:14-29 with maths/physics.py:535-544"""
    m_atom = atomic_mass(atom)
    ryd = con.Rydberg * (m_atom / (m_atom + con.m_e))
    return ryd * con.c * z_number(atom) ** 2. * (1. / n ** 2. - 1. / (n + delta_n) ** 2.)


def f_n1n2(n_1, delta_n):
    """maths/rrls.py:44-59"""
    m = {1: 0.190775, 2: 0.026332, 3: 0.0081056, 4: 0.0034918}[delta_n]
    return n_1 * m * (1. + 1.5 * delta_n / n_1)


def energy_n(n, atom):
    """erg.  maths/rrls.py:32-41"""
    return -2.17989724e-11 * z_number(atom) ** 2. / n ** 2.


def ni_from_ne(n_e, atom='H'):
    """maths/rrls.py:62-83"""
    xyz = {'H': 0.710, 'He': 0.276, 'CNO': 0.014}
    mu = (xyz['H'] / atomic_mass("H") * con.u + xyz['He'] / atomic_mass("He") * con.u +
          xyz['CNO'] / 14.24) ** -1.
    m_atom = atomic_mass(atom) / con.u
    return xyz[atom] * n_e * mu / m_atom


_GAUNT = None


def _gaunt_table():
    """maths/physics.py:626-663 (axes rebuilt with linspace of the rounded ends)."""
    global _GAUNT
    if _GAUNT is None:
        d = np.load(os.path.join(_DATA, "gaunt_vanhoof2014.npz"))
        g = d["gff"]
        n_u, n_g = g.shape
        step = float(d["step"])
        lus = np.linspace(np.round(float(d["log_u_start"]), decimals=1),
                          np.round(float(d["log_u_start"]) + step * (n_u - 1), decimals=1),
                          n_u)
        lgs = np.linspace(np.round(float(d["log_gamma2_start"]), decimals=1),
                          np.round(float(d["log_gamma2_start"]) + step * (n_g - 1),
                                   decimals=1), n_g)
        lg2, lu2 = np.meshgrid(lgs, lus)
        _GAUNT = (lg2, lu2, g)
    return _GAUNT


def gff(freq, temp, z=1.):
    """van Hoof (2014) Gaunt factor, scalar.  maths/physics.py:666-697, including the
    row clamp that uses the COLUMN count (:687-690).  `interp2d(kind='cubic')` on the
    5x5 scattered patch == FITPACK bisplrep(kx=ky=3, s=0) + bisplev."""
    from scipy.interpolate import bisplrep, bisplev
    ry = con.m_e * con.e ** 4. / (8 * con.epsilon_0 ** 2. * con.h ** 2.)
    logg2 = np.log10(z ** 2. * ry / (con.k * temp))
    logu = np.log10(con.h * freq / (con.k * temp))
    lg2s, lus, g = _gaunt_table()
    col = int(np.argmin(np.abs(lg2s[0] - logg2)))
    row = int(np.argmin(np.abs(lus[:, 0] - logu)))
    ncol = len(lg2s[0])
    col = min(max(col, 2), ncol - 3)
    row = min(max(row, 2), ncol - 3)
    sl = (slice(row - 2, row + 3), slice(col - 2, col + 3))
    tck = bisplrep(lg2s[sl].ravel(), lus[sl].ravel(), g[sl].ravel(), kx=3, ky=3, s=0.0)
    return float(np.ravel(bisplev(np.atleast_1d(logg2), np.atleast_1d(logu), tck))[0])


# ------------------------------------------------------------------ the model
class OracleJet:
    """Numpy restatement of JetModel's hot path (classes.py:42-1541).

    Grids are held as full (nx, ny, nz) float64 arrays like the reference does; use it
    at sizes that finish in seconds-minutes (<= 256^3)."""

    def __init__(self, params, time_s=0.0):
        import copy
        p = copy.deepcopy(params)
        self.p = p
        g, pl, pr, tg = p["geometry"], p["power_laws"], p["properties"], p["target"]
        # classes.py:168-180
        g["mod_r_0"] = mod_r_0(g["opang"], g["epsilon"], g["w_0"])
        pl["q_n"] = -pl["q_v"] - (2.0 * g["epsilon"])  # physics.py:17-35
        pl["q_tau"] = g["epsilon"] + 2.0 * pl["q_x"] + 2.0 * pl["q_n"] - 1.35 * pl["q_T"]
        # classes.py:188-213
        if p["grid"]["l_z"] is not None:
            self.nx, self.ny, self.nz = lz_to_grid_dims(p)
        else:
            self.nx = (p["grid"]["n_x"] + 1) // 2 * 2
            self.ny = (p["grid"]["n_y"] + 1) // 2 * 2
            self.nz = (p["grid"]["n_z"] + 1) // 2 * 2
        self.cs = p["grid"]["c_size"]
        # classes.py:228-242
        self.f_rb = pr["mlr_rj"] / pr["mlr_bj"]
        self.ss_bj = pr["mlr_bj"] * (1.989e30 / con.year)
        self.ss_rj = self.ss_bj * self.f_rb
        pr["n_0"] = n_0_from_mlr(pr["mlr_bj"], pr["v_0"], g["w_0"], pr["mu"],
                                 pl["q^d_n"], pl["q^d_v"], tg["R_1"], tg["R_2"])
        # classes.py:249-264: bursts (t_0 [s], peak [kg/s], half-life [s]) per jet
        self.bursts = {"R": [], "B": []}
        ej = p["ejection"]
        for i, t0 in enumerate(ej["t_0"]):
            which = ej["which"][i]
            if 'R' in which:
                self.bursts["R"].append((t0 * con.year, self.ss_rj * ej["chi"][i],
                                         ej["hl"][i] * con.year))
            if 'B' in which:
                self.bursts["B"].append((t0 * con.year, self.ss_bj * ej["chi"][i],
                                         ej["hl"][i] * con.year))
        self.time = time_s
        self._c = {}

    # -- coordinates (classes.py:465-557)
    def corners(self):
        cs = self.cs
        ix = np.arange(self.nx).reshape(-1, 1, 1)
        iy = np.arange(self.ny).reshape(1, -1, 1)
        iz = np.arange(self.nz).reshape(1, 1, -1)
        full = (self.nx, self.ny, self.nz)
        return tuple(np.ascontiguousarray(np.broadcast_to(cs * (i - n // 2), full))
                     for i, n in ((ix, self.nx), (iy, self.ny), (iz, self.nz)))

    def n_verts_inside(self):
        """classes.py:657-666 -- the integer count that must be reproduced bit-exactly."""
        if "nv" in self._c:
            return self._c["nv"]
        g = self.p["geometry"]
        cs = self.cs
        xx, yy, zz = self.corners()
        nv = np.zeros(xx.shape, dtype=int)
        for dx, dy, dz in ((0., 0., 0.), (cs, 0., 0.), (0., cs, 0.), (cs, cs, 0.),
                           (0., 0., cs), (cs, 0., cs), (0., cs, cs), (cs, cs, cs)):
            rv, wv = xyz_to_rwp(xx + dx, yy + dy, zz + dz, g["inc"], g["pa"])[:2]
            with np.errstate(all='ignore'):
                wrv = w_r(rv, g["w_0"], g["mod_r_0"], g["r_0"], g["epsilon"])
            nv = np.where((wrv >= wv) & (np.abs(rv) >= g["r_0"]), nv + 1, nv)
        self._c["nv"] = nv
        return nv

    def fill_factor(self):
        """classes.py:667-668, :763"""
        nv = self.n_verts_inside()
        ff = np.where(nv == 8, 1.0, np.where(nv > 0, 0.5, np.nan))
        return ff

    def areas(self):
        """classes.py:669, :764"""
        return np.where(self.n_verts_inside() > 0, 1.0, np.nan)

    def rwp(self):
        """classes.py:515-526"""
        if "rwp" not in self._c:
            xx, yy, zz = self.corners()
            h = self.cs / 2.
            g = self.p["geometry"]
            self._c["rwp"] = xyz_to_rwp(xx + h, yy + h, zz + h, g["inc"], g["pa"])
        return self._c["rwp"]

    def rreff(self):
        """classes.py:543-557 (note abs(r), not the base-shifted r)."""
        if "reff" not in self._c:
            g, tg = self.p["geometry"], self.p["target"]
            r, w, _ = self.rwp()
            with np.errstate(all='ignore'):
                self._c["reff"] = r_eff(w, tg["R_1"], tg["R_2"], g["w_0"], np.abs(r),
                                        g["mod_r_0"], g["r_0"], g["epsilon"])
        return self._c["reff"]

    def r_shifted(self):
        """classes.py:848-850 (= :884-886, :922-924, :1050-1052)"""
        r0 = self.p["geometry"]["r_0"]
        r = np.abs(self.rwp()[0])
        return np.where((r < r0) & ((r + self.cs / 2.) >= r0),
                        (r0 + r + self.cs / 2.) / 2., r)

    def travel_time(self):
        """Travel time from the jet base to the cell [s], i.e. the reference's
        ``_ts`` (classes.py:852 with maths/geometry.py:121-178).  hyp2f1 is called on
        whole arrays instead of through np.vectorize (same scipy routine)."""
        if "tt" in self._c:
            return self._c["tt"]
        g, pl, pr, tg = (self.p[k] for k in ("geometry", "power_laws", "properties",
                                              "target"))
        w_0 = g['w_0'] * con.au
        r_0 = g['r_0'] * con.au
        v_0 = pr["v_0"] * 1e3
        mr0 = g['mod_r_0'] * con.au
        eps = g['epsilon']
        r_1 = tg["R_1"] * con.au
        r_2 = tg["R_2"] * con.au
        q_v = pl["q_v"]
        q_vd = pl["q^d_v"]

        def indef(r_, w_):
            with np.errstate(all='ignore'):
                const = mr0 ** q_v / (v_0 * (1. - q_v + eps * q_vd))
                rad = r_ + mr0 - r_0
                p1 = rad ** (1. - q_v)
                p2 = (r_eff(w_, r_1, r_2, w_0, r_, mr0, r_0, eps) / r_1) ** -q_vd
                arg = (r_1 * w_0 * rad ** eps)
                wz = np.where(w_ == 0., 1., w_)
                p3 = (arg / (wz * mr0 ** eps * (r_2 - r_1)) + 1.) ** q_vd
                p4 = hyp2f1(q_vd, (1. - q_v + eps * q_vd) / eps,
                            (1. - q_v + eps + eps * q_vd) / eps,
                            arg / (wz * mr0 ** eps * (r_1 - r_2)))
                p3 = np.where(w_ == 0., 1.0, p3)
                p4 = np.where(w_ == 0., 1. + q_vd / (1. - q_v), p4)
                return const * p1 * p2 * p3 * p4

        r = self.r_shifted()
        w = self.rwp()[1]
        t_yr = (indef(np.abs(r) * con.au, w * con.au) -
                indef(np.full_like(w, r_0), w * con.au)) / con.year
        self._c["tt"] = t_yr * con.year
        return self._c["tt"]

    def ts(self):
        """Launch time of the material in each cell [s].  classes.py:838-859"""
        return self.time - self.travel_time()

    def _jml(self, which, t):
        """classes.py:442-448"""
        ss = self.ss_bj if which == 'B' else self.ss_rj
        out = ss
        for t0, peak, hl in self.bursts[which]:
            amp = peak - ss
            sigma = hl * 2. / (2. * np.sqrt(2. * np.log(2.)))
            out = out + amp * np.exp(-(t - t0) ** 2. / (2. * sigma ** 2.))
        return out

    def chi_xyz(self):
        """classes.py:861-870"""
        ts = self.ts()
        return np.where(self.rwp()[0] < 0, self._jml('R', ts) / self.ss_rj,
                        self._jml('B', ts) / self.ss_bj)

    def _masked_law(self, zero, q, qd, r_for_rho):
        """cell_value + masks: classes.py:889-897 / :928-934 / :961-967"""
        g, tg = self.p["geometry"], self.p["target"]
        with np.errstate(all='ignore'):
            v = zero * rho(r_for_rho, g["r_0"], g["mod_r_0"]) ** q * \
                (self.rreff() / tg["R_1"]) ** qd
        v = np.where(self.fill_factor() > 0, v, np.nan)
        v = np.where(v == 0, np.nan, v)
        return v

    def nd_base(self):
        """classes.py:872-899 without the burst factor."""
        if "nd" not in self._c:
            pl, pr = self.p["power_laws"], self.p["properties"]
            nd = self._masked_law(pr["n_0"], pl["q_n"], pl["q^d_n"], self.r_shifted())
            nd = np.where(self.rwp()[0] < 0, nd * self.f_rb, nd)
            self._c["nd"] = np.nan_to_num(nd, nan=np.nan, posinf=np.nan, neginf=np.nan)
        return self._c["nd"]

    def number_density(self):
        return self.nd_base() * self.chi_xyz()

    def ion_fraction(self):
        """classes.py:910-936"""
        if "xi" not in self._c:
            pl, pr = self.p["power_laws"], self.p["properties"]
            xi = self._masked_law(pr["x_0"], pl["q_x"], pl["q^d_x"], self.r_shifted())
            self._c["xi"] = np.nan_to_num(xi, nan=np.nan, posinf=np.nan, neginf=np.nan)
        return self._c["xi"]

    def temperature(self):
        """classes.py:942-969 incl. the cm-vs-au quirk at :957-959."""
        if "T" not in self._c:
            pl, pr, g = self.p["power_laws"], self.p["properties"], self.p["geometry"]
            r = np.abs(self.rwp()[0]) * con.au * 1e2
            r = np.where((r < g["r_0"]) & ((r + self.cs / 2.) >= g["r_0"]),
                         (g["r_0"] + r + self.cs / 2.) / 2., r)
            t = self._masked_law(pr["T_0"], pl["q_T"], pl["q^d_T"], r)
            self._c["T"] = np.nan_to_num(t, nan=np.nan, posinf=np.nan, neginf=np.nan)
        return self._c["T"]

    def vel(self):
        """km/s, (vx, v_los, vz).  classes.py:1009-1095, maths/physics.py:66-90"""
        if "vel" in self._c:
            return self._c["vel"]
        pl, pr, g, tg = (self.p[k] for k in ("power_laws", "properties", "geometry",
                                              "target"))
        r, w, ph = self.rwp()
        ffpos = self.fill_factor() > 0
        vz = self._masked_law(pr["v_0"], pl["q_v"], pl["q^d_v"], self.r_shifted())
        vz = np.nan_to_num(vz, nan=np.nan, posinf=np.nan, neginf=np.nan) * np.sign(r)
        with np.errstate(all='ignore'):
            vr = np.sqrt(con.G * tg["M_star"] * MSOL / (self.rreff() * con.au)) * \
                rho(r, g["r_0"], g["mod_r_0"]) ** -g["epsilon"] / 1e3
        sgn = 1 if g["rotation"].lower() == 'ccw' else -1
        vx = -vr * np.sin(ph) * sgn
        vy = vr * np.cos(ph) * sgn
        vx = np.where(ffpos, vx, np.nan)
        vy = np.where(ffpos, vy, np.nan)
        vz = np.where(ffpos, vz, np.nan)
        vxs, vys, vzs = xyz_rotate(vx, vy, vz, 90. - g["inc"], -g["pa"], order='xy')
        self._c["vel"] = (vxs, vys + tg["v_lsr"], vzs)
        return self._c["vel"]

    # -- line-of-sight integrals
    def _path_cm(self):
        # ff / areas == ff wherever the cell is in the jet (areas is 1 or NaN)
        return self.cs * con.au * 1e2 * (self.fill_factor() / self.areas())

    def emission_measure(self):
        """pc cm^-6.  classes.py:1116-1120"""
        ems = (self.number_density() * self.ion_fraction()) ** 2. * \
              (self.cs * con.au / con.parsec * (self.fill_factor() / self.areas()))
        return np.nansum(ems, axis=1)

    def optical_depth_ff(self, freq, collapse=True):
        """classes.py:1353-1447"""
        if not np.isscalar(freq):
            return np.array([self.optical_depth_ff(float(f), collapse) for f in freq])
        n_es = self.number_density() * self.ion_fraction()
        t = self.temperature()
        if self.p['power_laws']['q_T'] == 0.:
            g = gff(freq, self.p['properties']['T_0'])
        else:
            g = 11.95 * t ** 0.15 * freq ** -0.1
        with np.errstate(all='ignore'):
            tff = (0.018 * t ** -1.5 * freq ** -2. * n_es ** 2. * self._path_cm() * g)
        return np.nansum(tff, axis=1) if collapse else tff

    def mean_temperature(self):
        """classes.py:1471-1472 / :1254-1256"""
        t = self.temperature()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.nanmean(np.where(t > 0., t, np.nan), axis=1)

    def intensity_ff(self, freq):
        """W m^-2 Hz^-1 sr^-1.  classes.py:1449-1496"""
        if not np.isscalar(freq):
            return np.array([self.intensity_ff(float(f)) for f in freq])
        temp_b = self.mean_temperature() * (1. - np.exp(-self.optical_depth_ff(freq)))
        return 2. * freq ** 2. * con.k * temp_b / con.c ** 2.

    def pixel_solid_angle(self):
        return np.arctan((self.cs * con.au) /
                         (self.p["target"]["dist"] * con.parsec)) ** 2.

    def flux_ff(self, freq):
        """Jy/pixel.  classes.py:1498-1541"""
        return self.intensity_ff(freq) * self.pixel_solid_angle() / 1e-26

    def optical_depth_rrl(self, rrl, freq, collapse=True):
        """classes.py:1130-1229 with maths/rrls.py:86-118, :329-389"""
        element, n, dn = rrl_parser(rrl)
        v_los = self.vel()[1]
        rest = rrl_nu_0(element, n, dn) * (1. - v_los * 1000. / con.c)  # physics.py:547
        n_es = self.number_density() * self.ion_fraction()
        t = self.temperature()
        m = atomic_mass(element)
        with np.errstate(all='ignore'):
            fwhm_g = np.sqrt(4. * np.log(2.) * 2. * con.k * t / (m * con.c ** 2.)) * rest
            fwhm_l = 8.2 * n_es * (n / 100.) ** 4.5 * (1. + 4.5 / 2. * dn / n)
        fn = f_n1n2(n, dn)
        en = energy_n(n, element)
        z = z_number(element)
        n_i = ni_from_ne(n_es, element)
        sigma = fwhm_g / 2. / np.sqrt(2. * np.log(2))

        def one(f):
            with np.errstate(all='ignore'):
                phi = np.real(wofz(((f - rest) + 1j * fwhm_l / 2.) / sigma /
                                   np.sqrt(2.))) / sigma / np.sqrt(2. * np.pi)
                p1 = n ** 2. * fn * phi
                p2 = n_es * n_i / t ** 1.5
                p3 = np.exp((z ** 2. * en) / (k_cgs * t))
                p4 = 1. - np.exp(-h_cgs * f / (k_cgs * t))
                tau = 1.0991132675738456e-17 * p1 * p2 * p3 * p4 * self._path_cm()
            return np.nansum(tau, axis=1) if collapse else tau

        if np.isscalar(freq):
            return one(freq)
        return np.array([one(float(f)) for f in freq])

    def intensity_rrl(self, rrl, freq):
        """classes.py:1231-1290 (scalar branch, the one Pipeline reaches) with
        maths/rrls.py:428-449 and maths/physics.py:561-574"""
        if not np.isscalar(freq):
            return np.array([self.intensity_rrl(rrl, float(f)) for f in freq])
        av_t = self.mean_temperature()
        tau_l = self.optical_depth_rrl(rrl, freq)
        tau_c = self.optical_depth_ff(freq)
        with np.errstate(all='ignore'):
            p1 = 2. * con.h * 1e7 * freq ** 3. / (con.c * 1e2) ** 2.
            p2 = np.exp(con.h * 1e7 * freq / (con.k * 1e7 * av_t)) - 1.
            b_nu = p1 * p2 ** -1.
            return b_nu * np.exp(-tau_c) * (1. - np.exp(-tau_l)) * 1e-7 * 1e4

    def flux_rrl(self, rrl, freq, contsub=True):
        """Jy/pixel.  classes.py:1292-1351"""
        if not np.isscalar(freq):
            return np.array([self.flux_rrl(rrl, float(f), contsub) for f in freq])
        fl = self.intensity_rrl(rrl, freq) * self.pixel_solid_angle() / 1e-26
        if not contsub:
            fl = fl + self.flux_ff(freq)
        return fl


def chan_freqs(freq, bandwidth, chanwidth):
    """classes.py:1893-1900"""
    nchan = int(bandwidth / chanwidth)
    chan1 = freq - bandwidth / 2. + chanwidth / 2.
    return chan1 + np.arange(nchan) * chanwidth
