"""
`JetModel` -- drop-in replacement for the reference's `RaJePy.classes.JetModel`
(classes.py:42-1713) whose grid fill and line-of-sight radiative transfer run in
hand-written sm_100a CUDA kernels behind the C ABI of include/rajepy_b200.h.

Same constructor, properties, method names, argument meaning, return shapes
(numpy float64; 3-D fields (nx, ny, nz) with NaN outside the jet; images (nx, nz);
cubes (nfreq, nx, nz)) and error conventions as the reference.  Host code only derives
scalars (fp64, scipy.constants), owns device buffers through torch, and calls the
library through ctypes.  There is NO CPU fallback: without a CUDA device or without
the compiled library every compute call raises.

Multi-GPU: with `shard=(rank, world)` the model fills and integrates only an x-slab of
the grid and `gather=True` methods all-gather the image tiles over
torch.distributed (NCCL on GPUs); no other exchange exists on this path.
"""
import os
import pickle
import sys
import time as _time

import numpy as np
import scipy.constants as con

from . import _cabi
from . import hostmath as hm
from . import logger
from .sharding import balanced_bounds, even_bounds, gather_x

_TIE_CAPACITY = 1 << 16
_TIE_ASYNC = 2048         # ties read back with the asynchronous report (the rest on demand)

# number of kernel launches issued through the C ABI by this process (bench.py reports it)
LAUNCHES = {"count": 0}


def _launched(n=1):
    LAUNCHES["count"] += n


def _torch():
    import torch
    return torch


# Grid-state buffers (vertex counts 1 B/cell + packed cells 16 B/cell) are recycled between
# models of the same slab size together with their per-brick occupancy map: the fill then
# rewrites only the bricks around the jet (and zeroes bricks that held data of the previous
# model) instead of the whole 17 B/cell state.  A buffer that is not in the pool is allocated
# zero-filled, which is the state an all-zero occupancy map describes.
_STATE_POOL = {}
_PLANE_WEIGHTS = {}     # slab-balancing estimate per geometry (host work, ~20 ms once)
_SLABS = {}             # balanced slab boundaries per (plane weights, world)
_LINE_STRUCTS = {}      # line constants + per-channel device arrays
_TIE_DECISIONS = {}     # host re-decisions of near-tie vertices per (geometry, slab, ties)
_CONT_COEFFS = {}       # per-frequency continuum coefficients on the device
_N_ACTIVE = {}          # jet-crossing rays per (geometry, grid, slab): grid size of the line kernel


def _dev_index(torch, dev):
    return dev.index if dev.index is not None else torch.cuda.current_device()


# Hand-over of dense cubes to the host (JetModel._host_cube): measured rates of the two engines
# that fill the host array side by side -- host threads assembling planes from the packed
# jet-crossing columns, and the GPU's copy engine writing whole planes over PCIe.
_HANDOVER = {"cpu_gbs": None, "dma_gbs": None}


def _host_threads():
    env = os.environ.get("RAJEPY_B200_HOST_THREADS")
    if env:
        return max(1, int(env))
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    # one process per GPU: the ranks of a node share the cores
    n //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(n, 64))


def _take_state(torch, dev, nxs, ny, nz, nbricks):
    """Recycled (nverts, cells, brick map) of a slab with exactly this layout, else fresh
    zero-filled buffers.  The flat cell index and the brick id both depend on (ny, nz), so
    the key is the layout, not the cell count."""
    key = (_dev_index(torch, dev), nxs, ny, nz)
    hit = _STATE_POOL.pop(key, None)
    if hit is not None and hit[2].numel() == max(nbricks, 1):
        # the buffers may have been last written on another stream
        torch.cuda.current_stream(dev).wait_event(hit[3])
        return hit[:3]
    ncell = nxs * ny * nz
    return (torch.zeros(ncell, dtype=torch.uint8, device=dev),
            torch.zeros((ncell, 2), dtype=torch.float64, device=dev),
            torch.zeros(max(nbricks, 1), dtype=torch.uint8, device=dev))


def _give_state(d):
    """Return a model's state buffers to the pool (at most one set per device)."""
    try:
        torch = _torch()
        dev = d["device"]
        idx = _dev_index(torch, dev)
        if d["bricks"].numel() >= 1:
            for k in [k for k in _STATE_POOL if k[0] == idx]:
                del _STATE_POOL[k]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            _STATE_POOL[(idx,) + tuple(d["layout"])] = (d["nverts"], d["cells"], d["bricks"], ev)
    except Exception:  # noqa: BLE001 -- recycling is an optimisation only
        pass


def clear_state_pool():
    """Free the recycled grid-state buffers."""
    _STATE_POOL.clear()


class JetModel:
    """
    Class to handle physical model of an ionised jet from a young stellar object
    (API of the reference class, classes.py:42).
    """
    _arr_indexing = 'ij'  # numpy.meshgrid indexing type (classes.py:46)

    # ------------------------------------------------------------------ construction
    @classmethod
    def load_model(cls, model_file, **kwargs):
        """Load a model saved with `save` (classes.py:48-88)."""
        from .compat import load_pickle
        model_file = os.path.expanduser(model_file)
        loaded = load_pickle(model_file)     # also reads files written by the reference
        log = kwargs.pop('log', None)
        if log is None:
            log = loaded.get('log')
            if log is not None and not os.path.isdir(os.path.dirname(
                    os.path.abspath(log.filename))):
                # saved on another machine / in a directory that is gone: keep the entries,
                # continue the log beside the save file
                log._filename = os.path.join(os.path.dirname(os.path.abspath(model_file)),
                                             os.path.basename(log.filename))
        if log is None:
            log = logger.Log(os.path.expanduser('~') + os.sep + 'temp.log')
        new_jm = cls(loaded["params"], log=log, **kwargs)
        if loaded.get('ffs') is not None:
            new_jm._adopt_fill_factor(loaded['ffs'])
        new_jm.time = loaded['time']
        return new_jm

    @staticmethod
    def lz_to_grid_dims(params):
        """classes.py:90-122"""
        return hm.lz_to_grid_dims(params)

    @staticmethod
    def py_to_dict(py_file):
        """Parameter file (python module exposing `params`) -> dict (classes.py:124-142).
        Unlike the reference's validator (miscellaneous/functions.py:127-190) a missing
        properties.n_0 is accepted: it is always overwritten (classes.py:234-242) and the
        reference's own example file omits it."""
        if not os.path.exists(py_file):
            raise FileNotFoundError(py_file + " does not exist")
        import importlib.util
        name = "_rajepy_params_" + str(abs(hash(os.path.abspath(py_file))))
        spec = importlib.util.spec_from_file_location(name, py_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        err = check_model_params(getattr(mod, "params", None))
        if err is not None:
            raise err
        return mod.params

    def __init__(self, params, log=None, device=None, shard=None, balance=True,
                 host_ranks=None, shard_axis='x'):
        if isinstance(params, dict):
            self._params = params
        elif isinstance(params, str):
            self._params = JetModel.py_to_dict(params)
        else:
            raise TypeError("Supplied arg params must be dict or file path (str)")

        p = self._params
        self._name = p['target']['name']
        self._csize = p['grid']['c_size']

        # automatically calculated parameters (classes.py:168-180)
        mr0 = hm.mod_r_0(p['geometry']['opang'], p['geometry']['epsilon'],
                         p['geometry']['w_0'])
        qn = hm.q_n(p["geometry"]["epsilon"], p["power_laws"]["q_v"])
        qtau = hm.q_tau(p["geometry"]["epsilon"], p["power_laws"]["q_x"], qn,
                        p["power_laws"]["q_T"])
        p["geometry"]["mod_r_0"] = mr0
        p["power_laws"]["q_n"] = qn
        p["power_laws"]["q_tau"] = qtau

        if log is not None:
            self._log = log
        else:
            self._log = logger.Log(os.path.expanduser('~') + os.sep + 'temp.log',
                                   verbose=True)

        # grid dimensions (classes.py:188-213)
        if p['grid']['l_z'] is not None:
            nx, ny, nz = JetModel.lz_to_grid_dims(p)
            self.log.add_entry("INFO",
                               'For a (bipolar) jet length of {:.1f}", cell '
                               'size of {:.2f}au and distance of {:.0f}pc, a '
                               'grid size of (n_x, n_y, n_z) = ({}, {}, {}) '
                               'voxels is calculated'
                               ''.format(p['grid']['l_z'], p["grid"]["c_size"],
                                         p["target"]["dist"], nx, ny, nz))
        else:
            nx = (p['grid']['n_x'] + 1) // 2 * 2
            ny = (p['grid']['n_y'] + 1) // 2 * 2
            nz = (p['grid']['n_z'] + 1) // 2 * 2
        p['grid']['n_x'], p['grid']['n_y'], p['grid']['n_z'] = nx, ny, nz
        self._nx, self._ny, self._nz = int(nx), int(ny), int(nz)

        # steady-state mass-loss rates (classes.py:228-242)
        self._ss_jml_rb_frac = p["properties"]["mlr_rj"] / p["properties"]["mlr_bj"]
        self._ss_jml_bj = p["properties"]["mlr_bj"] * (1.989e30 / con.year)
        self._ss_jml_rj = self._ss_jml_bj * self._ss_jml_rb_frac
        p["properties"]["n_0"] = hm.n_0_from_mlr(
            p["properties"]["mlr_bj"], p["properties"]["v_0"], p["geometry"]["w_0"],
            p["properties"]["mu"], p["power_laws"]["q^d_n"], p["power_laws"]["q^d_v"],
            p["target"]["R_1"], p["target"]["R_2"])

        # ejection bursts (classes.py:244-264)
        self._bursts = {'R': [], 'B': []}
        self._jml_t_bj = lambda t: self._ss_jml_bj
        self._jml_t_rj = lambda t: self._ss_jml_rj
        self._ejections = {}
        for idx, ejn_t0 in enumerate(p['ejection']['t_0']):
            which = p['ejection']['which'][idx]
            if 'R' in which:
                self.add_ejection_event(ejn_t0 * con.year,
                                        self._ss_jml_rj * p['ejection']['chi'][idx],
                                        p['ejection']['hl'][idx] * con.year, which='R')
            if 'B' in which:
                self.add_ejection_event(ejn_t0 * con.year,
                                        self._ss_jml_bj * p['ejection']['chi'][idx],
                                        p['ejection']['hl'][idx] * con.year, which='B')

        self._time = 0. * con.year

        # device side
        self._device_arg = device
        if shard is None:
            shard = (0, 1)
        if shard_axis not in ('x', 'tile', 'channel'):
            raise ValueError("shard_axis must be 'x', 'tile' or 'channel'")
        # 'tile': x-slabs like 'x', but the line cubes STAY sky tiles (nchan, nx_slab, nz) on
        # their ranks -- nothing but the small sky images and per-channel totals is exchanged on
        # the device; a host cube is assembled through an all-to-all of the packed jet-crossing
        # columns (`_host_cube`).  'x' completes full-size cubes on every rank.
        self._tiles = shard_axis == 'tile'
        # shard_axis='channel': every rank holds the whole grid and integrates a contiguous
        # block of the channels of a line cube (sharding.chan_bounds); continuum products are
        # replicated, cube planes stay with their rank (no exchange).  'x': x-slabs.
        self._chan_rank, self._chan_world = 0, 1
        if shard_axis == 'channel':
            self._chan_rank, self._chan_world = int(shard[0]), int(shard[1])
            if not (0 <= self._chan_rank < self._chan_world):
                raise ValueError("bad (rank, world)")
            shard = (0, 1)
        self._rank, self._world = int(shard[0]), int(shard[1])
        # sharded models: ranks that receive the products as host (numpy) arrays; the others
        # take part in the exchange and get None (e.g. host_ranks=(0,) when only rank 0 writes
        # the FITS files).  None = every rank, the drop-in behaviour.
        self._host_ranks = None if host_ranks is None else {int(r) for r in host_ranks}
        # x-slabs of equal estimated work (in-jet cells), not equal width: see _plane_weights
        # (a plane of empty sky is not free: its rows of every cube plane are constants to be
        # written.  Measured per slab at N = 8 (tools/slab_probe.py, 1024^3 x 512 channels):
        # 1.5 ns per ray of sky (16 B x 512 channels at ~5.5 TB/s) against 1.3 ns per in-jet
        # cell of the channel loop, i.e. ~1.2 in-jet cells per ray of the plane; the two costs
        # overlap, but not for free, see balanced_bounds)
        self._bounds = self._balanced_slabs() \
            if (balance and self._world > 1) else even_bounds(self._nx, self._world)
        self._x_lo, self._x_hi = self._bounds[self._rank]
        self._dev = None       # dict of device buffers once filled
        self._fields = {}      # cached host copies of 3-D property grids
        self._overrides = {}   # user-assigned grids (setters)
        self._cont = None      # cached continuum pass (device tensors)
        self._line = None      # cached line pass
        self._timings = {}
        self._coeff_cache = {}

    def _balanced_slabs(self):
        """Work-balanced x-slabs of this geometry for `world` ranks; a pure function of the
        parameters, remembered per geometry (the search is ~ms of host time, a step is too)."""
        w = self._plane_weights()
        key = (id(w), self._world, self._nz)
        hit = _SLABS.get(key)
        if hit is None or hit[0] is not w:
            if len(_SLABS) > 64:
                _SLABS.clear()
            hit = (w, balanced_bounds(w, self._world, plane_cost=1.2 * self._nz, overlap=0.3))
            _SLABS[key] = hit
        return list(hit[1])

    def _plane_weights(self):
        """Estimated number of in-jet cells of every x-plane from a coarse sample of cell
        centres (classes.py:661-666 at the centroid): the work per plane of the sparse fill and
        of the ray walk.  Pure function of the parameters, so every rank derives the same
        slab boundaries without communicating."""
        g = self._params["geometry"]
        nx, ny, nz, cs = self._nx, self._ny, self._nz, self._csize
        key = (nx, ny, nz, float(cs), float(g["inc"]), float(g["pa"]), float(g["w_0"]),
               float(g["r_0"]), float(g["mod_r_0"]), float(g["epsilon"]))
        hit = _PLANE_WEIGHTS.get(key)
        if hit is not None:
            return hit
        sx, sy, sz = max(1, nx // 256), max(1, ny // 64), max(1, nz // 64)
        ix = np.arange(sx // 2, nx, sx)
        iy = np.arange(sy // 2, ny, sy)
        iz = np.arange(sz // 2, nz, sz)
        x = cs * (ix - nx // 2 + 0.5)
        y = cs * (iy - ny // 2 + 0.5)
        z = cs * (iz - nz // 2 + 0.5)
        xx, yy, zz = np.meshgrid(x, y, z, indexing="ij")
        r, w, _ = hm.xyz_to_rwp(xx, yy, zz, g["inc"], g["pa"])
        with np.errstate(all="ignore"):
            inside = (w <= hm.w_r(r, g["w_0"], g["mod_r_0"], g["r_0"], g["epsilon"])) & \
                     (np.abs(r) >= g["r_0"])
        coarse = inside.sum(axis=(1, 2)).astype(np.float64) * (sx * sy * sz) / sx
        out = np.interp(np.arange(nx), ix, coarse)
        if len(_PLANE_WEIGHTS) > 32:
            _PLANE_WEIGHTS.clear()
        _PLANE_WEIGHTS[key] = out
        return out

    # ------------------------------------------------------------------ text table
    def __str__(self):
        """Same table as the reference (classes.py:268-361); it becomes the FITS HISTORY."""
        p = self.params
        h = ['Parameter', 'Value']
        d = [('epsilon', format(p['geometry']['epsilon'], '+.3f')),
             ('opang', format(p['geometry']['opang'], '+.0f') + ' deg'),
             ('q_v', format(p['power_laws']['q_v'], '+.3f')),
             ('q_T', format(p['power_laws']['q_T'], '+.3f')),
             ('q_x', format(p['power_laws']['q_x'], '+.3f')),
             ('q_n', format(p['power_laws']['q_n'], '+.3f')),
             ('q^d_v', format(p['power_laws']['q^d_v'], '+.3f')),
             ('q^d_T', format(p['power_laws']['q^d_T'], '+.3f')),
             ('q^d_x', format(p['power_laws']['q^d_x'], '+.3f')),
             ('q^d_n', format(p['power_laws']['q^d_n'], '+.3f')),
             ('q_tau', format(p['power_laws']['q_tau'], '+.3f')),
             ('cell', format(p['grid']['c_size'], '.1f') + ' au'),
             ('w_0', format(p['geometry']['w_0'], '.2f') + ' au'),
             ('r_0', format(p['geometry']['r_0'], '.2f') + ' au'),
             ('v_0', format(p['properties']['v_0'], '.0f') + ' km/s'),
             ('x_0', format(p['properties']['x_0'], '.3f')),
             ('n_0', format(p['properties']['n_0'], '.3e') + ' cm^-3'),
             ('T_0', format(p['properties']['T_0'], '.0e') + ' K'),
             ('f_R2B', format(self._ss_jml_rb_frac, '.2e')),
             ('i', format(p['geometry']['inc'], '+.1f') + ' deg'),
             ('theta', format(p['geometry']['pa'], '+.1f') + ' deg'),
             ('D', format(p['target']['dist'], '+.0f') + ' pc'),
             ('M*', format(p['target']['M_star'], '+.1f') + ' Msol'),
             ('R_1', format(p['target']['R_1'], '+.1f') + ' au'),
             ('R_2', format(p['target']['R_2'], '+.1f') + ' au')]
        if len(p['ejection']['t_0']) > 0:
            d.append(('t_now', format(self.time / con.year, '+.3f') + ' yr'))

        c1 = max(len(s) for s in [h[0]] + [r[0] for r in d]) + 2
        c2 = max(len(s) for s in [h[1]] + [r[1] for r in d]) + 2
        width = c1 + c2 + 3
        hline = width * '-'

        def row(cols, widths):
            return '|' + '|'.join(format(c, '^' + str(w)) for c, w in zip(cols, widths)) + '|\n'

        s = hline + '\n' + '/' + format('JET MODEL', '^' + str(width - 2)) + '/\n' + hline + '\n'
        s += row(h, (c1, c2)) + hline + '\n'
        for line_ in d:
            s += row(line_, (c1, c2))
        s += hline + '\n'
        s += '/' + format('BURSTS', '^' + str(width - 2)) + '/\n' + hline + '\n'
        db = [(format(t, '.2f'), format(p["ejection"]["hl"][i], '.2f'),
               format(p["ejection"]["chi"][i], '.2f'))
              for i, t in enumerate(p["ejection"]["t_0"])]
        if len(db) == 0:
            return s + '|' + format(' None ', '-^' + str(width - 2)) + '|\n' + hline + '\n'
        b1 = b2 = b3 = (width - 4) // 3
        if (width - 4) % 3 > 0:
            b1 += 1
            if (width - 4) % 3 == 2:
                b2 += 1
        for line_ in (['t_0', 'FWHM', 'chi'], ['[yr]', '[yr]', '']):
            s += row(line_, (b1, b2, b3))
        s += hline + '\n'
        for line_ in db:
            s += row(line_, (b1, b2, b3))
        return s + hline + '\n'

    # ------------------------------------------------------------------ simple attributes
    @property
    def los_axis(self):
        if self._arr_indexing == 'ij':
            return 1
        elif self._arr_indexing == 'xy':
            return 0
        raise ValueError(f"Unknown numpy array indexing ({self._arr_indexing})")

    @property
    def time(self):
        """Model time in seconds"""
        return self._time

    @time.setter
    def time(self, new_time):
        self._time = new_time

    @property
    def log(self):
        return self._log

    @log.setter
    def log(self, new_log):
        self._log = new_log

    @property
    def csize(self):
        return self._csize

    @property
    def nx(self):
        return self._nx

    @property
    def ny(self):
        return self._ny

    @property
    def nz(self):
        return self._nz

    @property
    def params(self):
        return self._params

    @property
    def name(self):
        return self._name

    @property
    def ejections(self):
        return self._ejections

    @property
    def slab(self):
        """x-range [x_lo, x_hi) of the grid held by this process."""
        return self._x_lo, self._x_hi

    def ss_jml(self, which):
        """classes.py:1694-1702"""
        if which == 'R':
            return self._ss_jml_rj
        elif which == 'B':
            return self._ss_jml_bj
        elif 'R' in which and 'B' in which:
            return self._ss_jml_rj + self._ss_jml_bj
        raise ValueError("which must be one of 'R', 'B', or 'RB'")

    def jml_t(self, which):
        """Callable t [s] -> jet mass-loss rate [kg/s] (classes.py:383-397)"""
        def inner_func(t):
            jml = 0.
            if 'R' in which:
                jml += self._jml_t_rj(t)
            if 'B' in which:
                jml += self._jml_t_bj(t)
            return jml
        return inner_func

    def add_ejection_event(self, t_0, peak_jml, half_life, which):
        """Gaussian ejection burst (classes.py:399-463).  t_0, half_life in s,
        peak_jml in kg/s, which in ('R', 'B')."""
        assert which in ('R', 'B')
        if len(self._bursts[which]) >= _cabi.MAX_BURSTS:
            raise ValueError(f"at most {_cabi.MAX_BURSTS} bursts per jet are supported")
        ss_jml = self._ss_jml_bj if which == 'B' else self._ss_jml_rj
        amp = peak_jml - ss_jml
        sigma = half_life * 2. / (2. * np.sqrt(2. * np.log(2.)))
        prev = self._jml_t_rj if which == 'R' else self._jml_t_bj

        def func2(t, _prev=prev, _amp=amp, _t0=t_0, _sigma=sigma):
            return _prev(t) + _amp * np.exp(-(t - _t0) ** 2. / (2. * _sigma ** 2.))

        if which == 'R':
            self._jml_t_rj = func2
        else:
            self._jml_t_bj = func2
        self._bursts[which].append((float(t_0), float(amp / ss_jml),
                                    float(1. / (2. * sigma ** 2.))))
        self._ejections[str(len(self._ejections) + 1)] = {
            't_0': t_0, 'peak_jml': peak_jml, 'half_life': half_life, 'which': which}
        self._cont = None
        self._line = None

    # ------------------------------------------------------------------ index/coordinate grids
    @property
    def indices(self):
        """classes.py:465-474 (host numpy; never needed by the CUDA path)"""
        return tuple(np.meshgrid(np.arange(self.nx), np.arange(self.ny), np.arange(self.nz),
                                 indexing=self._arr_indexing))

    @property
    def ix(self):
        return self.indices[0]

    @property
    def iy(self):
        return self.indices[1]

    @property
    def iz(self):
        return self.indices[2]

    @property
    def grid(self):
        """Corner coordinates in au (classes.py:488-501)"""
        ix, iy, iz = self.indices
        return (self.csize * (ix - self.nx // 2), self.csize * (iy - self.ny // 2),
                self.csize * (iz - self.nz // 2))

    @property
    def xx(self):
        return self.grid[0]

    @property
    def yy(self):
        return self.grid[1]

    @property
    def zz(self):
        return self.grid[2]

    @property
    def xs(self):
        return self.csize * (np.arange(self.nx) - self.nx // 2)

    @property
    def ys(self):
        return self.csize * (np.arange(self.ny) - self.ny // 2)

    @property
    def zs(self):
        return self.csize * (np.arange(self.nz) - self.nz // 2)

    @property
    def grid_rwp(self):
        return self._field('r'), self._field('w'), self._field('phi')

    @property
    def rr(self):
        return self._field('r')

    @property
    def ww(self):
        return self._field('w')

    @property
    def pp(self):
        return self._field('phi')

    @property
    def rreff(self):
        return self._field('reff')

    # ------------------------------------------------------------------ device plumbing
    def _device(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _cabi.EngineError(
                "rajepy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self._device_arg is not None:
            return torch.device(self._device_arg)
        return torch.device("cuda", torch.cuda.current_device())

    def _stream(self):
        return _torch().cuda.current_stream(self._device()).cuda_stream

    def _model_struct(self):
        p = self._params
        g, pl, pr, tg = p["geometry"], p["power_laws"], p["properties"], p["target"]
        m = _cabi.Model()
        m.nx, m.ny, m.nz = self._nx, self._ny, self._nz
        m.x_lo, m.x_hi = self._x_lo, self._x_hi
        m.cs = float(self._csize)
        m.w0, m.r0, m.mr0, m.eps = (float(g["w_0"]), float(g["r_0"]), float(g["mod_r_0"]),
                                    float(g["epsilon"]))
        m.ca, m.sa, m.cb, m.sb = hm.rotation_trig(g["inc"] - 90., g["pa"])
        m.cva, m.sva, m.cvb, m.svb = hm.rotation_trig(90. - g["inc"], -g["pa"])
        m.R1, m.R2 = float(tg["R_1"]), float(tg["R_2"])
        m.q_n, m.q_x, m.q_T, m.q_v = (float(pl["q_n"]), float(pl["q_x"]), float(pl["q_T"]),
                                      float(pl["q_v"]))
        m.qd_n, m.qd_x, m.qd_T, m.qd_v = (float(pl["q^d_n"]), float(pl["q^d_x"]),
                                          float(pl["q^d_T"]), float(pl["q^d_v"]))
        m.n0, m.x0, m.T0, m.v0 = (float(pr["n_0"]), float(pr["x_0"]), float(pr["T_0"]),
                                  float(pr["v_0"]))
        m.f_rb = float(self._ss_jml_rb_frac)
        m.gm_over_au = float(con.G * tg["M_star"] * hm.MSOL / con.au)
        m.rot_sign = 1.0 if g["rotation"].lower() == 'ccw' else -1.0
        m.v_lsr = float(tg["v_lsr"])
        m.au_m, m.year_s, m.au_cm = float(con.au), float(con.year), float(con.au * 1e2)
        a = m.qd_v
        b = (1. - m.q_v + m.eps * a) / m.eps
        m.hyp_b = b
        m.hyp_c1 = m.hyp_c2 = 0.0
        m.hyp_degenerate = 1
        if a != 0.0:
            if tg["R_2"] <= tg["R_1"]:
                raise ValueError("q^d_v != 0 requires R_2 > R_1")
            from scipy.special import gamma, rgamma
            d = b - a
            if abs(d - round(d)) > 1e-3:
                m.hyp_degenerate = 0
                m.hyp_c1 = b / (b - a)
                m.hyp_c2 = float(gamma(b + 1.) * gamma(a - b) * rgamma(a))
        m.need_reff = int(any(pl[k] != 0. for k in ("q^d_n", "q^d_x", "q^d_T")))
        return m

    def _epoch_struct(self):
        e = _cabi.Epoch()
        e.time = float(self._time)
        e.n_blue, e.n_red = len(self._bursts['B']), len(self._bursts['R'])
        for key, arr in (('B', e.blue), ('R', e.red)):
            for i, (t0, amp, inv2s2) in enumerate(self._bursts[key]):
                arr[i].t0, arr[i].amp, arr[i].inv2s2 = t0, amp, inv2s2
        return e

    def _ensure_filled(self, sync=True):
        """Run the grid fill (K1+K2) once; resolve near-tie vertices on the host with
        the reference's own numpy expression so that the counts are bit-exact.

        The fill reports its near-tie vertices through an asynchronous copy into pinned host
        memory.  `sync=False` (the line-of-sight passes) returns without waiting for it: the
        pass is queued right behind the fill and `_validate_fill` is consulted afterwards --
        only if a host decision changes a vertex count (none does on the BASELINE grids) is the
        pass repeated.  Every other consumer validates first."""
        if self._dev is None:
            self._launch_fill()
        if sync:
            self._validate_fill()
        return self._dev

    def _launch_fill(self, tie_cap=_TIE_CAPACITY):
        torch = _torch()
        lib = _cabi.load()
        dev = self._device()
        nxs = self._x_hi - self._x_lo
        self._fill_t0 = _time.time()
        if self.log:
            self._log.add_entry(mtype="INFO",
                                entry="Calculating cells' fill factors/projected areas")
        with torch.cuda.device(dev):
            m = self._model_struct()
            nbricks = int(lib.rjp_brick_count(m))
            nverts, cells, bricks = _take_state(torch, dev, nxs, self._ny, self._nz, nbricks)
            # [0:8] counters ([0] = number of near-tie vertices), [8:] the tie list
            tb = torch.empty(8 + 4 * tie_cap, dtype=torch.int32, device=dev)
            tb[:8].zero_()
            work = torch.empty(nbricks + 4, dtype=torch.int32, device=dev)   # fill work list
            nray = nxs * self._nz
            extents = torch.empty((nray, 2), dtype=torch.int32, device=dev)
            st = lib.rjp_fill_grid(m, nverts.data_ptr(), cells.data_ptr(),
                                   bricks.data_ptr(), work.data_ptr(), tb.data_ptr() + 32,
                                   tie_cap, tb.data_ptr(), extents.data_ptr(), self._stream())
            _cabi.check(st, "rjp_fill_grid")
            _launched(3)   # init_extents, fill_classify, fill_bricks kernels
            n_async = min(tie_cap, _TIE_ASYNC)
            pin = torch.empty(8 + 4 * n_async, dtype=torch.int32, pin_memory=True)
            pin.copy_(tb[:8 + 4 * n_async], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._dev = {"nverts": nverts, "cells": cells, "bricks": bricks,
                         "extents": extents, "model": m, "layout": (nxs, self._ny, self._nz),
                         "device": dev, "stream2": torch.cuda.Stream(device=dev),
                         "n_ties": None, "n_patched": 0,
                         "tie_pending": (pin, ev, tb, tie_cap)}
            self._apply_overrides()
            self._build_ray_list()
        return self._dev

    def _apply_overrides(self):
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        dev = d["device"]
        for name, field in (('xi', 8), ('temp', 9)):     # RJP_FIELD_XI, RJP_FIELD_TEMP
            if name in self._overrides:
                arr = np.ascontiguousarray(
                    np.asarray(self._overrides[name], dtype=np.float64)
                    [self._x_lo:self._x_hi])
                if arr.shape != (self._x_hi - self._x_lo, self._ny, self._nz):
                    raise ValueError(f"assigned grid has shape {arr.shape}")
                vals = torch.from_numpy(arr).to(dev)
                st = lib.rjp_override_cells(d["model"], d["nverts"].data_ptr(), field,
                                            vals.data_ptr(), d["cells"].data_ptr(),
                                            self._stream())
                _cabi.check(st, "rjp_override_cells")
                _launched()
        # user-assigned travel-time / velocity grids (the `ts` / `vel` setters): the integrators
        # read them per cell instead of recomputing the analytic laws
        d["travel_grid"] = d["vlos_grid"] = None
        shape = (self._x_hi - self._x_lo, self._ny, self._nz)
        if 'travel' in self._overrides:
            arr = np.ascontiguousarray(np.asarray(self._overrides['travel'], dtype=np.float64)
                                       [self._x_lo:self._x_hi])
            if arr.shape != shape:
                raise ValueError(f"assigned ts grid has shape {arr.shape}")
            d["travel_grid"] = torch.from_numpy(arr).to(dev)
        if 'vel' in self._overrides:
            arr = np.ascontiguousarray(np.asarray(self._overrides['vel'][1], dtype=np.float64)
                                       [self._x_lo:self._x_hi])
            if arr.shape != shape:
                raise ValueError(f"assigned v_los grid has shape {arr.shape}")
            d["vlos_grid"] = torch.from_numpy(arr).to(dev)

    def _validate_fill(self):
        """Wait for the fill's tie report (not for anything queued behind it), let the host
        decide the reported vertices and patch the cells whose count changes.  Returns True if
        the state changed (passes launched meanwhile are stale and have been dropped)."""
        d = self._dev
        pend = d.pop("tie_pending", None) if d is not None else None
        if pend is None:
            return False
        pin, ev, tb, tie_cap = pend
        ev.synchronize()
        n_ties = int(pin[0])
        if n_ties > tie_cap:
            # the list overflowed: fill again with room for every tie (never seen on real
            # grids: dyadic axis-aligned grids report tens of vertices)
            self._dev = None
            self._cont = self._line = None
            _give_state(d)
            self._launch_fill(tie_cap=n_ties + 1024)
            self._validate_fill()
            return True
        d["n_ties"] = n_ties
        changed = False
        if n_ties > 0:
            if n_ties <= (pin.numel() - 8) // 4:
                ties = pin[8:8 + 4 * n_ties].view(n_ties, 4).numpy()
            else:
                ties = tb[8:8 + 4 * n_ties].view(n_ties, 4).cpu().numpy()
            changed = self._resolve_ties(ties.astype(np.int64))
        if changed:
            torch = _torch()
            with torch.cuda.device(d["device"]):
                self._apply_overrides()
                self._build_ray_list()
            self._cont = self._line = None
            self._fields.clear()
        if self.log:
            self.log.add_entry(mtype="INFO",
                               entry=_time.strftime('Finished in %Hh%Mm%Ss',
                                                    _time.gmtime(_time.time() - self._fill_t0)))
        return changed

    def _build_ray_list(self):
        """Ordered list of the rays that cross the jet; its length stays on the device (the
        ray kernels read it there)."""
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        dev = d["device"]
        nray = d["extents"].shape[0]
        rays = torch.empty(max(nray, 1), dtype=torch.int32, device=dev)
        counts = torch.empty(max(int(lib.rjp_ray_list_chunks(nray)), 1), dtype=torch.int32,
                             device=dev)
        n_act = torch.zeros(1, dtype=torch.int32, device=dev)
        st = lib.rjp_ray_list(d["extents"].data_ptr(), nray, rays.data_ptr(),
                              counts.data_ptr(), n_act.data_ptr(), self._stream())
        _cabi.check(st, "rjp_ray_list")
        _launched(2)
        d["rays"], d["n_active_dev"], d["n_active_host"] = rays, n_act, None
        d["ray_cells_host"] = None
        d["ray_meta"] = None

    def _n_active(self):
        """Number of jet-crossing rays of the slab on the HOST (blocks until the ray list is
        built)."""
        d = self._ensure_filled()
        if d["n_active_host"] is None:
            torch = _torch()
            ext = d["extents"]
            cells = (ext[:, 1] - ext[:, 0]).clamp_min(0).sum(dtype=torch.int64)
            both = torch.stack([d["n_active_dev"][0].to(torch.int64), cells]).cpu()
            d["n_active_host"], d["ray_cells_host"] = int(both[0]), int(both[1])
            if len(_N_ACTIVE) > 256:
                _N_ACTIVE.clear()
            if not d.get("custom_fill"):
                _N_ACTIVE[self._geometry_key()] = (d["n_active_host"], d["ray_cells_host"])
        return d["n_active_host"]

    def _geometry_key(self):
        g = self._params["geometry"]
        return (self._nx, self._ny, self._nz, self._x_lo, self._x_hi, float(self._csize),
                float(g["inc"]), float(g["pa"]), float(g["w_0"]), float(g["r_0"]),
                float(g["mod_r_0"]), float(g["epsilon"]))

    def _n_active_hint(self):
        """Grid size of the line kernel (one CTA per jet-crossing ray is fastest).  The count is
        a pure function of the geometry and the grid, so it is remembered per geometry: the
        first model of a geometry reads it back once, later ones (time series, repeated runs)
        never wait for the device.  The kernel itself reads the true count on the device, so a
        stale hint could only cost time, not correctness."""
        return self._ray_counts()[0]

    def _ray_counts(self):
        """(jet-crossing rays, summed lengths of their extents) of the slab, from the
        per-geometry memory when this geometry has been filled before (see _n_active_hint)."""
        d = self._dev
        if d["n_active_host"] is None:
            hit = None if d.get("custom_fill") else _N_ACTIVE.get(self._geometry_key())
            if hit is not None:
                return hit
            self._n_active()
        return d["n_active_host"], d["ray_cells_host"]

    def _decide_ties(self, I, J, K, dec):
        """{flat slab cell index: change of its vertex count} from the reference's own numpy
        expression (classes.py:658-666) for every (cell, corner) touching a near-tie vertex."""
        g = self._params["geometry"]
        cs = self._csize
        ny, nz = self._ny, self._nz
        delta = {}
        for a in (0, 1):
            for b in (0, 1):
                for c in (0, 1):
                    i, j, k = I - a, J - b, K - c
                    ok = (i >= self._x_lo) & (i < self._x_hi) & (j >= 0) & (j < ny) & \
                         (k >= 0) & (k < nz)
                    if not ok.any():
                        continue
                    ii, jj, kk = i[ok], j[ok], k[ok]
                    x = cs * (ii - self._nx // 2) + (cs if a else 0.)
                    y = cs * (jj - self._ny // 2) + (cs if b else 0.)
                    z = cs * (kk - self._nz // 2) + (cs if c else 0.)
                    rv, wv = hm.xyz_to_rwp(x, y, z, g["inc"], g["pa"])[:2]
                    with np.errstate(all='ignore'):
                        wrv = hm.w_r(rv, g["w_0"], g["mod_r_0"], g["r_0"], g["epsilon"])
                    ref = ((wrv >= wv) & (np.abs(rv) >= g["r_0"])).astype(np.int64)
                    diff = ref - dec[ok]
                    flat = ((ii - self._x_lo) * ny + jj) * nz + kk
                    for f, df in zip(flat[diff != 0], diff[diff != 0]):
                        delta[int(f)] = delta.get(int(f), 0) + int(df)
        return {f: v for f, v in delta.items() if v != 0}

    def _resolve_ties(self, ties):
        """Vertices whose inside test the device could not call: decide every
        (cell, corner) that touches them exactly as classes.py:658-666 does -- corner
        coordinate cs*(i - n//2) + {0|cs}, maths/geometry.py:181-209 and :96-118 in numpy
        -- and patch the cells whose count changes."""
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        g = self._params["geometry"]
        cs = self._csize
        I, J, K, dec = ties[:, 0], ties[:, 1], ties[:, 2], ties[:, 3]
        ny, nz = self._ny, self._nz
        # the decisions are a pure function of the geometry, the slab and the reported ties:
        # models built from the same parameters (time series, repeated runs) reuse them
        order = np.lexsort((K, J, I))
        key = (self._nx, ny, nz, float(cs), self._x_lo, self._x_hi, float(g["inc"]),
               float(g["pa"]), float(g["w_0"]), float(g["r_0"]), float(g["mod_r_0"]),
               float(g["epsilon"]), ties[order].tobytes())
        delta = _TIE_DECISIONS.get(key)
        if delta is None:
            delta = self._decide_ties(I, J, K, dec)
            if len(_TIE_DECISIONS) > 64:
                _TIE_DECISIONS.clear()
            _TIE_DECISIONS[key] = delta
        d["n_patched"] = len(delta)
        if not delta:
            return False
        dev = d["device"]
        idx = torch.tensor(sorted(delta), dtype=torch.int64, device=dev)
        old = d["nverts"][idx].to(torch.int64)
        new = old + torch.tensor([delta[f] for f in sorted(delta)], dtype=torch.int64,
                                 device=dev)
        if int(new.min()) < 0 or int(new.max()) > 8:
            raise _cabi.EngineError("tie resolution produced an impossible vertex count")
        new8 = new.to(torch.uint8)
        st = lib.rjp_patch_cells(d["model"], idx.data_ptr(), new8.data_ptr(), idx.numel(),
                                 d["nverts"].data_ptr(), d["cells"].data_ptr(),
                                 d["bricks"].data_ptr(), d["extents"].data_ptr(),
                                 self._stream())
        _cabi.check(st, "rjp_patch_cells")
        _launched()
        return True

    def _adopt_fill_factor(self, ffs):
        """Resume path (classes.py:78-84): take fill factors from a saved model instead
        of recomputing them.  Counts 1..7 are not recoverable from ff = 0.5; any value in
        that range gives the same physics (ff is all the integrators use)."""
        torch = _torch()
        ffs = np.asarray(ffs, dtype=np.float64)
        if ffs.shape != (self._nx, self._ny, self._nz):
            raise ValueError("saved fill factors do not match the grid dimensions")
        saved_overrides, self._overrides = self._overrides, {}
        self._dev = None
        d = self._ensure_filled()
        self._overrides = saved_overrides
        sl = ffs[self._x_lo:self._x_hi].reshape(-1)
        want = np.where(sl == 1.0, 8, np.where(sl > 0, 4, 0)).astype(np.uint8)
        have = d["nverts"].cpu().numpy()
        cls_have = np.where(have == 8, 8, np.where(have > 0, 4, 0))
        diff = np.flatnonzero(cls_have != want)
        if diff.size:
            lib = _cabi.load()
            dev = d["device"]
            idx = torch.from_numpy(diff.astype(np.int64)).to(dev)
            new8 = torch.from_numpy(want[diff]).to(dev)
            st = lib.rjp_patch_cells(d["model"], idx.data_ptr(), new8.data_ptr(),
                                     idx.numel(), d["nverts"].data_ptr(),
                                     d["cells"].data_ptr(), d["bricks"].data_ptr(),
                                     d["extents"].data_ptr(), self._stream())
            _cabi.check(st, "rjp_patch_cells")
            _launched()
            # cells entered / left the jet: the ray list of the model's own fill is stale
            d["custom_fill"] = True
            with torch.cuda.device(d["device"]):
                self._build_ray_list()
        self._fields.clear()
        self._cont = self._line = None

    def n_verts_inside(self, gather=True):
        """Number of cell vertices inside the jet, uint8 (nx, ny, nz): the integer the
        reference forms at classes.py:657-666 before mapping it to fill factors."""
        d = self._ensure_filled()
        t = d["nverts"].view(self._x_hi - self._x_lo, self._ny, self._nz)
        if gather and self._world > 1:
            t = gather_x(t, self._nx, self._rank, self._world, dim=0, bounds=self._bounds)
        return t.cpu().numpy()

    def _field_device(self, name):
        torch = _torch()
        lib = _cabi.load()
        d = self._ensure_filled()
        dev = d["device"]
        ncell = d["nverts"].numel()
        out = torch.empty(ncell, dtype=torch.float64, device=dev)
        ep = self._epoch_struct()
        st = lib.rjp_cell_field(d["model"], ep, d["nverts"].data_ptr(), _cabi.FIELDS[name],
                                out.data_ptr(), self._stream())
        _cabi.check(st, "rjp_cell_field")
        _launched()
        return out.view(self._x_hi - self._x_lo, self._ny, self._nz)

    def _field(self, name, cache=True, gather=True):
        if name in self._overrides:
            return self._overrides[name]
        if cache and name in self._fields:
            return self._fields[name]
        t = self._field_device(name)
        if gather and self._world > 1:
            t = gather_x(t, self._nx, self._rank, self._world, dim=0, bounds=self._bounds)
        arr = t.cpu().numpy()
        if cache:
            self._fields[name] = arr
        return arr

    # ------------------------------------------------------------------ 3-D properties
    @property
    def fill_factor(self):
        """classes.py:571-769"""
        return self._field('fill_factor')

    @property
    def areas(self):
        """classes.py:771-784"""
        return self._field('areas')

    @property
    def ts(self):
        """Launch time of the material in each cell [s] (classes.py:838-859)"""
        if 'travel' in self._overrides:
            return self.time - self._overrides['travel']
        return self.time - self._field('travel')

    @ts.setter
    def ts(self, new_ts):
        self._overrides['travel'] = new_ts
        self._invalidate()

    @property
    def chi_xyz(self):
        """Burst factor per cell (classes.py:861-870)"""
        if 'travel' in self._overrides:
            # user-assigned launch times: the reference's own expression on the host
            ts = self.ts
            with np.errstate(all='ignore'):
                return np.where(self.rr < 0, self._jml_t_rj(ts) / self._ss_jml_rj,
                                self._jml_t_bj(ts) / self._ss_jml_bj)
        return self._field('chi', cache=False)

    @property
    def number_density(self):
        """cm^-3, including the burst factor (classes.py:872-899)"""
        return self._field('nd_base') * self.chi_xyz

    @property
    def mass_density(self):
        """g cm^-3 (classes.py:901-908)"""
        av_m_particle = self.params['properties']['mu'] * hm.atomic_mass("H")
        return av_m_particle * 1e3 * self.number_density

    @property
    def ion_fraction(self):
        """classes.py:910-936"""
        return self._field('xi')

    @ion_fraction.setter
    def ion_fraction(self, new_xis):
        self._overrides['xi'] = new_xis
        self._invalidate()

    @property
    def temperature(self):
        """K (classes.py:942-969)"""
        return self._field('temp')

    @temperature.setter
    def temperature(self, new_ts):
        self._overrides['temp'] = new_ts
        self._invalidate()

    @property
    def pressure(self):
        """Barye (classes.py:1002-1007)"""
        return self.number_density * self.temperature * con.k * 1e7

    @property
    def vel(self):
        """(vx, v_los, vz) in km/s (classes.py:1009-1095)"""
        if 'vel' in self._overrides:
            return self._overrides['vel']
        return self._field('vx'), self._field('vlos'), self._field('vz')

    @vel.setter
    def vel(self, new_vs):
        self._overrides['vel'] = new_vs
        self._invalidate()

    def release(self):
        """Drop every device buffer held by the model (they are re-created on demand)."""
        self._invalidate()

    def _invalidate(self):
        if self._dev is not None:
            _give_state(self._dev)
        self._dev = None
        self._cont = None
        self._line = None
        self._fields.clear()

    # ------------------------------------------------------------------ line-of-sight passes
    def _continuum_struct(self):
        ct = _cabi.Continuum()
        ct.em_scale = float(self._csize * con.au / con.parsec)
        ct.tau_scale = float(0.018 * (self._csize * con.au * 1e2))
        ct.t_exponent = -1.5 if self._params['power_laws']['q_T'] == 0. else -1.35
        return ct

    def _ff_coeff(self, freqs):
        """tau_ff(nu) = coeff(nu) * K with K the frequency-independent ray sum
        (classes.py:1388-1399, :1421-1429).  Cached: the Gaunt-factor spline fits are
        host work that would otherwise sit between kernel launches."""
        freqs = np.asarray(freqs, dtype=np.float64)
        key = (freqs.tobytes(), float(self._params['properties']['T_0']),
               float(self._params['power_laws']['q_T']))
        hit = self._coeff_cache.get(key)
        if hit is not None:
            return hit
        if self._params['power_laws']['q_T'] == 0.:
            g = hm.gff(freqs, self._params['properties']['T_0'])
            out = freqs ** -2. * g
        else:
            out = 11.95 * freqs ** -0.1 * freqs ** -2.
        if len(self._coeff_cache) > 64:
            self._coeff_cache.clear()
        self._coeff_cache[key] = out
        return out

    def _pixel_solid_angle(self):
        return float(np.arctan((self._csize * con.au) /
                               (self._params["target"]["dist"] * con.parsec)) ** 2.)

    def _pass(self, line=None, freqs=None, contsub=True, want_tau=True, want_flux=True):
        """One fused line-of-sight pass over the packed state; caches the continuum sums
        of the current epoch and the last line cube."""
        torch = _torch()
        lib = _cabi.load()
        d = self._ensure_filled(sync=False)
        dev = d["device"]
        nxs, nz = self._x_hi - self._x_lo, self._nz
        key_c = (float(self._time), len(self._ejections))
        if line is None:
            if self._cont is not None and self._cont["key"] == key_c:
                return self._cont
        else:
            key_l = key_c + (line, np.asarray(freqs, np.float64).tobytes(), bool(contsub))
            if self._line is not None and self._line["key"] == key_l and \
                    (not want_tau or self._line["tau"] is not None) and \
                    (not want_flux or self._line["flux"] is not None):
                return self._line
        npix = nxs * nz
        with torch.cuda.device(dev):
            em = torch.empty(npix, dtype=torch.float64, device=dev)
            kff = torch.empty(npix, dtype=torch.float64, device=dev)
            tsum = torch.empty(npix, dtype=torch.float64, device=dev)
            cnt = torch.empty(npix, dtype=torch.int32, device=dev)
            ep = self._epoch_struct()
            ct = self._continuum_struct()
            tau = flux = None
            if line is None:
                st = lib.rjp_integrate(d["model"], ep, ct, d["cells"].data_ptr(),
                                       d["extents"].data_ptr(), d["rays"].data_ptr(),
                                       d["n_active_dev"].data_ptr(), -1,
                                       em.data_ptr(), kff.data_ptr(),
                                       tsum.data_ptr(), cnt.data_ptr(), None, None, 0, 1,
                                       None, None, 0, 0, *self._cell_grid_ptrs(), None, 0,
                                       self._stream(), None)
            else:
                c_lo, c_hi = 0, len(freqs)
                dn_all = None
                if self._chan_world > 1:
                    from .sharding import chan_bounds
                    c_lo, c_hi = chan_bounds(len(freqs), self._chan_rank, self._chan_world)
                    all_f = np.asarray(freqs, np.float64)
                    dn_all = float(np.max(np.abs(all_f - hm.rrl_nu_0(*hm.rrl_parser(line)))))
                    freqs = all_f[c_lo:c_hi]
                ln, chans, keep = self._line_structs(line, freqs, dev, dn_max=dn_all)
                nch = len(freqs)
                # sharded: every rank writes its slab straight into full-size cubes; the
                # other slabs arrive through the sparse exchange below
                full = self._world > 1 and not self._tiles
                rows = self._nx if full else nxs
                plane, coff = (self._nx * nz, self._x_lo * nz) if full else (0, 0)
                if want_tau:
                    tau = torch.empty((nch, rows, nz), dtype=torch.float64, device=dev)
                if want_flux:
                    flux = torch.empty((nch, rows, nz), dtype=torch.float64, device=dev)
                side = self._fill_remote_constants(tau, flux) if full else None
                lines = nch > 0       # (more ranks than channels: continuum sums only)
                n_hint, max_cells = self._ray_counts()
                scratch = torch.empty(int(lib.rjp_line_scratch_bytes(d["model"], max_cells)),
                                      dtype=torch.uint8, device=dev)
                st = lib.rjp_integrate(d["model"], ep, ct, d["cells"].data_ptr(),
                                       d["extents"].data_ptr(), d["rays"].data_ptr(),
                                       d["n_active_dev"].data_ptr(), n_hint,
                                       em.data_ptr(), kff.data_ptr(),
                                       tsum.data_ptr(), cnt.data_ptr(),
                                       ln if lines else None, chans if lines else None,
                                       nch, 1 if contsub else 0,
                                       tau.data_ptr() if (want_tau and lines) else None,
                                       flux.data_ptr() if (want_flux and lines) else None,
                                       plane, coff, *self._cell_grid_ptrs(),
                                       scratch.data_ptr(), max_cells,
                                       self._stream(), d["stream2"].cuda_stream)
                scratch.record_stream(d["stream2"])
                if full:
                    _cabi.check(st, "rjp_integrate")
                    self._exchange_cubes(tau, flux, side)
                del keep
            _cabi.check(st, "rjp_integrate")
            # writer + (ray walk | prepare + channel-loop launches)
            _launched(2 if line is None else 2 + (len(freqs) + 2047) // 2048)
        self._cont = {"key": key_c, "em": em, "kff": kff, "tsum": tsum, "cnt": cnt}
        if line is not None:
            self._line = {"key": key_l, "tau": tau, "flux": flux, "c_lo": c_lo, "c_hi": c_hi}
        if self._validate_fill():
            # a host-resolved near-tie changed a vertex count after this pass was queued:
            # the state has been patched, integrate again
            return self._pass(line, freqs, contsub, want_tau, want_flux)
        return self._cont if line is None else self._line

    def _cell_grid_ptrs(self):
        d = self._dev
        return tuple(None if d.get(k) is None else d[k].data_ptr()
                     for k in ("travel_grid", "vlos_grid"))

    def _line_structs(self, line, freqs, dev, dn_max=None):
        """Host scalars of the LTE line opacity (classes.py:1159-1169; rrls.py) and the
        per-channel device arrays.  Cached per (line, channels, model constants, device): the
        Gaunt-factor fits and the host->device copy would otherwise sit between the kernel
        launches of every pass."""
        torch = _torch()
        freqs = np.asarray(freqs, dtype=np.float64)
        key = (line, freqs.tobytes(), dn_max, str(dev), float(self._csize),
               float(self._params["target"]["dist"]),
               float(self._params['properties']['T_0']),
               float(self._params['power_laws']['q_T']),
               float(self._params['power_laws']['q^d_T']))
        hit = _LINE_STRUCTS.get(key)
        if hit is not None:
            return hit
        element, n, dn = hm.rrl_parser(line)
        nu0 = hm.rrl_nu_0(element, n, dn)
        m_atom = hm.atomic_mass(element)
        ln = _cabi.Line()
        ln.nu0 = float(nu0)
        ln.dopp = float(1000. / con.c)
        ln.width_g = float(np.sqrt(2. * con.k / (m_atom * con.c ** 2.)))
        ln.stark = float(hm.deltanu_l(1.0, n, dn) / 2.)
        z = hm.z_number(element)
        ln.kappa0 = float(1.0991132675738456e-17 * n ** 2. * hm.f_n1n2(n, dn) *
                          hm.ni_from_ne(1.0, element) * (self._csize * con.au * 1e2) /
                          np.sqrt(np.pi))
        ln.en_over_k = float(z ** 2. * hm.energy_n(n, element) / hm.k_cgs)
        ln.h_over_k = float(hm.h_cgs / hm.k_cgs)
        # (a rank of a channel-sharded cube passes the maximum over ALL channels, so that
        # every rank classifies a cell the same way)
        ln.dn_max = dn_max if dn_max is not None else \
            (float(np.max(np.abs(freqs - nu0))) if freqs.size else 0.0)
        # equally spaced channels (ContinuumRun.chan_freqs, classes.py:1893-1900): the kernels
        # form the offsets on the fly instead of reading them
        ln.chan_dnu0, ln.chan_step = float(freqs[0] - nu0) if freqs.size else 0.0, 0.0
        if freqs.size > 1:
            step = (freqs[-1] - freqs[0]) / (freqs.size - 1)
            dev_max = np.max(np.abs(freqs - (freqs[0] + step * np.arange(freqs.size))))
            if step != 0.0 and dev_max <= 1e-9 * abs(step):
                ln.chan_step = float(step)
        # isothermal jet (q_T = q^d_T = 0 and no assigned temperature grid: T = T_0 exactly in
        # every cell): the temperature-only factors are formed here once
        pl = self._params['power_laws']
        ln.t_common = 0.0
        if pl['q_T'] == 0. and pl['q^d_T'] == 0.:
            t0 = float(self._params['properties']['T_0'])
            ln.t_common = t0
            ln.tc_sqrt = float(np.sqrt(t0))
            ln.tc_boltz = float(np.exp(ln.en_over_k / t0))
            ln.tc_hk = float(ln.h_over_k / t0)
            ln.tc_p0 = float(-np.expm1(-ln.tc_hk * ln.nu0))
        omega_jy = self._pixel_solid_angle() / 1e-26
        host = np.stack([
            freqs - nu0,
            freqs,
            self._ff_coeff(freqs),
            2. * freqs ** 2. * con.k / con.c ** 2. * omega_jy,
            2. * con.h * 1e7 * freqs ** 3. / (con.c * 1e2) ** 2. * (1e-7 * 1e4) * omega_jy,
        ])
        devarr = torch.from_numpy(np.ascontiguousarray(host)).to(dev)
        ch = _cabi.Channels()
        base, step = devarr.data_ptr(), freqs.size * 8
        ch.dnu, ch.nu, ch.cff, ch.aff, ch.bnu = (base, base + step, base + 2 * step,
                                                 base + 3 * step, base + 4 * step)
        if len(_LINE_STRUCTS) > 16:
            _LINE_STRUCTS.clear()
        _LINE_STRUCTS[key] = (ln, ch, devarr)
        return ln, ch, devarr

    def _host_image(self, t, lead=None):
        """Device tile(s) -> full host numpy array, all-gathering x-slabs if sharded."""
        nxs, nz = self._x_hi - self._x_lo, self._nz
        wanted = self._host_ranks is None or self._rank in self._host_ranks
        if self._world > 1 and t.numel() == (lead or 1) * self._nx * nz:
            # already complete (cube finished by _exchange_cubes, images from _sky_sums)
            shape = (self._nx, nz) if lead is None else (lead, self._nx, nz)
            return _to_host(t.view(shape)) if wanted else None
        t = t.view(nxs, nz) if lead is None else t.view(lead, nxs, nz)
        if self._world > 1:
            t = gather_x(t, self._nx, self._rank, self._world, dim=0 if lead is None else 1,
                         bounds=self._bounds)
        return _to_host(t) if wanted else None

    def _host_cube(self, cube, fill, nch, sibling=None):
        """Device cube of a line pass -> numpy (nch, nx, nz) on the host.

        One GPU: `_handover` into a pinned array.  Channel-sharded (`shard_axis='channel'`):
        every rank hands ITS planes over ITS PCIe link into one host array in shared memory
        (`_shared_cube`), which the host ranks return; nothing travels between GPUs.
        x-slabs: the cube was completed on the device by the sparse exchange, plain copy."""
        torch = _torch()
        nxs, nz = self._x_hi - self._x_lo, self._nz
        plane = nxs * nz
        wanted = self._host_ranks is None or self._rank in self._host_ranks
        if self._chan_world > 1:
            return self._shared_cube(cube, fill, nch, sibling)
        if self._tiles and self._world > 1:
            return self._shared_cube_from_tiles(cube, fill, nch)
        if self._world > 1 or cube.numel() != nch * plane:
            return self._host_image(cube, lead=nch)
        out = torch.empty((nch, plane), dtype=torch.float64, pin_memory=True)
        self._handover(cube.view(nch, plane), fill, out,
                       None if sibling is None else sibling.view(nch, plane))
        return out.view(nch, nxs, nz).numpy() if wanted else None

    def _stage_columns(self, cube):
        """Queue (asynchronously) the packed columns of the jet-crossing rays of `cube`
        (n, plane) for the host: rjp_pack_rays -> copy into page-locked memory.  Returns
        (host columns, event); staged once per cube of the current line pass."""
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        dev = d["device"]
        staged = self._line.setdefault("staged", {}) if self._line is not None else {}
        hit = staged.get(cube.data_ptr())
        if hit is not None:
            return hit
        nch, plane = cube.shape
        n = self._n_active()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            if d.get("rays_host") is None:
                ids_h = torch.empty(n, dtype=torch.int32, pin_memory=True)
                ids_h.copy_(d["rays"][:n], non_blocking=True)
                d["rays_host"] = ids_h
            cols = torch.empty((nch, n), dtype=torch.float64, device=dev)
            _cabi.check(lib.rjp_pack_rays(cube.data_ptr(), plane, d["rays"].data_ptr(), n, n,
                                          nch, cols.data_ptr(), stream.cuda_stream),
                        "rjp_pack_rays")
            _launched()
            cols_h = torch.empty((nch, n), dtype=torch.float64, pin_memory=True)
            cols_h.copy_(cols, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        staged[cube.data_ptr()] = (cols_h, ev)
        return cols_h, ev

    def _handover(self, cube, fill, out, sibling=None):
        """cube (n, plane) on the device -> out (n, plane), a page-locked host tensor.

        The dense cube is mostly the constant `fill` (tau: 0, flux: NaN -- the rays that miss the
        jet), so only the packed columns of the jet-crossing rays cross PCIe (`_stage_columns`)
        and host threads produce the array: constants with streaming stores, columns dropped in
        (rjp_host_assemble; measured 135-170 GB/s with 16 threads against 57 GB/s for the plain
        copy of the dense cube, tools/e2e_probe.py).  `sibling`: the other cube of the same line
        pass; its columns are staged meanwhile, so that handing it over next starts at once.
        RAJEPY_B200_HOST_SPLIT < 1 leaves that fraction of the planes to the copy engine instead
        (no gain on the boxes measured: the two engines share the host memory bandwidth).
        The result is bit-identical to the plain copy."""
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        dev = d["device"]
        nch, plane = cube.shape
        if nch == 0:
            return
        n = self._n_active()
        split = float(os.environ.get("RAJEPY_B200_HOST_SPLIT", "1.0"))
        if plane % 2 or n == 0 or n > 0.3 * plane:   # (a jet that covers the sky: nothing to gain)
            split = 0.0
        a = max(0, min(nch, int(round(nch * max(0.0, min(1.0, split))))))
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            cols_h = ev = None
            if a > 0:
                cols_h, ev = self._stage_columns(cube)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            if a < nch:
                out[a:].copy_(cube[a:], non_blocking=True)
            e1.record(stream)
            if a > 0 and sibling is not None and sibling.shape == cube.shape:
                self._stage_columns(sibling)
            if a > 0:
                ev.synchronize()
                t0 = _time.perf_counter()
                st = lib.rjp_host_assemble(out.data_ptr(), a, plane, d["rays_host"].data_ptr(), n,
                                           cols_h.data_ptr(), n, float(fill), _host_threads())
                _cabi.check(st, "rjp_host_assemble")
                t_cpu = _time.perf_counter() - t0
                if t_cpu > 0:
                    _HANDOVER["cpu_gbs"] = a * plane * 8 / t_cpu / 1e9
            if a < nch:
                stream.synchronize()
                ms = e0.elapsed_time(e1)
                if ms > 0:
                    _HANDOVER["dma_gbs"] = (nch - a) * plane * 8 / (ms * 1e-3) / 1e9

    def _shared_cube(self, cube, fill, nch, sibling=None):
        """Channel-sharded hand-over: (nch, nx, nz) float64 in POSIX shared memory, every rank
        writing the planes it integrated (collective over the default process group).  The
        host ranks get the array, the others None."""
        import torch.distributed as dist
        from . import hostshare
        from .sharding import chan_bounds
        torch = _torch()
        plane = self._nx * self._nz
        c_lo, c_hi = chan_bounds(nch, self._chan_rank, self._chan_world)
        seg = hostshare.segment(nch * plane * 8, self._chan_rank)
        mine = seg.tensor(c_lo * plane * 8, (c_hi - c_lo, plane))     # page-locked view
        if c_hi > c_lo:
            self._handover(cube.view(c_hi - c_lo, plane), fill, mine,
                           None if sibling is None else sibling.view(c_hi - c_lo, plane))
        dist.barrier()
        wanted = self._host_ranks is None or self._chan_rank in self._host_ranks
        if not wanted:
            seg.release()
            return None
        return seg.array((nch, self._nx, self._nz))

    def _shared_cube_from_tiles(self, cube, fill, nch):
        """Tile-sharded hand-over.  The ranks hold sky tiles (nch, nx_slab, nz); the host cube is
        assembled by CHANNEL blocks (an even share of the host work whatever the slabs look
        like): one all-to-all over NVLink moves the packed columns of the jet-crossing rays --
        the only data that differs from the constant -- from the rank that integrated the ray
        to the rank that assembles the channel; then every rank copies its block's columns to the
        host and writes its planes of the shared cube (constants + columns)."""
        import torch.distributed as dist
        from . import hostshare
        from .sharding import chan_bounds
        torch = _torch()
        lib = _cabi.load()
        d = self._dev
        dev = d["device"]
        world, rank = self._world, self._rank
        nxs, nz = self._x_hi - self._x_lo, self._nz
        plane_tile, plane = nxs * nz, self._nx * nz
        n = self._n_active()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            # who lists how many rays, and which (global ids, ascending: the slabs are ordered)
            cnt = torch.tensor([n], dtype=torch.int64, device=dev)
            cnts = torch.empty(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(cnts, cnt)
            counts = [int(c) for c in cnts.cpu()]
            n_all, nmax = sum(counts), max(max(counts), 1)
            ids = torch.full((nmax,), -1, dtype=torch.int32, device=dev)
            ids[:n] = d["rays"][:n] + self._x_lo * nz
            all_ids = torch.empty((world, nmax), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(all_ids.view(-1), ids)
            ids_all = torch.cat([all_ids[r, :counts[r]] for r in range(world)])
            # own rays, all channels -> the channel blocks' owners
            cols = torch.empty((nch, max(n, 1)), dtype=torch.float64, device=dev)
            if n > 0:
                _cabi.check(lib.rjp_pack_rays(cube.data_ptr(), plane_tile, d["rays"].data_ptr(),
                                              n, n, nch, cols.data_ptr(), stream.cuda_stream),
                            "rjp_pack_rays")
                _launched()
            blocks = [chan_bounds(nch, r, world) for r in range(world)]
            c_lo, c_hi = blocks[rank]
            nb = c_hi - c_lo
            send = cols[:, :n].contiguous().view(-1) if n > 0 else cols.view(-1)[:0]
            recv = torch.empty(nb * n_all, dtype=torch.float64, device=dev)
            dist.all_to_all_single(recv, send,
                                   output_split_sizes=[nb * c for c in counts],
                                   input_split_sizes=[(hi - lo) * n for lo, hi in blocks])
            # [source][channel][ray of source] -> (channel, all rays)
            parts, off = [], 0
            for c in counts:
                parts.append(recv[off: off + nb * c].view(nb, c))
                off += nb * c
            mine = torch.cat(parts, dim=1).contiguous() if nb > 0 else recv.view(0, n_all)
            ids_h = torch.empty(n_all, dtype=torch.int32, pin_memory=True)
            ids_h.copy_(ids_all, non_blocking=True)
            cols_h = torch.empty((nb, n_all), dtype=torch.float64, pin_memory=True)
            cols_h.copy_(mine, non_blocking=True)
            stream.synchronize()
        seg = hostshare.segment(nch * plane * 8, rank)
        if nb > 0:
            out = seg.tensor(c_lo * plane * 8, (nb, plane), pin=False)
            st = lib.rjp_host_assemble(out.data_ptr(), nb, plane, ids_h.data_ptr(), n_all,
                                       cols_h.data_ptr(), n_all, float(fill), _host_threads())
            _cabi.check(st, "rjp_host_assemble")
        dist.barrier()
        wanted = self._host_ranks is None or rank in self._host_ranks
        if not wanted:
            seg.release()
            return None
        return seg.array((nch, self._nx, nz))

    def _continuum_images_device(self, freqs, want):
        """K5 for a list of frequencies; `want` in ('tau', 'intensity', 'flux').
        Returns a device tensor (nfreq, nx_slab * nz)."""
        torch = _torch()
        lib = _cabi.load()
        c = self._pass()
        if self._tiles and self._world > 1:
            c = self._sky_sums()
        dev = c["em"].device
        freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
        nf = freqs.size
        npix = c["em"].numel()
        coeff = self._cont_coeff_device(freqs, dev)
        with torch.cuda.device(dev):
            out = torch.empty((nf, npix), dtype=torch.float64, device=dev)
            ptrs = {k: (out.data_ptr() if k == want else None)
                    for k in ('tau', 'intensity', 'flux')}
            st = lib.rjp_continuum_images(c["kff"].data_ptr(), c["tsum"].data_ptr(),
                                          c["cnt"].data_ptr(), npix, coeff.data_ptr(),
                                          coeff.data_ptr() + nf * 8,
                                          self._pixel_solid_angle() / 1e-26, nf,
                                          ptrs['tau'], ptrs['intensity'], ptrs['flux'],
                                          self._stream())
            _cabi.check(st, "rjp_continuum_images")
            _launched()
        return out

    def _cont_coeff_device(self, freqs, dev):
        """(2, nfreq) device tensor: tau_ff = [0] * K (`_ff_coeff`) and the Rayleigh-Jeans factor
        2 nu^2 k / c^2 of the intensity (classes.py:1488); cached per frequency list."""
        torch = _torch()
        ckey = (freqs.tobytes(), str(dev), float(self._params['properties']['T_0']),
                float(self._params['power_laws']['q_T']))
        coeff = _CONT_COEFFS.get(ckey)
        if coeff is None:
            coeff = torch.from_numpy(np.stack([self._ff_coeff(freqs),
                                               2. * freqs ** 2. * con.k / con.c ** 2.])).to(dev)
            if len(_CONT_COEFFS) > 64:
                _CONT_COEFFS.clear()
            _CONT_COEFFS[ckey] = coeff
        return coeff

    def _continuum_epochs_device(self, times_s, freqs, want='flux', with_em=False):
        """Continuum images for a BATCH of model times without touching `self.time`: one ray
        walk for all of them (rjp_integrate_epochs: the epochs only differ in the burst factor
        chi(t) of each cell, classes.py:861-870) and one epilogue launch.  Returns a device
        tensor (n_epochs, nfreq, nx_slab * nz) of `want` in ('tau', 'intensity', 'flux')
        [, the emission measures (n_epochs, nx_slab * nz)]; each epoch equals what
        `time = t; flux_ff(freqs)` gives."""
        torch = _torch()
        lib = _cabi.load()
        if self._tiles and self._world > 1:
            raise NotImplementedError("epoch batches are sharded by epoch, not by sky tile")
        times_s = np.atleast_1d(np.asarray(times_s, dtype=np.float64))
        freqs = np.atleast_1d(np.asarray(freqs, dtype=np.float64))
        ne, nf = times_s.size, freqs.size
        while True:
            d = self._ensure_filled(sync=False)
            dev = d["device"]
            npix = (self._x_hi - self._x_lo) * self._nz
            with torch.cuda.device(dev):
                coeff = self._cont_coeff_device(freqs, dev)
                t_dev = torch.from_numpy(times_s).to(dev, non_blocking=True)
                kff = torch.zeros((ne, npix), dtype=torch.float64, device=dev)
                em = torch.zeros((ne, npix), dtype=torch.float64, device=dev) if with_em else None
                tsum = torch.zeros(npix, dtype=torch.float64, device=dev)
                cnt = torch.zeros(npix, dtype=torch.int32, device=dev)
                out = torch.empty((ne, nf, npix), dtype=torch.float64, device=dev)
                n_hint = self._ray_counts()[0] if ne > 0 else 0
                st = lib.rjp_integrate_epochs(d["model"], self._epoch_struct(),
                                              self._continuum_struct(), d["cells"].data_ptr(),
                                              d["extents"].data_ptr(), d["rays"].data_ptr(),
                                              d["n_active_dev"].data_ptr(), int(n_hint or 0),
                                              t_dev.data_ptr(), ne,
                                              None if em is None else em.data_ptr(),
                                              kff.data_ptr(), tsum.data_ptr(), cnt.data_ptr(),
                                              self._cell_grid_ptrs()[0], self._stream())
                _cabi.check(st, "rjp_integrate_epochs")
                ptrs = {k: (out.data_ptr() if k == want else None)
                        for k in ('tau', 'intensity', 'flux')}
                st = lib.rjp_continuum_images_epochs(kff.data_ptr(), ne, tsum.data_ptr(),
                                                     cnt.data_ptr(), npix, coeff.data_ptr(),
                                                     coeff.data_ptr() + nf * 8,
                                                     self._pixel_solid_angle() / 1e-26, nf,
                                                     ptrs['tau'], ptrs['intensity'],
                                                     ptrs['flux'], self._stream())
                _cabi.check(st, "rjp_continuum_images_epochs")
                _launched(2)
            if not self._validate_fill():      # (a host-resolved near-tie changed the state)
                break
        return (out, em) if with_em else out

    def _sky_sums(self):
        """Tile-sharded models: the four ray sums of the WHOLE sky (EM, K, sum T, count), from
        one exchange of the slabs' tiles (29 MB at 1024^2 rays) -- every per-frequency image is
        then formed locally instead of being gathered (16 frequencies x 2 products: 268 MB)."""
        torch = _torch()
        c = self._pass()
        if c.get("sky") is None:
            nxs, nz = self._x_hi - self._x_lo, self._nz
            pack = torch.stack([c["em"], c["kff"], c["tsum"],
                                c["cnt"].to(torch.float64)]).view(4, nxs, nz)
            full = gather_x(pack, self._nx, self._rank, self._world, dim=1, bounds=self._bounds)
            if hasattr(full, "contiguous"):
                full = full.contiguous()
            full = full.reshape(4, self._nx * nz)
            c["sky"] = {"em": full[0], "kff": full[1], "tsum": full[2],
                        "cnt": full[3].to(torch.int32)}
        return c["sky"]

    def _continuum_images(self, freqs, want):
        out = self._continuum_images_device(freqs, want)
        return self._host_image(out, lead=out.shape[0])

    def los_means(self):
        """Line-of-sight means of the cell properties, what the reference's model plot draws
        (plotting/functions.py:539-590) -- without materialising any 3-D grid:
        {'number_density', 'temperature', 'ion_fraction', 'v_los'} -> (nx, nz) arrays equal to
        np.nanmean(<property>, axis=1) (v_los relative to v_lsr, km/s), plus the scalars
        'n_min', 'n_max', 't_max' = nanmin / nanmax over the whole grid."""
        torch = _torch()
        lib = _cabi.load()
        if self._overrides:
            raise NotImplementedError("los_means() evaluates the model's own power laws; it does "
                                      "not see user-assigned grids")
        d = self._ensure_filled()
        dev = d["device"]
        npix = (self._x_hi - self._x_lo) * self._nz
        with torch.cuda.device(dev):
            out = torch.empty((7, npix), dtype=torch.float64, device=dev)
            st = lib.rjp_los_means(d["model"], self._epoch_struct(), d["nverts"].data_ptr(),
                                   d["extents"].data_ptr(), d["rays"].data_ptr(),
                                   d["n_active_dev"].data_ptr(), out.data_ptr(), self._stream())
            _cabi.check(st, "rjp_los_means")
            _launched()
        maps = self._host_image(out, lead=7)
        if maps is None:
            return None
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                return {"number_density": maps[0], "temperature": maps[1],
                        "ion_fraction": maps[2], "v_los": maps[3],
                        "n_min": float(np.nanmin(maps[4])), "n_max": float(np.nanmax(maps[5])),
                        "t_max": float(np.nanmax(maps[6]))}

    def rt_products(self, cont_freqs=None, line=None, chan_freqs=None, contsub=False,
                    host=True):
        """Batched driver for one epoch (what Pipeline.execute asks of the model per run,
        classes.py:2397-2453, in ONE sweep over the grid): emission measure, continuum
        tau/flux at `cont_freqs`, and -- if `line` is given -- the RRL tau and flux cubes
        at `chan_freqs`.  `host=False` keeps the (x-gathered) results on the device."""
        out = {}
        if line is not None:
            res = self._pass(line, np.asarray(chan_freqs, np.float64), contsub=contsub)
            nch = len(chan_freqs)
        cont = self._pass()
        conv = self._host_image if host else self._device_image
        out["em"] = conv(self._sky_sums()["em"] if (self._tiles and self._world > 1)
                         else cont["em"])
        if cont_freqs is not None:
            nf = len(np.atleast_1d(cont_freqs))
            out["tau_ff"] = conv(self._continuum_images_device(cont_freqs, 'tau'), lead=nf)
            out["flux_ff"] = conv(self._continuum_images_device(cont_freqs, 'flux'), lead=nf)
        if line is not None:
            if host:
                out["tau_rrl"] = self._host_cube(res["tau"], 0.0, nch, sibling=res["flux"])
                out["flux_rrl"] = self._host_cube(res["flux"], float("nan"), nch)
            elif self._chan_world > 1:
                # channel-sharded: the planes [c_lo, c_hi) this rank integrated stay with it
                out["tau_rrl"], out["flux_rrl"] = res["tau"], res["flux"]
                out["channels"] = (res["c_lo"], res["c_hi"])
            elif self._tiles and self._world > 1:
                # sky tiles: rows [x_lo, x_hi) of every plane stay with this rank
                out["tau_rrl"], out["flux_rrl"] = res["tau"], res["flux"]
                out["rows"] = (self._x_lo, self._x_hi)
            else:
                out["tau_rrl"] = conv(res["tau"], lead=nch)
                out["flux_rrl"] = conv(res["flux"], lead=nch)
        return out

    def _ray_meta(self):
        from . import sharding
        d = self._dev
        if d.get("ray_meta") is None:
            d["ray_meta"] = sharding.build_ray_meta(d["extents"], d["rays"][:self._n_active()],
                                                    self._x_lo, self._nx, self._nz, self._rank,
                                                    self._world, bounds=self._bounds)
        return d["ray_meta"]

    def _fill_remote_constants(self, tau, flux):
        """Sharded line pass: the constants (tau = 0, flux = NaN) of the OTHER slabs' rays that
        miss the jet, from the all-gathered extents, on a side stream so that they are
        written beside this slab's channel loop.  Returns the side stream."""
        torch = _torch()
        from . import sharding
        lib = _cabi.load()
        d = self._dev
        nz, nx = self._nz, self._nx
        meta = self._ray_meta()
        if d.get("stream3") is None:
            d["stream3"] = torch.cuda.Stream(device=d["device"])
        side = d["stream3"]
        side.wait_stream(torch.cuda.current_stream(d["device"]))
        nch = (tau if tau is not None else flux).shape[0]
        # beside a long channel loop a light store grid interferes least; when this slab's loop
        # is shorter than the constant fill itself, the fill should run at full bandwidth
        ncube = (tau is not None) + (flux is not None)
        t_loop = self._n_active() * nch * 1.7e-7                    # ms, measured rate
        t_fill = (nx - (self._x_hi - self._x_lo)) * nz * nch * ncube * 8 / 3.4e9   # ms, light grid
        light = 1 if t_loop > t_fill else 0
        # ONE launch for all other slabs (own rays skipped), queued before the channel loop so
        # that its CTAs are resident beside it: a kernel launched after the loop has filled
        # the SMs starves until the loop's whole CTA queue has drained
        ext = meta["extents"]
        st = lib.rjp_fill_missed(ext.data_ptr(), ext.shape[0], nch, nx * nz, 0,
                                 self._x_lo * nz, self._x_hi * nz,
                                 tau.data_ptr() if tau is not None else None,
                                 flux.data_ptr() if flux is not None else None,
                                 light, side.cuda_stream)
        _cabi.check(st, "rjp_fill_missed")
        _launched()
        return side

    def _exchange_cubes(self, tau, flux, side):
        """Sparse all-gather of the line cubes between x-slabs (sharding.exchange_ray_columns)
        with the C-ABI pack / scatter kernels; the constants of the other slabs were written
        by _fill_remote_constants on `side`."""
        torch = _torch()
        from . import sharding
        lib = _cabi.load()
        d = self._dev
        nz, nx = self._nz, self._nx
        stream = self._stream()
        plane = nx * nz

        class Ops:
            @staticmethod
            def pack(cube, ids, out):
                _cabi.check(lib.rjp_pack_rays(cube.data_ptr(), plane, ids.data_ptr(), ids.numel(),
                                              out.shape[1], out.shape[0], out.data_ptr(),
                                              stream), "rjp_pack_rays")
                _launched()

            @staticmethod
            def scatter(src, ids, cube):
                _cabi.check(lib.rjp_scatter_rays(src.data_ptr(), src.shape[1], ids.data_ptr(),
                                                 ids.numel(), src.shape[0], cube.data_ptr(),
                                                 plane, stream), "rjp_scatter_rays")
                _launched()

            @staticmethod
            def fill_missed(extents, offset, cubes, values):
                pass    # done on the side stream, beside the channel loop

        views = [c.view(c.shape[0], plane) if c is not None else None for c in (tau, flux)]
        sharding.exchange_ray_columns(views, [0.0, float("nan")], self._ray_meta(), nx, nz,
                                      self._rank, self._world, ops=Ops, bounds=self._bounds)
        torch.cuda.current_stream(d["device"]).wait_stream(side)

    def _device_image(self, t, lead=None):
        nxs, nz = self._x_hi - self._x_lo, self._nz
        if self._world > 1 and t.numel() == (lead or 1) * self._nx * nz:
            return t.view((self._nx, nz) if lead is None else (lead, self._nx, nz))
        t = t.view(nxs, nz) if lead is None else t.view(lead, nxs, nz)
        if self._world > 1:
            t = gather_x(t, self._nx, self._rank, self._world, dim=0 if lead is None else 1,
                         bounds=self._bounds)
        return t

    # ------------------------------------------------------------------ public RT methods
    def emission_measure(self, savefits=False):
        """Emission measure viewed along the y-axis [pc cm^-6] (classes.py:1101-1128)"""
        ems = self._host_image(self._sky_sums()["em"] if (self._tiles and self._world > 1)
                               else self._pass()["em"])
        if ems is None:
            return None
        if savefits:
            self.save_fits(reorder_axes(ems, ra_axis=0, dec_axis=1), savefits, 'em')
        return ems

    def optical_depth_ff(self, freq, savefits=False, collapse=True):
        """Free-free optical depth along the y-axis (classes.py:1353-1447)"""
        scalar = np.isscalar(freq)
        if not collapse:
            tff = self._cellwise_tau_ff(np.atleast_1d(freq))
            return tff[0] if scalar else tff
        tff = self._continuum_images(freq, 'tau')
        if tff is None:
            return None
        if scalar:
            tff = tff[0]
        if savefits:
            self._save_image(tff, savefits, 'tau', freq, scalar)
        return tff

    def intensity_ff(self, freq, savefits=False):
        """Intensity along the y-axis [W m^-2 Hz^-1 sr^-1] (classes.py:1449-1496)"""
        scalar = np.isscalar(freq)
        ints = self._continuum_images(freq, 'intensity')
        if ints is None:
            return None
        if scalar:
            ints = ints[0]
        if savefits:
            self._save_image(ints, savefits, 'intensity', freq, scalar)
        return ints

    def flux_ff(self, freq, savefits=False):
        """Flux [Jy/pixel] (classes.py:1498-1541)"""
        scalar = np.isscalar(freq)
        fluxes = self._continuum_images(freq, 'flux')
        if fluxes is None:
            return None
        if scalar:
            fluxes = fluxes[0]
        if savefits:
            self._save_image(fluxes, savefits, 'flux', freq, scalar)
        return fluxes

    def optical_depth_rrl(self, rrl, freq, lte=True, savefits=False, collapse=True):
        """RRL optical depth along the y-axis (classes.py:1130-1229)"""
        scalar = np.isscalar(freq)
        freqs = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        if not collapse:
            tau = self._cellwise_tau_rrl(rrl, freqs)
            return tau[0] if scalar else tau
        res = self._pass(rrl, freqs, contsub=self._line_contsub_hint(), want_tau=True,
                         want_flux=True)
        tau = self._host_cube(res["tau"], 0.0, freqs.size, sibling=res["flux"])
        if tau is None:
            return None
        if scalar:
            tau = tau[0]
        if savefits:
            self._save_image(tau, savefits, 'tau', freq, scalar)
        return tau

    def _line_contsub_hint(self):
        # Pipeline asks for tau then for flux with contsub=False (classes.py:2437-2453):
        # computing that flux cube in the same pass avoids a second sweep.
        return False

    def intensity_rrl(self, rrl, freq, lte=True, savefits=False):
        """Line intensity [W m^-2 Hz^-1 sr^-1] (classes.py:1231-1290)"""
        if not lte:
            raise ValueError("Non-LTE RRL calculations not yet supported")
        scalar = np.isscalar(freq)
        freqs = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        res = self._pass(rrl, freqs, contsub=True, want_tau=True, want_flux=True)
        ints = self._host_cube(res["flux"], float("nan"), freqs.size)
        if ints is None:
            return None
        ints = ints * (1e-26 / self._pixel_solid_angle())
        if scalar:
            ints = ints[0]
        if savefits:
            self._save_image(ints, savefits, 'intensity', freq, scalar)
        return ints

    def flux_rrl(self, rrl, freq, lte=True, contsub=True, savefits=False):
        """RRL flux [Jy/pixel]; contsub=False adds the continuum (classes.py:1292-1351)"""
        if not lte:
            raise ValueError("Non-LTE RRL calculations not yet supported")
        scalar = np.isscalar(freq)
        freqs = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        res = self._pass(rrl, freqs, contsub=contsub, want_tau=True, want_flux=True)
        fluxes = self._host_cube(res["flux"], float("nan"), freqs.size)
        if fluxes is None:
            return None
        if scalar:
            fluxes = fluxes[0]
        if savefits:
            self._save_image(fluxes, savefits, 'flux', freq, scalar)
        return fluxes

    def rrl_flux_totals(self, rrl, freq, contsub=False, host=True):
        """Sky-summed flux of every channel [Jy], i.e. what Pipeline stores as a line run's
        `results['flux']` (np.nansum over both sky axes, classes.py:2468-2472), reduced on the
        device.  Sharded models all-gather the per-channel sums (the only collective of a
        channel-sharded run), so every rank gets all `len(freq)` values."""
        torch = _torch()
        freqs = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        res = self._pass(rrl, freqs, contsub=contsub, want_tau=True, want_flux=True)
        flux = res["flux"]
        d = self._dev
        lib = _cabi.load()
        nloc = flux.shape[0]
        local = torch.zeros(nloc, dtype=torch.float64, device=flux.device)
        full = self._world > 1 and not self._tiles          # full-size cube, own rows only
        plane = flux.numel() // max(nloc, 1)
        if nloc > 0:
            with torch.cuda.device(flux.device):
                st = lib.rjp_column_totals(flux.data_ptr(), plane,
                                           self._x_lo * self._nz if full else 0,
                                           d["rays"].data_ptr(), d["n_active_dev"].data_ptr(),
                                           nloc, local.data_ptr(), self._stream())
                _cabi.check(st, "rjp_column_totals")
                _launched()
        if self._chan_world > 1:
            from .sharding import gather_channel_totals
            local = gather_channel_totals(local, freqs.size, self._chan_rank, self._chan_world)
        elif self._world > 1:
            import torch.distributed as dist
            dist.all_reduce(local, op=dist.ReduceOp.SUM)      # the slabs partition the sky
        return local.cpu().numpy() if host else local

    def _cellwise_tau_rrl(self, rrl, freqs):
        """`optical_depth_rrl(..., collapse=False)` (classes.py:1159-1189, rrls.py:329-389):
        un-summed (nfreq, nx, ny, nz) line optical depths, composed on the host from the fp64
        field planes with scipy's wofz (never used by Pipeline; not a hot path)."""
        from scipy.special import wofz
        element, n, dn = hm.rrl_parser(rrl)
        with np.errstate(all='ignore'):
            rest = hm.rrl_nu_0(element, n, dn) * (1. - self.vel[1] * 1000. / con.c)
            n_es = self.number_density * self.ion_fraction
            t = self.temperature
            fwhm_g = hm.deltanu_g(rest, t, element)
            fwhm_l = hm.deltanu_l(n_es, n, dn)
            sigma = fwhm_g / 2. / np.sqrt(2. * np.log(2))
            n_i = hm.ni_from_ne(n_es, element)
            path = self._csize * con.au * 1e2 * (self.fill_factor / self.areas)
            fn, en, z = hm.f_n1n2(n, dn), hm.energy_n(n, element), hm.z_number(element)
            out = np.empty((len(freqs),) + n_es.shape)
            for i, f in enumerate(freqs):
                phi = np.real(wofz(((f - rest) + 1j * fwhm_l / 2.) / sigma / np.sqrt(2.))) / \
                    sigma / np.sqrt(2. * np.pi)
                out[i] = 1.0991132675738456e-17 * (n ** 2. * fn * phi) * \
                    (n_es * n_i / t ** 1.5) * np.exp((z ** 2. * en) / (hm.k_cgs * t)) * \
                    (1. - np.exp(-hm.h_cgs * f / (hm.k_cgs * t))) * path
        return out

    def _cellwise_tau_ff(self, freqs):
        """collapse=False branch of optical_depth_ff (classes.py:1383, :1395-1397): the
        un-summed 3-D optical depths, composed on the host from the fp64 field planes
        (never used by Pipeline; not a hot path)."""
        n_es = self.number_density * self.ion_fraction
        t = self.temperature
        path = self._csize * con.au * 1e2 * (self.fill_factor / self.areas)
        out = np.empty((len(freqs),) + n_es.shape)
        with np.errstate(all='ignore'):
            for i, nu in enumerate(freqs):
                if self._params['power_laws']['q_T'] == 0.:
                    g = hm.gff(float(nu), self._params['properties']['T_0'])
                else:
                    g = 11.95 * t ** 0.15 * nu ** -0.1
                out[i] = 0.018 * t ** -1.5 * nu ** -2. * n_es ** 2. * path * g
        return out

    # ------------------------------------------------------------------ products
    def _save_image(self, data, filename, image_type, freq, scalar):
        if scalar:
            self.save_fits(reorder_axes(data, ra_axis=0, dec_axis=1), filename, image_type,
                           freq)
        else:
            self.save_fits(reorder_axes(data, ra_axis=1, dec_axis=2, axis3=0,
                                        axis3_type='freq'), filename, image_type, freq)

    def save_fits(self, data, filename, image_type, freq=None):
        """Write a FITS image/cube with the reference's header (classes.py:1543-1652)."""
        from .fitsio import write_model_fits
        if image_type not in ('flux', 'tau', 'em', 'intensity'):
            raise ValueError("arg image_type must be one of 'flux', 'tau' or 'em'")
        write_model_fits(self, data, filename, image_type, freq)
        return None

    def save(self, filename):
        """Pickle params / fill factors / areas / time / log (classes.py:1704-1713)"""
        filled = self._dev is not None
        ps = {'params': self._params,
              'areas': self.areas if filled else None,
              'ffs': self.fill_factor if filled else None,
              'time': self.time,
              'log': self.log}
        self.log.add_entry("INFO", "Saving physical model to {}".format(filename))
        with open(filename, "wb") as f:
            pickle.dump(ps, f)
        return None


def flux_ff_time_series(params, epochs_s, freq, rank=0, world=1, device=None, log=None,
                        host=True):
    """Continuum flux images of a variable-ejection time series (BASELINE config 4; what
    Pipeline does per run year, classes.py:2347-2453): (n_epochs, nx, nz) at frequency `freq`
    for the model times `epochs_s` [s].  Sharded by EPOCH: every rank holds the whole grid
    (epochs only change the burst factor chi(t), the filled state is reused), integrates the
    epochs `sharding.epoch_shares` deals to it in ONE batched ray walk
    (`JetModel._continuum_epochs_device`) and the images are all-gathered."""
    from .sharding import epoch_shares, gather_epochs
    torch = _torch()
    epochs_s = np.asarray(epochs_s, dtype=np.float64)
    jm = JetModel(params, log=log, device=device)
    mine = list(epoch_shares(len(epochs_s), rank, world))
    dev = jm._device()
    npix = jm.nx * jm.nz
    if mine:
        local = jm._continuum_epochs_device(epochs_s[mine], float(freq), 'flux')[:, 0]
    else:
        local = torch.empty((0, npix), dtype=torch.float64, device=dev)
    full = gather_epochs(local, len(epochs_s), rank, world).view(len(epochs_s), jm.nx, jm.nz)
    jm.release()
    return _to_host(full) if host else full


def _to_host(t):
    """Device tensor -> numpy through a pinned staging buffer."""
    torch = _torch()
    if hasattr(t, "to_host"):          # folded all-gather view (sharding._FoldedView)
        return t.to_host().numpy()
    if t.device.type != "cuda":
        return t.contiguous().numpy()
    t = t.contiguous()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def reorder_axes(data, ra_axis, dec_axis, axis3=None, axis4=None, axis3_type=None,
                 axis4_type=None):
    """Copy of `data` with axes ordered (..., dec, ra) as FITS wants them
    (miscellaneous/functions.py:236-301)."""
    cur = {'ra': ra_axis, 'dec': dec_axis}
    req = {'ra': 1, 'dec': 0}
    if axis3 is not None:
        cur[axis3_type] = axis3
        req = {k: v + 1 for k, v in req.items()}
        req[axis3_type] = 0
        if axis4 is not None:
            cur[axis4_type] = axis4
            req = {k: v + 1 for k, v in req.items()}
            req[axis4_type] = 0
    order = [None] * len(req)
    for k, pos in req.items():
        order[pos] = cur[k]
    return np.ascontiguousarray(np.transpose(np.asarray(data), order))


def check_model_params(params):
    """Structural validation of a model parameter dict; returns an exception instance
    (or None) like miscellaneous/functions.py:127-190 does."""
    if not isinstance(params, dict):
        return TypeError("model params must be dict")
    need = {'target': ('name', 'ra', 'dec', 'epoch', 'dist', 'v_lsr', 'M_star', 'R_1', 'R_2'),
            'grid': ('n_x', 'n_y', 'n_z', 'l_z', 'c_size'),
            'geometry': ('epsilon', 'opang', 'w_0', 'r_0', 'inc', 'pa', 'rotation'),
            'power_laws': ('q_v', 'q_T', 'q_x', 'q^d_n', 'q^d_T', 'q^d_v', 'q^d_x'),
            'properties': ('v_0', 'x_0', 'T_0', 'mu', 'mlr_bj', 'mlr_rj'),
            'ejection': ('t_0', 'hl', 'chi', 'which')}
    for section, keys in need.items():
        if section not in params:
            return KeyError("{} keyword not found in params dict".format(section))
        for key in keys:
            if key not in params[section]:
                return KeyError("{} keyword not found in {} section of params "
                                "dict".format(key, section))
    n = len(params['ejection']['t_0'])
    for key in ('hl', 'chi', 'which'):
        if len(params['ejection'][key]) != n:
            return ValueError("ejection arrays must have equal lengths")
    return None
