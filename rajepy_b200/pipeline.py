"""
Radiative-transfer block of the reference's `Pipeline` as a driver around the CUDA
`JetModel`: the run list (`ContinuumRun` / `RRLRun`,
classes.py:1715-1967, built as `Pipeline.__init__` does, :2116-2172) and `run_rt`, which is
`Pipeline.execute`'s per-run RT section (:2386-2479) without plots and CASA: same product
directories and file names, the same skip-if-exists / clobber / resume rules, the same
`results['flux']` reductions and the same model / pipeline save files, so that `-rt` keeps its
on-disk layout.  Everything numerical goes through `JetModel`; all runs of one epoch share
one grid fill and one line-of-sight pass per (line, channel set).
"""
import os
import pickle
from collections.abc import Iterable

import numpy as np
import scipy.constants as con

from . import hostmath as hm


def freq_str(freq, fmt='.0f'):
    """'5GHz', '33GHz', ... (miscellaneous/functions.py:193-233); names the product
    directories of continuum runs."""
    suffixes = (('Hz', 1., 1e3), ('kHz', 1e3, 1e6), ('MHz', 1e6, 1e9), ('GHz', 1e9, 1e12),
                ('THz', 1e12, 1e15), ('PHz', 1e15, 1e18))

    def one(f):
        for name, lo, hi in suffixes:
            if lo <= f < hi:
                return f'{{:{fmt}}}{{}}'.format(f / lo, name)
        raise ValueError(f"frequency {f} Hz outside 1 Hz .. 1e18 Hz")

    if isinstance(freq, Iterable):
        return [one(f) for f in freq]
    return one(freq)


class ContinuumRun:
    """One (epoch, frequency) radiative-transfer / synthetic-observation job
    (classes.py:1715-1900).  Attribute names follow the reference, so run lists pickled by
    the reference's `Pipeline.save` load into this class (`compat.load_pickle`)."""

    def __init__(self, dcy, year, freq=None, bandwidth=None, chanwidth=None, t_obs=None,
                 t_int=None, tscop=None):
        self._year = year
        self._dcy = dcy
        self._obs_type = 'continuum'
        self._freq = freq
        self._t_obs = t_obs
        self._t_int = t_int
        self._tscop = tscop
        self._products = {}
        self._results = {}
        # bandwidth / channel width default to 1 Hz (classes.py:1736-1745)
        self._bandwidth = bandwidth if bandwidth is not None else 1.
        self._chanwidth = chanwidth if chanwidth is not None else 1.
        self.completed = False
        self.radiative_transfer = freq is not None
        self.simobserve = all(v is not None for v in (tscop, bandwidth, chanwidth, t_obs, t_int))

    line = None

    def __str__(self):
        tscop = self._tscop if self._tscop is None else tuple(self._tscop)
        vals = (('Year [yr]', format(self._year, '.2f')),
                ('Type', self._obs_type.capitalize()),
                ('Telescope', '-' if tscop is None else str(tscop)),
                ('t_obs [s]', '-' if self._t_obs is None else format(self._t_obs, '.0f')),
                ('t_int [s]', '-' if self._t_int is None else format(self._t_int, '.0f')),
                ('Line', '-' if self.line is None else self.line),
                ('Frequency [Hz]', '-' if self._freq is None else format(self._freq, '.3e')),
                ('Bandwidth [Hz]', format(self._bandwidth, '.3e')),
                ('Channel width [Hz]', format(self._chanwidth, '.3e')),
                ('Radiative Transfer?', str(self.radiative_transfer)),
                ('Synthetic Obs.?', str(self.simobserve)),
                ('Completed?', str(self.completed)))
        return ', '.join(f'{k}: {v}' for k, v in vals)

    @property
    def results(self):
        return self._results

    @results.setter
    def results(self, new_results):
        if not isinstance(new_results, dict):
            raise TypeError("setter method for results attribute requires dict")
        self._results = new_results

    @property
    def products(self):
        return self._products

    @products.setter
    def products(self, new_products):
        if not isinstance(new_products, dict):
            raise TypeError("setter method for products attribute requires dict")
        self._products = new_products

    @property
    def obs_type(self):
        return self._obs_type

    @property
    def dcy(self):
        return self._dcy

    @dcy.setter
    def dcy(self, path):
        self._dcy = path

    @property
    def model_dcy(self):
        return os.sep.join([self.dcy, f'Day{self.day}'])

    @property
    def rt_dcy(self):
        if not self.radiative_transfer:
            return None
        return os.sep.join([self.model_dcy, self._tag()])

    def _tag(self):
        return freq_str(self.freq)

    @property
    def year(self):
        return self._year

    @property
    def day(self):
        return int(self.year * 365.)

    @property
    def freq(self):
        return self._freq

    @property
    def bandwidth(self):
        return self._bandwidth

    @property
    def chanwidth(self):
        return self._chanwidth

    @property
    def t_obs(self):
        return self._t_obs

    @property
    def t_int(self):
        return self._t_int

    @property
    def tscop(self):
        return self._tscop

    def _fits(self, kind):
        return self.rt_dcy + os.sep + '_'.join([kind, 'Day' + str(self.day), self._tag()]) + '.fits'

    @property
    def fits_flux(self):
        return self._fits('Flux')

    @property
    def fits_tau(self):
        return self._fits('Tau')

    @property
    def fits_em(self):
        return self._fits('EM')

    @property
    def nchan(self):
        return int(self.bandwidth / self.chanwidth)

    @property
    def chan_freqs(self):
        """classes.py:1897-1900"""
        chan1 = self.freq - self.bandwidth / 2. + self.chanwidth / 2.
        return chan1 + np.arange(self.nchan) * self.chanwidth


class RRLRun(ContinuumRun):
    """classes.py:1903-1967: centred on the (un-shifted) rest frequency of `line`."""

    def __init__(self, dcy, year, line=None, bandwidth=None, chanwidth=None, t_obs=None,
                 t_int=None, tscp=None):
        self.line = line
        freq = hm.rrl_nu_0(*hm.rrl_parser(line))
        super().__init__(dcy, year, freq, bandwidth, chanwidth, t_obs, t_int, tscp)
        self._obs_type = 'rrl'

    def _tag(self):
        return self.line


def _pick(v, idx):
    return v[idx] if isinstance(v, Iterable) and not isinstance(v, str) else v


def build_runs(dcy, params):
    """Run list of a pipeline parameter dict, in the reference's order (classes.py:2116-2172):
    continuum runs (times sorted, frequencies inner), then RRL runs."""
    dcy = dcy.rstrip(os.sep)
    runs = []
    for band, cls in (('continuum', ContinuumRun), ('rrls', RRLRun)):
        sec = params[band]
        times = sec['times']
        times = np.array([]) if times is None else np.sort(np.asarray(times))
        what = sec['freqs'] if band == 'continuum' else sec['lines']
        for t in times:
            for i, item in enumerate(what):
                runs.append(cls(dcy, t, item, _pick(sec['bws'], i), _pick(sec['chanws'], i),
                                _pick(sec['t_obs'], i), _pick(sec['t_ints'], i),
                                _pick(sec['tscps'], i)))
    return runs


def total_flux(run, fluxes):
    """`results['flux']` of a run (classes.py:2461-2472): continuum -> sum over the sky of the
    channel-averaged flux; RRL -> per-channel sums over the sky."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        if run.obs_type == 'continuum':
            return np.nansum(np.nanmean(fluxes, axis=0))
        return np.nansum(np.nansum(fluxes, axis=1), axis=1)


def save_pipeline(save_file, runs, params, model_file, log=None, absolute_directories=False):
    """Pickle of the pipeline state with the reference's keys (classes.py:2215-2258)."""
    home = os.path.expanduser('~')
    mf = model_file
    if not absolute_directories:
        for run in runs:
            run.dcy = run.dcy.replace(home, '~')
        params['dcys']['model_dcy'] = params['dcys']['model_dcy'].replace(home, '~')
        mf = mf.replace(home, '~')
    if log is not None:
        log.add_entry(mtype="INFO", entry="Saving pipeline to " + save_file)
    with open(save_file, 'wb') as f:
        pickle.dump({"runs": runs, "params": params, "model_file": mf, 'log': log}, f)


def load_pipeline(load_file):
    """(runs, params, model_file, log) of a pipeline save file written by `save_pipeline` or by
    the reference's `Pipeline.save` (classes.py:1976-2017)."""
    from .compat import load_pickle
    home = os.path.expanduser('~')
    loaded = load_pickle(os.path.expanduser(load_file))
    for run in loaded['runs']:
        run.dcy = run.dcy.replace('~', home)
    loaded['model_file'] = loaded['model_file'].replace('~', home)
    loaded['params']['dcys']['model_dcy'] = \
        loaded['params']['dcys']['model_dcy'].replace('~', home)
    return loaded['runs'], loaded['params'], loaded['model_file'], loaded.get('log')


def run_rt(model, runs, params=None, clobber=False, resume=False, dryrun=False,
           model_file=None, save_file=None):
    """The radiative-transfer section of `Pipeline.execute` (classes.py:2386-2479) for the
    runs in `runs` (from `build_runs` / `load_pipeline`), against a `rajepy_b200.JetModel`:

      * model time = run.year, product directory `Day<d>/<freq|line>/` created on demand;
      * `EM_`, `Tau_`, `Flux_` FITS files written unless they exist (then the flux cube is
        read back for the totals) or `clobber`; RRL fluxes with `contsub=False`;
      * `run.results['flux']` as the reference reduces it; a completed run is skipped when
        `resume` and not `clobber`;
      * the model (and, if `params` / `save_file` are given, the pipeline state) is saved
        after every successful run like the reference does.

    Runs of one epoch share the grid fill and the line-of-sight passes (JetModel caches them
    per model time).  Returns the list of runs."""
    from .fitsio import read_fits_data
    log = model.log
    dcy = runs[0].dcy if runs else None
    if model_file is None and dcy is not None:
        model_file = dcy + os.sep + "jetmodel.save"
    for idx, run in enumerate(runs):
        model.time = run.year * con.year
        log.add_entry(mtype="INFO", entry="Executing run #{} -> Details:\n{}"
                                          "".format(idx + 1, run.__str__()))
        if run.completed and resume and not clobber:
            log.add_entry(mtype="INFO", entry="Run #{} previously completed, skipping"
                                              "".format(idx + 1), timestamp=False)
            continue
        if not run.radiative_transfer:
            run.completed = True
            continue
        if not os.path.exists(run.rt_dcy):
            log.add_entry(mtype="INFO", entry="{} doesn't exist, creating".format(run.rt_dcy),
                          timestamp=False)
            os.makedirs(run.rt_dcy)
        if dryrun:
            run.completed = True
            continue
        log.add_entry(mtype="INFO",
                      entry="Conducting radiative transfer at "
                            f"{run.freq / 1e9:.1f}GHz for a model time of {run.year:.1f}yr")
        if not os.path.exists(run.fits_em) or clobber:
            log.add_entry(mtype="INFO", entry=f"Emission measures saved to {run.fits_em}")
            model.emission_measure(savefits=run.fits_em)
        else:
            log.add_entry(mtype="INFO", entry=f"Emission measures already exist -> {run.fits_em}",
                          timestamp=False)
        cont = run.obs_type == 'continuum'
        if not os.path.exists(run.fits_tau) or clobber:
            log.add_entry(mtype="INFO",
                          entry=f"Computing optical depths and saving to {run.fits_tau}")
            if cont:
                model.optical_depth_ff(run.chan_freqs, savefits=run.fits_tau)
            else:
                model.optical_depth_rrl(run.line, run.chan_freqs, savefits=run.fits_tau)
        else:
            log.add_entry(mtype="INFO", entry=f"Optical depths already exist -> {run.fits_tau}",
                          timestamp=False)
        if not os.path.exists(run.fits_flux) or clobber:
            log.add_entry(mtype="INFO",
                          entry=f"Calculating fluxes and saving to {run.fits_flux}")
            if cont:
                fluxes = model.flux_ff(run.chan_freqs, savefits=run.fits_flux)
            else:
                fluxes = model.flux_rrl(run.line, run.chan_freqs, contsub=False,
                                        savefits=run.fits_flux)
        else:
            log.add_entry(mtype="INFO", entry=f"Fluxes already exist -> {run.fits_flux}",
                          timestamp=False)
            fluxes = read_fits_data(run.fits_flux)
        flux = total_flux(run, fluxes)
        if cont:
            log.add_entry(mtype="INFO",
                          entry=f"Total, average, channel flux of {flux:.2e}Jy calculated")
        run.results['flux'] = flux
        run.products.update({'em': run.fits_em, 'tau': run.fits_tau, 'flux': run.fits_flux})
        if model_file is not None and not os.path.exists(model_file):
            model.save(model_file)
        # like the reference, the state is saved BEFORE the run is marked complete
        # (classes.py:2474-2479 vs :2853)
        if params is not None and save_file is not None:
            save_pipeline(save_file, runs, params, model_file, log=log,
                          absolute_directories=True)
        run.completed = True
    return runs
