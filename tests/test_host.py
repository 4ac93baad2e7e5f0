"""CPU tests of the host side: derived scalars, the C-ABI library and its header, FITS
products, parameter files, sharding helpers.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest
import scipy.constants as con

import rajepy_b200 as rb
from rajepy_b200 import _cabi, fitsio, hostmath as hm, sharding
from oracle import rajepy_oracle as orc
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(params, **kw):
    log = rb.logger.Log(os.path.join(tempfile.mkdtemp(), "m.log"), verbose=False)
    return rb.JetModel(params, log=log, **kw)


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    hdr = open(os.path.join(ROOT, "include", "rajepy_b200.h")).read()
    declared = set(re.findall(r"\b(rjp_[a-z_]+)\s*\(", hdr))
    assert {"rjp_fill_grid", "rjp_patch_cells", "rjp_cell_field", "rjp_integrate",
            "rjp_continuum_images", "rjp_strerror"} <= declared
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert lib.rjp_abi_version() == _cabi.ABI_VERSION == 6
    assert lib.rjp_strerror(0) == b"ok"
    assert lib.rjp_strerror(-1) == b"invalid argument"


def test_abi_rejects_bad_arguments_without_gpu():
    lib = _cabi.load()
    m = _cabi.Model()  # all zero: invalid dims
    assert lib.rjp_fill_grid(m, None, None, None, None, None, 0, None, None, None) == _cabi.ERR_ARG
    assert lib.rjp_brick_count(m) < 0
    assert lib.rjp_voigt_profile(None, None, 5, None, None) == _cabi.ERR_ARG
    assert lib.rjp_continuum_images(None, None, None, 0, None, None, 1.0, 0, None, None,
                                    None, None) == _cabi.ERR_ARG


def test_abi_rejects_bad_epoch_batches_without_gpu():
    lib = _cabi.load()
    m, ep, ct = _cabi.Model(), _cabi.Epoch(), _cabi.Continuum()
    assert lib.rjp_integrate_epochs(m, ep, ct, None, None, None, None, 0, None, 4, None, None,
                                    None, None, None, None) == _cabi.ERR_ARG
    assert lib.rjp_continuum_images_epochs(None, 3, None, None, 10, None, None, 1.0, 1, None,
                                           None, None, None) == _cabi.ERR_ARG


def _assemble(lib, nchan, plane, ids, cols, fill, threads):
    out = np.full((nchan, plane), 123.0)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.float64)
    st = lib.rjp_host_assemble(out.ctypes.data, nchan, plane, ids.ctypes.data, ids.size,
                               cols.ctypes.data, cols.shape[1] if cols.ndim == 2 else 0,
                               float(fill), threads)
    return st, out


@pytest.mark.parametrize("fill", [0.0, np.nan])
def test_host_assemble_builds_dense_cubes_from_packed_columns(fill):
    """rjp_host_assemble (host threads, no CUDA call): the dense (nchan, nx*nz) product equals
    constants + scattered columns, for ragged run patterns, odd alignments and any thread
    count -- what optical_depth_rrl / flux_rrl hand to numpy (classes.py:1215-1229, :1340-1351)."""
    lib = _cabi.load()
    rng = np.random.default_rng(5)
    for plane, nchan, frac in ((64, 1, 0.5), (1000, 7, 0.06), (4098, 33, 0.3), (130, 5, 1.0),
                               (2, 3, 0.5), (777, 16, 0.0)):
        n = int(round(frac * plane))
        ids = np.sort(rng.choice(plane, size=n, replace=False))
        if n > 8:                                   # make some long runs of neighbouring rays
            ids[: n // 2] = ids[0] + np.arange(n // 2)
            ids = np.unique(ids)
            n = ids.size
        cols = rng.normal(size=(nchan, max(n, 1)))
        if n:
            cols[0, 0] = np.nan                     # data may hold NaN too
        want = np.full((nchan, plane), fill)
        want[:, ids] = cols[:, :n]
        for threads in (1, 3, 16):
            st, got = _assemble(lib, nchan, plane, ids, cols, fill, threads)
            assert st == _cabi.OK
            assert np.array_equal(got, want, equal_nan=True), (plane, nchan, threads)


def test_host_assemble_rejects_bad_ray_lists():
    lib = _cabi.load()
    cols = np.zeros((2, 3))
    for ids in ([3, 2, 5], [1, 1, 2], [0, 4, 10], [-1, 2, 3]):      # not ascending / outside
        st, _ = _assemble(lib, 2, 10, ids, cols, 0.0, 2)
        assert st == _cabi.ERR_ARG
    assert lib.rjp_host_assemble(None, 1, 4, None, 0, None, 0, 0.0, 1) == _cabi.ERR_ARG


def test_struct_mirrors_match_header_sizes():
    lib = _cabi.load()
    sizes = [ctypes.c_int32() for _ in range(6)]
    lib.rjp_struct_sizes(*[ctypes.byref(s) for s in sizes])
    mirrors = (_cabi.Model, _cabi.Epoch, _cabi.Continuum, _cabi.Line, _cabi.Channels)
    assert [s.value for s in sizes[:5]] == [ctypes.sizeof(t) for t in mirrors]
    assert sizes[5].value == 16  # the 16-byte cell state


@pytest.mark.parametrize("name", ["small", "inclined", "powerlaws", "nobursts"])
def test_derived_parameters_match_oracle(name):
    factory = cases.CASES[name][0]
    jm, oj = _model(factory()), orc.OracleJet(factory())
    assert (jm.nx, jm.ny, jm.nz) == (oj.nx, oj.ny, oj.nz)
    for sec, key in (("geometry", "mod_r_0"), ("power_laws", "q_n"),
                     ("power_laws", "q_tau"), ("properties", "n_0")):
        assert jm.params[sec][key] == oj.p[sec][key]
    assert jm.ss_jml('B') == oj.ss_bj and jm.ss_jml('R') == oj.ss_rj
    t = np.linspace(-1, 3, 9) * con.year
    np.testing.assert_allclose(jm.jml_t('B')(t), oj._jml('B', t), rtol=1e-14)
    np.testing.assert_allclose(jm.jml_t('RB')(t), oj._jml('B', t) + oj._jml('R', t),
                               rtol=1e-14)
    ep = jm._epoch_struct()
    assert ep.n_blue == len(oj.bursts['B']) and ep.n_red == len(oj.bursts['R'])


def test_lz_grid_dims_and_param_file(tmp_path):
    p = cases.base_params()
    p["grid"]["l_z"] = 2.
    assert (_model(p).nx, _model(p).ny, _model(p).nz) == (108, 110, 588)
    f = tmp_path / "my-model-params.py"
    f.write_text("import numpy as np\nparams = " + repr(cases.case_small())
                 .replace("array", "np.array") + "\n")
    jm = _model(str(f))
    assert (jm.nx, jm.ny, jm.nz) == (20, 40, 60)
    with pytest.raises(TypeError):
        rb.JetModel(42)
    with pytest.raises(FileNotFoundError):
        rb.JetModel(str(tmp_path / "missing.py"))
    bad = cases.case_small()
    del bad["geometry"]["inc"]
    assert isinstance(rb.check_model_params(bad), KeyError)


def test_hostmath_matches_oracle_scalars():
    assert hm.rrl_nu_0('H', 58, 1) == orc.rrl_nu_0('H', 58, 1)
    assert hm.rrl_nu_0('He', 42, 2) == orc.rrl_nu_0('He', 42, 2)
    assert hm.gff(5e9, 1e4) == pytest.approx(orc.gff(5e9, 1e4), rel=1e-14)
    fr = np.logspace(9, 11.5, 7)
    np.testing.assert_allclose(hm.gff(fr, 1e4), [orc.gff(f, 1e4) for f in fr], rtol=1e-13)
    assert hm.ni_from_ne(1.0, 'H') == orc.ni_from_ne(1.0, 'H')
    assert hm.f_n1n2(58, 1) == orc.f_n1n2(58, 1)
    assert hm.rrl_parser('He42b') == ('He', 42, 2)
    np.testing.assert_array_equal(hm.chan_freqs(1e10, 8e6, 1e6), orc.chan_freqs(1e10, 8e6, 1e6))


def test_n_0_from_mlr_known_answer_of_the_reference_tests():
    """The reference's own known-answer test for this path (test/test_physics.py:38-59):
    n_0_from_mlr against the quadrature of the mass flux over the jet base, 1e-3 relative, for
    the 9 x 9 grid of cross-sectional indices -- here for the product's host math AND the
    oracle (the only golden vectors the reference's tests hold for the path)."""
    from scipy.integrate import quad
    msol = 1.989e30
    mlr, mu, w0, v0, r1, r2 = 1e-6, 1.3, 5.0, 400., 0.5, 5.0
    const = 2. * con.pi * mu * v0 * 1e3 * hm.atomic_mass("H")

    def flux(w, w0_, qnd_, qnv_, r1_, r2_):
        return w * (1. + w * (r2_ - r1_) / (w0_ * r1_)) ** (qnd_ + qnv_)

    for qnd in np.linspace(-2, 2, 9):
        for qnv in np.linspace(-2, 2, 9):
            integral = quad(flux, 0., w0 * con.au,
                            args=(w0 * con.au, qnd, qnv, r1 * con.au, r2 * con.au))[0]
            want = ((mlr * msol / con.year) / (integral * const)) * 1e-6
            for impl in (hm.n_0_from_mlr, orc.n_0_from_mlr):
                got = impl(mlr, v0, w0, mu, qnd, qnv, r1, r2)
                assert abs(got - want) <= 1e-3 * want, (impl.__module__, qnd, qnv, got, want)
    assert hm.atomic_mass("H") == orc.atomic_mass("H")


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    jm = _model(cases.case_small())
    with pytest.raises(rb.EngineError):
        jm.emission_measure()
    with pytest.raises(rb.EngineError):
        _ = jm.fill_factor
    with pytest.raises(ValueError):
        jm.intensity_rrl('H58a', 3e10, lte=False)


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rajepy_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


def test_reorder_axes_semantics():
    a = np.arange(2 * 3).reshape(2, 3).astype(float)          # (nx, nz)
    assert np.array_equal(rb.reorder_axes(a, ra_axis=0, dec_axis=1), a.T)
    c = np.arange(4 * 2 * 3).reshape(4, 2, 3).astype(float)   # (nf, nx, nz)
    out = rb.reorder_axes(c, ra_axis=1, dec_axis=2, axis3=0, axis3_type='freq')
    assert out.shape == (4, 3, 2)
    assert np.array_equal(out, np.swapaxes(c, 1, 2))


def test_fits_round_trip(tmp_path):
    jm = _model(cases.case_small())
    img = np.random.default_rng(0).normal(size=(jm.nx, jm.nz))
    img[0, 0] = np.nan
    f = str(tmp_path / "em.fits")
    jm.save_fits(rb.reorder_axes(img, 0, 1), f, 'em')
    hdr, data = fitsio.read_fits(f)
    assert os.path.getsize(f) % 2880 == 0
    assert data.shape == (jm.nz, jm.nx)
    assert np.array_equal(np.nan_to_num(data), np.nan_to_num(img.T))
    assert hdr['BITPIX'] == -64 and hdr['NAXIS1'] == jm.nx and hdr['NAXIS2'] == jm.nz
    assert hdr['CTYPE1'] == 'RA---TAN' and hdr['CTYPE2'] == 'DEC--TAN'
    assert hdr['CRPIX1'] == jm.nx / 2 + 0.5 and hdr['CRPIX2'] == jm.nz / 2 + 0.5
    assert hdr['BUNIT'] == 'pc cm^-6' and hdr['OBJECT'] == 'test2'
    ra, dec = fitsio.parse_sexagesimal("04:31:34.07736", "+18:08:04.9020")
    assert hdr['CRVAL1'] == pytest.approx(15 * (4 + 31 / 60 + 34.07736 / 3600), abs=1e-12)
    assert hdr['CRVAL2'] == pytest.approx(dec, abs=1e-12)
    cdelt = np.degrees(np.arctan(0.5 * con.au / (120. * con.parsec)))
    assert hdr['CDELT1'] == pytest.approx(-cdelt, rel=1e-14)
    assert ''.join(hdr['HISTORY']).count('JET MODEL') == 1
    # cube with a frequency axis (classes.py:1615-1629)
    fr = np.array([1e9, 2e9, 3e9, 4e9])
    cube = np.zeros((4, jm.nx, jm.nz))
    f2 = str(tmp_path / "flux.fits")
    jm.save_fits(rb.reorder_axes(cube, 1, 2, axis3=0, axis3_type='freq'), f2, 'flux', fr)
    hdr, data = fitsio.read_fits(f2)
    assert data.shape == (4, jm.nz, jm.nx)
    assert hdr['CTYPE3'] == 'FREQ' and hdr['CDELT3'] == 1e9 and hdr['CRPIX3'] == 2.5
    assert hdr['CRVAL3'] == 2e9 + 0.5e9 and hdr['BUNIT'] == 'Jy pixel^-1'
    with pytest.raises(ValueError):
        jm.save_fits(cube, f2, 'nonsense')


def test_model_table_matches_reference_layout():
    s = str(_model(cases.case_small()))
    lines = s.split('\n')
    assert lines[1].strip('/').strip() == 'JET MODEL'
    assert len({len(ln) for ln in lines if ln}) == 1        # rectangular table
    assert '|  epsilon  |' in s and '+0.778' in s and 'BURSTS' in s and '0.50' in s
    assert 'None' in str(_model(cases.case_nobursts()))


def test_slab_bounds_cover_grid():
    for nx in (50, 64, 1024, 7):
        for world in (1, 2, 3, 4, 7):
            if world > nx:
                continue
            b = [sharding.slab_bounds(nx, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nx
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.epoch_shares(64, 3, 8) == list(range(3, 64, 8))
    with pytest.raises(ValueError):
        sharding.slab_bounds(4, 0, 8)


def test_balanced_bounds_tile_the_grid_and_balance_work():
    rng = np.random.default_rng(3)
    for nx, world in ((50, 8), (1024, 8), (17, 17), (64, 3)):
        w = np.zeros(nx)
        c = nx // 2
        w[max(0, c - nx // 16 - 1): c + nx // 16 + 1] = rng.uniform(1, 5, size=len(
            w[max(0, c - nx // 16 - 1): c + nx // 16 + 1]))
        b = sharding.balanced_bounds(w, world)
        assert b[0][0] == 0 and b[-1][1] == nx
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(hi > lo for lo, hi in b)
        if nx >= 8 * world:
            share = np.array([w[lo:hi].sum() for lo, hi in b])
            assert share.max() <= w.sum() / world + w.max() + 1e-9
    z = sharding.balanced_bounds(np.zeros(12), 4)          # no jet at all: near-even split
    assert z[0][0] == 0 and z[-1][1] == 12 and all(2 <= hi - lo <= 4 for lo, hi in z)
    with pytest.raises(ValueError):
        sharding.balanced_bounds(np.ones(3), 4)


def test_balanced_bounds_minimise_the_busiest_slab():
    """cost(slab) = max(L, W) + overlap min(L, W), L = in-jet cells, W = plane_cost x planes;
    the split is the optimum of an exhaustive search on small cases."""
    import itertools
    rng = np.random.default_rng(11)
    for _ in range(60):
        nx, world = int(rng.integers(4, 13)), int(rng.integers(2, 5))
        w = rng.uniform(0, 9, nx) * (rng.random(nx) < 0.4)
        pc, ov = float(rng.uniform(0, 2)), float(rng.choice([0.0, 0.3, 1.0]))
        cum = np.concatenate([[0.0], np.cumsum(w)])

        def cost(lo, hi):
            a, b = cum[hi] - cum[lo], pc * (hi - lo)
            return max(a, b) + ov * min(a, b)

        best = min(max(cost(c[i], c[i + 1]) for i in range(world))
                   for cuts in itertools.combinations(range(1, nx), world - 1)
                   for c in [(0,) + cuts + (nx,)])
        b = sharding.balanced_bounds(w, world, plane_cost=pc, overlap=ov)
        assert b[0][0] == 0 and b[-1][1] == nx and all(hi > lo for lo, hi in b)
        got = max(cost(lo, hi) for lo, hi in b)
        assert got <= best * (1 + 1e-4) + 1e-6, (w, pc, ov, world, b)
    # empty sky is not free: a slab of sky only is as expensive as a share of the jet
    w = np.zeros(1024)
    w[480:544] = 6e4
    b = sharding.balanced_bounds(w, 8, plane_cost=1500.0)
    cost = [w[lo:hi].sum() + 1500.0 * (hi - lo) for lo, hi in b]
    assert max(cost) <= 1.1 * sum(cost) / 8


def test_sharded_models_agree_on_work_balanced_slabs():
    p = cases.with_grid(cases.base_params(), 64, 64, 64)
    slabs = [_model(p, shard=(r, 4)).slab for r in range(4)]
    assert slabs[0][0] == 0 and slabs[-1][1] == 64
    assert all(slabs[i][1] == slabs[i + 1][0] for i in range(3))
    widths = [hi - lo for lo, hi in slabs]
    assert min(widths[1:3]) < min(widths[0], widths[3])      # the jet sits in the middle planes
    even = [_model(p, shard=(r, 4), balance=False).slab for r in range(4)]
    assert even == [(0, 16), (16, 32), (32, 48), (48, 64)]


def test_reynolds_1986_analytic_flux_matches_reference():
    """maths/physics.py:297-374 restated in hostmath against values computed by the unmodified
    reference (tests/golden/r86.npz, tools/make_golden_r86.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "r86.npz"))
    for name in ("small", "inclined", "nobursts", "tgrad"):
        jm = _model(cases.CASES[name][0]())
        for i, f in enumerate(g["freqs"]):
            for j, which in enumerate("RB"):
                for k, ymax in enumerate(g["ymax"]):
                    for m, ymin in enumerate((None, 0.05)):
                        got = hm.flux_expected_r86(jm, float(f), which, float(ymax), ymin)
                        assert got == pytest.approx(g[name][i, j, k, m], rel=1e-12), \
                            (name, f, which, ymax, ymin)
