"""
ctypes binding of include/rajepy_b200.h (the C ABI of the CUDA engine).

The structs below mirror the header field by field; `load()` verifies every mirror
against `rjp_struct_sizes()` and fails loudly when the library is missing -- there is
no CPU fallback on the product path.
"""
import ctypes as C
import os

from . import build as _build

MAX_BURSTS = 16

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4

FIELDS = {name: i for i, name in enumerate(
    ["fill_factor", "areas", "r", "w", "phi", "reff", "travel", "nd_base", "xi", "temp",
     "vx", "vlos", "vz", "chi"])}


class Model(C.Structure):
    _fields_ = (
        [(n, C.c_int32) for n in ("nx", "ny", "nz", "x_lo", "x_hi")] +
        [(n, C.c_double) for n in (
            "cs", "w0", "r0", "mr0", "eps", "ca", "sa", "cb", "sb", "cva", "sva", "cvb",
            "svb", "R1", "R2", "q_n", "q_x", "q_T", "q_v", "qd_n", "qd_x", "qd_T", "qd_v",
            "n0", "x0", "T0", "v0", "f_rb", "gm_over_au", "rot_sign", "v_lsr", "au_m",
            "year_s", "au_cm", "hyp_b", "hyp_c1", "hyp_c2")] +
        [("hyp_degenerate", C.c_int32), ("reserved0", C.c_int32),
         ("need_reff", C.c_int32), ("reserved1", C.c_int32)])


class Burst(C.Structure):
    _fields_ = [("t0", C.c_double), ("amp", C.c_double), ("inv2s2", C.c_double)]


class Epoch(C.Structure):
    _fields_ = [("time", C.c_double), ("n_blue", C.c_int32), ("n_red", C.c_int32),
                ("blue", Burst * MAX_BURSTS), ("red", Burst * MAX_BURSTS)]


class Continuum(C.Structure):
    _fields_ = [("em_scale", C.c_double), ("tau_scale", C.c_double),
                ("t_exponent", C.c_double)]


class Line(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("nu0", "dopp", "width_g", "stark", "kappa0",
                                          "en_over_k", "h_over_k", "dn_max", "chan_dnu0",
                                          "chan_step", "t_common", "tc_sqrt", "tc_boltz",
                                          "tc_hk", "tc_p0")]


class Channels(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dnu", "nu", "cff", "aff", "bnu")]


class EngineError(RuntimeError):
    pass


_lib = None

EXPORTS = ("rjp_strerror", "rjp_last_cuda_error", "rjp_abi_version", "rjp_struct_sizes",
           "rjp_fill_grid", "rjp_patch_cells", "rjp_cell_field", "rjp_ray_list", "rjp_integrate",
           "rjp_continuum_images", "rjp_voigt_profile", "rjp_brick_count", "rjp_pack_rays",
           "rjp_scatter_rays", "rjp_fill_missed", "rjp_los_means", "rjp_override_cells",
           "rjp_ray_list_chunks", "rjp_host_assemble", "rjp_line_scratch_bytes",
           "rjp_column_totals", "rjp_integrate_epochs", "rjp_continuum_images_epochs")

ABI_VERSION = 6      # RJP_ABI_VERSION of include/rajepy_b200.h this binding was written for


def library_path():
    return _build.LIB


def load():
    """Load librajepy_b200.so (building it in-tree if nvcc is available and it is
    stale or missing).  Raises EngineError if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    stale = False
    if os.path.exists(path) and not os.environ.get("RAJEPY_B200_LIB"):
        try:
            stale = _build.have_nvcc() and _build.needs_build()
        except OSError:
            stale = False
    if not os.path.exists(path) or stale or (os.environ.get("RAJEPY_B200_REBUILD") == "1"):
        try:
            _build.build(force=True)
        except Exception as exc:  # noqa: BLE001
            raise EngineError(f"CUDA library {path} is missing and could not be built: "
                              f"{exc}") from exc
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise EngineError(f"cannot load {path}: {exc}") from exc
    for sym in EXPORTS:
        if not hasattr(lib, sym):
            raise EngineError(f"{path} does not export {sym}")
    lib.rjp_strerror.restype = C.c_char_p
    lib.rjp_strerror.argtypes = [C.c_int]
    lib.rjp_last_cuda_error.restype = C.c_char_p
    lib.rjp_abi_version.restype = C.c_int
    if lib.rjp_abi_version() != ABI_VERSION:
        raise EngineError(f"ABI mismatch: {path} implements version {lib.rjp_abi_version()}, "
                          f"this binding expects {ABI_VERSION} (rebuild: python -m "
                          f"rajepy_b200.build --force)")
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.rjp_struct_sizes.argtypes = [C.POINTER(i32)] * 6
    lib.rjp_fill_grid.argtypes = [C.POINTER(Model), vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.rjp_patch_cells.argtypes = [C.POINTER(Model), vp, vp, i32, vp, vp, vp, vp, vp]
    lib.rjp_los_means.argtypes = [C.POINTER(Model), C.POINTER(Epoch), vp, vp, vp, vp, vp, vp]
    lib.rjp_los_means.restype = C.c_int
    lib.rjp_override_cells.argtypes = [C.POINTER(Model), vp, i32, vp, vp, vp]
    lib.rjp_override_cells.restype = C.c_int
    lib.rjp_brick_count.argtypes = [C.POINTER(Model)]
    lib.rjp_brick_count.restype = i64
    lib.rjp_cell_field.argtypes = [C.POINTER(Model), C.POINTER(Epoch), vp, i32, vp, vp]
    lib.rjp_ray_list.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.rjp_ray_list_chunks.argtypes = [i64]
    lib.rjp_ray_list_chunks.restype = i64
    lib.rjp_integrate.argtypes = [C.POINTER(Model), C.POINTER(Epoch), C.POINTER(Continuum),
                                  vp, vp, vp, vp, i32, vp, vp, vp, vp, C.POINTER(Line),
                                  C.POINTER(Channels), i32, i32, vp, vp, i64, i64, vp, vp, vp,
                                  i64, vp, vp]
    lib.rjp_line_scratch_bytes.argtypes = [C.POINTER(Model), i64]
    lib.rjp_line_scratch_bytes.restype = i64
    lib.rjp_pack_rays.argtypes = [vp, i64, vp, i32, i32, i32, vp, vp]
    lib.rjp_column_totals.argtypes = [vp, i64, i64, vp, vp, i32, vp, vp]
    lib.rjp_column_totals.restype = C.c_int
    lib.rjp_scatter_rays.argtypes = [vp, i32, vp, i32, i32, vp, i64, vp]
    lib.rjp_fill_missed.argtypes = [vp, i64, i32, i64, i64, i64, i64, vp, vp, i32, vp]
    lib.rjp_continuum_images.argtypes = [vp, vp, vp, i64, vp, vp, dbl, i32, vp, vp, vp, vp]
    lib.rjp_integrate_epochs.argtypes = [C.POINTER(Model), C.POINTER(Epoch),
                                         C.POINTER(Continuum), vp, vp, vp, vp, i32, vp, i32, vp,
                                         vp, vp, vp, vp, vp]
    lib.rjp_integrate_epochs.restype = C.c_int
    lib.rjp_continuum_images_epochs.argtypes = [vp, i32, vp, vp, i64, vp, vp, dbl, i32, vp, vp,
                                                vp, vp]
    lib.rjp_continuum_images_epochs.restype = C.c_int
    lib.rjp_voigt_profile.argtypes = [vp, vp, i64, vp, vp]
    lib.rjp_host_assemble.argtypes = [vp, i64, i64, vp, i64, vp, i64, dbl, i32]
    lib.rjp_host_assemble.restype = C.c_int
    for f in ("rjp_struct_sizes", "rjp_fill_grid", "rjp_patch_cells", "rjp_cell_field",
              "rjp_ray_list", "rjp_integrate", "rjp_continuum_images", "rjp_voigt_profile",
              "rjp_pack_rays", "rjp_scatter_rays", "rjp_fill_missed", "rjp_los_means", "rjp_override_cells"):
        getattr(lib, f).restype = C.c_int
    sizes = [i32() for _ in range(6)]
    lib.rjp_struct_sizes(*[C.byref(s) for s in sizes])
    mirrors = (Model, Epoch, Continuum, Line, Channels)
    for s, t in zip(sizes, mirrors):
        if s.value != C.sizeof(t):
            raise EngineError(f"ABI mismatch: sizeof({t.__name__}) = {C.sizeof(t)} in "
                              f"Python, {s.value} in {path}")
    if sizes[5].value != 16:
        raise EngineError("ABI mismatch: rjp_cell is not 16 bytes")
    _lib = lib
    return lib


def check(status, what):
    if status == OK:
        return
    lib = load()
    msg = lib.rjp_strerror(status).decode()
    if status == ERR_CUDA:
        msg += " (" + lib.rjp_last_cuda_error().decode() + ")"
    raise EngineError(f"{what}: {msg}")
