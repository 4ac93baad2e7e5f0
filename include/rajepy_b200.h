/*
 * rajepy_b200 -- C ABI of the B200-native engine for RaJePy's hot path
 * (jet-grid fill + line-of-sight radiative transfer).
 *
 * The reference (SimonP2207/RaJePy) is pure Python/numpy and has NO FFI layer for this
 * path; its boundary is the Python class JetModel (classes.py:42-1713).  This header is
 * therefore the interface a maintainer binds from Python with ctypes (see
 * INTEGRATION.md); each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions.
 *   - every function returns an int status: 0 = ok, < 0 = error (rjp_strerror()).
 *   - all buffers are caller-owned DEVICE pointers unless the name ends in _host;
 *     launches are asynchronous on the cudaStream_t passed as `void* stream`
 *     (the caller synchronises).  No global mutable state.
 *   - arrays follow the reference layout: cells (nx, ny, nz) C-order, y = line of
 *     sight (classes.py:46, :465-474); images (nx, nz); cubes (nchan, nx, nz).
 *   - a device may hold an x-slab [x_lo, x_hi) of the grid (multi-GPU sharding);
 *     per-cell buffers then have (x_hi - x_lo) * ny * nz entries and image buffers
 *     (x_hi - x_lo) * nz.
 */
#ifndef RAJEPY_B200_H
#define RAJEPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RJP_ABI_VERSION 6
#define RJP_MAX_BURSTS 16

enum {
  RJP_OK = 0,
  RJP_ERR_ARG = -1,       /* null pointer / bad dimension / bad enum          */
  RJP_ERR_CUDA = -2,      /* a CUDA runtime call failed (see rjp_last_cuda_error) */
  RJP_ERR_CAPACITY = -3,  /* a caller-provided list was too small              */
  RJP_ERR_UNSUPPORTED = -4
};

/* Per-cell packed state written by rjp_fill_grid and streamed (once) by
 * rjp_integrate: 16 bytes = the algorithmic bytes per cell of SURVEY.md 8(d).
 *   ne0 : n_base * x [cm^-3] = electron density WITHOUT the burst factor chi(t)
 *         (classes.py:889-897 times :928-934); 0 = density invalid (NaN in the reference)
 *   temp: T [K] (classes.py:957-967); sign bit set = half-filled cell (ff = 0.5);
 *         |temp| == 0 = temperature invalid
 * An all-zero cell is outside the jet.  Everything else the integrators need per cell
 * (jet side, travel time -> burst factor, line-of-sight velocity) is an analytic
 * function of the cell indices and is recomputed in fp64 for the few per cent of cells
 * that are inside the jet, so the state carries full double precision in 16 bytes. */
typedef struct { double ne0; double temp; } rjp_cell;

/* Everything the grid fill needs (host scalars derived as JetModel.__init__ does,
 * classes.py:168-242; angles as maths/geometry.py:243-247). */
typedef struct {
  int32_t nx, ny, nz;        /* full grid (even, classes.py:201-205)                */
  int32_t x_lo, x_hi;        /* slab held by this device                            */
  double cs;                 /* cell size [au]                                      */
  double w0, r0, mr0, eps;   /* geometry; mr0 = mod_r_0 (geometry.py:12-31)         */
  double ca, sa, cb, sb;     /* cos/sin(radians(inc-90)), cos/sin(radians(pa))      */
  double cva, sva, cvb, svb; /* cos/sin(radians(90-inc)), cos/sin(radians(-pa))     */
  double R1, R2;             /* disc radii [au]                                     */
  double q_n, q_x, q_T, q_v; /* axial power-law indices                             */
  double qd_n, qd_x, qd_T, qd_v; /* cross-sectional indices                         */
  double n0, x0, T0, v0;     /* base values (cm^-3, -, K, km/s)                     */
  double f_rb;               /* mlr_rj / mlr_bj (classes.py:228-229, :895)          */
  double gm_over_au;         /* G * M_star * MSOL / au  (physics.py:90)             */
  double rot_sign;           /* +1 CCW, -1 CW (classes.py:1069-1071)                */
  double v_lsr;              /* km/s                                                */
  double au_m, year_s;       /* scipy.constants.au / .year                          */
  double au_cm;              /* au * 1e2: the temperature quirk (classes.py:957)    */
  /* 2F1 connection constants for the travel time when qd_v != 0 (geometry.py:166-171):
   * 2F1(a, b; b+1; z), a = qd_v, b = (1 - q_v + eps*qd_v)/eps                      */
  double hyp_b;              /* b                                                   */
  double hyp_c1;             /* b / (b - a)                                         */
  double hyp_c2;             /* Gamma(b+1) Gamma(a-b) / Gamma(a)                    */
  int32_t hyp_degenerate;    /* 1: b-a (near-)integer -> series only                */
  int32_t reserved0;
  int32_t need_reff;         /* any of qd_n, qd_x, qd_T != 0                        */
  int32_t reserved1;
} rjp_model;

/* One Gaussian ejection burst of one jet (classes.py:399-463). */
typedef struct {
  double t0;        /* peak time [s]                                   */
  double amp;       /* (peak_jml - ss_jml) / ss_jml  = chi - 1         */
  double inv2s2;    /* 1 / (2 sigma^2), sigma = hl * 2 / (2 sqrt(2 ln 2)) [s^-2] */
} rjp_burst;

typedef struct {
  double time;                       /* model time [s] (JetModel.time)           */
  int32_t n_blue, n_red;
  rjp_burst blue[RJP_MAX_BURSTS];
  rjp_burst red[RJP_MAX_BURSTS];
} rjp_epoch;

/* Continuum integration constants (classes.py:1116-1120, :1395-1399). */
typedef struct {
  double em_scale;      /* cs * au / pc                                           */
  double tau_scale;     /* 0.018 * cs * au * 100                                  */
  double t_exponent;    /* -1.5 (q_T == 0: van Hoof g_ff) or -1.35 (Reynolds g_ff folded) */
} rjp_continuum;

/* LTE recombination-line constants (classes.py:1159-1189; maths/rrls.py). */
typedef struct {
  double nu0;           /* rest frequency [Hz] (rrls.py:14-29)                    */
  double dopp;          /* 1000 / c   (physics.py:547-558)                        */
  double width_g;       /* sqrt(2 k / (m_atom c^2)): sigma*sqrt(2) = width_g*sqrt(T)*nu0_cell */
  double stark;         /* 8.2 (n/100)^4.5 (1 + 2.25 dn/n) / 2  (rrls.py:86-101)  */
  double kappa0;        /* 1.0991132675738456e-17 n^2 f (X mu'/m_amu) cs au 100 / sqrt(pi) */
  double en_over_k;     /* Z^2 E_n / k_cgs  [K]  (rrls.py:386)                    */
  double h_over_k;      /* h / k [K s]                                            */
  double dn_max;        /* max_k |nu_k - nu0| [Hz] over the channels of this call  */
  double chan_dnu0;     /* equally spaced channels: dnu[k] = chan_dnu0 + k * chan_step     */
  double chan_step;     /* [Hz]; 0 = not equally spaced (the kernels then read dnu[])      */
  /* functions of the temperature alone, precomputed for the temperature most cells have
   * (isothermal jets, q_T = q^d_T = 0); t_common = 0: none, the kernels evaluate them per cell */
  double t_common;      /* [K]                                                             */
  double tc_sqrt;       /* sqrt(t_common)                                                  */
  double tc_boltz;      /* exp(en_over_k / t_common)                                       */
  double tc_hk;         /* h_over_k / t_common                                             */
  double tc_p0;         /* 1 - exp(-tc_hk * nu0)                                           */
} rjp_line;

/* Per-channel host-prepared scalars, each a DEVICE array of nchan doubles. */
typedef struct {
  const double* dnu;     /* nu_k - nu0 [Hz]                                        */
  const double* nu;      /* nu_k [Hz]                                              */
  const double* cff;     /* tau_ff(nu_k) = cff_k * K  (nu^-2 g_ff or 11.95 nu^-2.1)*/
  const double* aff;     /* S_ff = aff_k * Tmean * (1 - exp(-tau_ff))  [Jy/pixel]  */
  const double* bnu;     /* 2 h nu^3 / c^2 [cgs] * 1e-3 * Omega_pix / 1e-26        */
} rjp_channels;

const char* rjp_strerror(int status);
const char* rjp_last_cuda_error(void);
int rjp_abi_version(void);
/* sizeof() of the ABI structs, so a ctypes binding can verify its mirror. */
int rjp_struct_sizes(int32_t* model, int32_t* epoch, int32_t* continuum, int32_t* line,
                     int32_t* channels, int32_t* cell);

/* Grid fill (K1+K2).  Replaces JetModel.fill_factor/areas (classes.py:571-784) and
 * the per-cell property chain ts/number_density/ion_fraction/temperature/vel
 * (classes.py:838-1099; maths/geometry.py:121-336).
 *   nverts [slab cells]  : number of cell vertices inside the jet (0..8), bit-exact
 *                          apart from the vertices reported in `ties`
 *   cells  [slab cells]  : packed state
 *   ties   [tie_capacity*3] int32 (I,J,K) lattice indices of vertices whose inside
 *                          test is too close to call in device arithmetic; *n_ties
 *                          (device) receives the number found (may exceed capacity:
 *                          then re-run with a larger list)
 *   extents [slab rays][2] int32: per ray (x, z) the half-open y-range [y_lo, y_hi)
 *                          that contains all of its in-jet cells (y_lo >= y_hi: the ray
 *                          misses the jet); written here, read by the ray kernels
 *   brick_state [rjp_brick_count()] uint8, optional (NULL: every cell is written):
 *                          occupancy map of the caller's nverts / cells buffers in bricks of
 *                          4 x 8 x 32 cells.  0 = this brick of both buffers is all zero,
 *                          non-zero = it holds data.  The fill skips bricks that lie outside
 *                          the jet and are already zero, zeroes bricks that lie outside but
 *                          hold data, and updates the map -- so on a zero-initialised or
 *                          recycled buffer it writes only the bricks around the jet.
 *   brick_work  [rjp_brick_count() + 4] int32 scratch, optional (needs brick_state): enables
 *                          the two-level sparse fill -- bricks are classified by one thread
 *                          each into this work list and a persistent grid pulls the bricks
 *                          that need work one at a time (even load over the SMs).           */
int rjp_fill_grid(const rjp_model* m_host, uint8_t* nverts, rjp_cell* cells,
                  uint8_t* brick_state, int32_t* brick_work, int32_t* ties,
                  int32_t tie_capacity, int32_t* n_ties, int32_t* extents, void* stream);

/* Number of bricks (entries of brick_state) of the slab described by m_host; < 0 = error. */
int64_t rjp_brick_count(const rjp_model* m_host);

/* Apply host-resolved vertex decisions: for n cells (flat slab indices `cell_idx`,
 * device) set nverts to `new_count` (device, uint8), recompute the packed state and
 * widen the ray extents where a cell enters the jet. */
int rjp_patch_cells(const rjp_model* m_host, const int64_t* cell_idx,
                    const uint8_t* new_count, int32_t n, uint8_t* nverts,
                    rjp_cell* cells, uint8_t* brick_state /* optional */, int32_t* extents,
                    void* stream);

/* Full-precision 3-D property planes on demand (float64, NaN outside the jet where the
 * reference has NaN), for the JetModel properties the plotting code reads. */
enum {
  RJP_FIELD_FILL_FACTOR = 0,  /* classes.py:667-668,763 */
  RJP_FIELD_AREAS = 1,        /* classes.py:669,764     */
  RJP_FIELD_R = 2, RJP_FIELD_W = 3, RJP_FIELD_PHI = 4,     /* classes.py:515-541 */
  RJP_FIELD_REFF = 5,         /* classes.py:543-557     */
  RJP_FIELD_TRAVEL = 6,       /* seconds; JetModel.ts = time - this (classes.py:838-859) */
  RJP_FIELD_ND_BASE = 7,      /* classes.py:872-897 (without chi) */
  RJP_FIELD_XI = 8,           /* classes.py:910-936     */
  RJP_FIELD_TEMP = 9,         /* classes.py:942-969     */
  RJP_FIELD_VX = 10, RJP_FIELD_VLOS = 11, RJP_FIELD_VZ = 12, /* classes.py:1009-1095 */
  RJP_FIELD_CHI = 13,         /* classes.py:861-870 (needs epoch) */
  RJP_FIELD_COUNT = 14
};
int rjp_cell_field(const rjp_model* m_host, const rjp_epoch* ep_host,
                   const uint8_t* nverts, int32_t field, double* out, void* stream);

/* User-assigned grids (the `temperature` / `ion_fraction` setters, classes.py:936-940,
 * :994-1000): replace T (field = RJP_FIELD_TEMP) or the ionisation fraction (RJP_FIELD_XI) of
 * every in-jet cell of the packed state by values[slab cell] (NaN / non-positive = invalid). */
int rjp_override_cells(const rjp_model* m_host, const uint8_t* nverts, int32_t field,
                       const double* values, rjp_cell* cells, void* stream);

/* Ordered list of the rays whose extent is non-empty (slab-local ray index x_local * nz + z,
 * ASCENDING): the ray kernels pull their work from it.  `list` holds up to nray entries,
 * `chunk_counts` is scratch of rjp_ray_list_chunks(nray) int32, *n_active (DEVICE) receives the
 * count and STAYS on the device -- rjp_integrate / rjp_los_means read it there, so the host
 * never has to wait for it.  Call after rjp_fill_grid / rjp_patch_cells. */
int rjp_ray_list(const int32_t* extents, int64_t nray, int32_t* list, int32_t* chunk_counts,
                 int32_t* n_active, void* stream);
int64_t rjp_ray_list_chunks(int64_t nray);

/* Line-of-sight pass (K3+K4+K5).  The ray kernels walk only the per-ray in-jet extents recorded
 * by the fill (the CTAs stride over `ray_list`; *n_active is read on the device) and run on
 * `stream2` when one is given (NULL: same stream) beside the constant writer, which streams
 * the 0 / NaN of the rays that miss the jet with TMA bulk stores on `stream`.
 * Replaces emission_measure (classes.py:1101-1128), optical_depth_ff (:1353-1447),
 * the nanmean temperature of intensity_ff (:1471-1473), optical_depth_rrl (:1130-1229)
 * and intensity_rrl/flux_rrl (:1231-1351).
 *   em, kff, tsum [nxs*nz] double, tcount [nxs*nz] int32 (always written)
 *   extents / ray_list / n_active (from rjp_fill_grid / rjp_ray_list); with any of them NULL
 *   (continuum-only passes) every cell of the state is swept instead (dense sweep).
 *   n_active_hint: the caller's guess of *n_active on the HOST, or < 0 if it has none
 *   (0 is a count: no ray of the slab crosses the jet, the pass is the constant writer alone).
 *   It only sizes the grids (one CTA per listed ray is fastest for the line kernel; the
 *   constant writer gets the SMs the channel loop will not need); any value gives correct
 *   results, so a count remembered from an earlier model of the same geometry is fine and
 *   nothing has to be read back from the device.
 *   line/ch may be NULL/nchan = 0 for a continuum-only pass; otherwise
 *   tau_rrl and/or flux_rrl ([nchan][nxs][nz] double) may each be NULL.
 *   contsub: 0 -> flux_rrl includes S_ff (what Pipeline requests, classes.py:2450).
 *   cube_plane / cube_offset: 0 / 0 -> the cubes are slab tiles [nchan][nxs*nz]; otherwise
 *   ray i of the slab is element cube_offset + i of every plane of cube_plane elements, so
 *   a slab can write straight into its rows of a full-size [nchan][nx*nz] cube
 *   (cube_plane = nx*nz, cube_offset = x_lo*nz).
 *   travel_cells / vlos_cells: optional (NULL) user-assigned per-cell grids [slab cells] double
 *   -- the `ts` and `vel` setters of the reference (classes.py:857-859, :1097-1099): travel time
 *   from the jet base [s] (NaN: the cell's density is dropped like the reference's nansum does)
 *   and line-of-sight velocity incl. v_lsr [km/s]; NULL = recomputed from the cell indices.
 *   line_scratch / line_max_cells (line passes only): DEVICE scratch of
 *   rjp_line_scratch_bytes(m, line_max_cells) bytes, where line_max_cells >= the summed
 *   lengths of the extents of the listed rays (the prepared cells of every jet-crossing ray
 *   are staged there between the two kernels of a line pass: ray_prepare_kernel evaluates the
 *   per-cell line constants once, integrate_line_kernel is the channel loop alone).  A ray
 *   whose cells do not fit gets NaN in every channel of both cubes.
 * On return all work is ordered on `stream` (stream2 is joined back).              */
int64_t rjp_line_scratch_bytes(const rjp_model* m_host, int64_t line_max_cells);
int rjp_integrate(const rjp_model* m_host, const rjp_epoch* ep_host,
                  const rjp_continuum* cont_host, const rjp_cell* cells,
                  const int32_t* extents, const int32_t* ray_list, const int32_t* n_active,
                  int32_t n_active_hint, double* em, double* kff, double* tsum,
                  int32_t* tcount, const rjp_line* line_host, const rjp_channels* ch_host,
                  int32_t nchan, int32_t contsub, double* tau_rrl, double* flux_rrl,
                  int64_t cube_plane, int64_t cube_offset, const double* travel_cells,
                  const double* vlos_cells, void* line_scratch, int64_t line_max_cells,
                  void* stream, void* stream2);

/* Sparse exchange of cube tiles between x-slabs (multi-GPU): 94 % of the rays
 * of the BASELINE jet miss the jet and carry constants (tau_L = 0, flux = NaN) that every rank
 * can write itself once it knows the extents, so only the columns of jet-crossing rays travel.
 *   rjp_pack_rays    out[c * n_stride + k] = cube[c * cube_plane + ray_ids[k]],  k < n
 *   rjp_scatter_rays cube[c * cube_plane + ray_ids[k]] = in[c * n_stride + k]
 *   rjp_fill_missed  for the nray rays described by `extents` (any slab's, e.g. all-gathered):
 *                    tau = 0 / flux = NaN in every channel plane where the extent is empty;
 *                    ray i is element cube_offset + i of a plane; rays in [skip_lo, skip_hi)
 *                    are left untouched (the caller's own slab when `extents` covers the
 *                    whole image: ONE launch then writes every other slab's constants).
 *                    tau or flux may be NULL.
 *                    Whole tiles of 1024 consecutive constant rays go out as TMA bulk stores
 *                    (needs even cube_plane / cube_offset / nray and 16-byte aligned cubes,
 *                    else predicated scalar stores).
 *                    light != 0: a small grid meant to run on a side stream beside a long
 *                    channel loop; 0: a grid that reaches the HBM write bandwidth alone.
 * ray_ids index into a plane (global ray = x * nz + z for a full-size cube).          */
/* totals[c] = sum over the listed rays of cube[c * cube_plane + cube_offset + ray], NaN skipped:
 * the sky-summed flux of every channel (Pipeline's results['flux'], classes.py:2468-2472) --
 * every other ray of a line cube holds the constant 0 / NaN.  Deterministic summation order. */
int rjp_column_totals(const double* cube, int64_t cube_plane, int64_t cube_offset,
                      const int32_t* ray_list, const int32_t* n_active, int32_t nchan,
                      double* totals, void* stream);
int rjp_pack_rays(const double* cube, int64_t cube_plane, const int32_t* ray_ids, int32_t n,
                  int32_t n_stride, int32_t nchan, double* out, void* stream);
int rjp_scatter_rays(const double* in, int32_t n_stride, const int32_t* ray_ids, int32_t n,
                     int32_t nchan, double* cube, int64_t cube_plane, void* stream);
int rjp_fill_missed(const int32_t* extents, int64_t nray, int32_t nchan, int64_t cube_plane,
                    int64_t cube_offset, int64_t skip_lo, int64_t skip_hi, double* tau,
                    double* flux, int32_t light, void* stream);

/* HOST-side assembly of a dense product cube (the hand-over to numpy, i.e. what
 * optical_depth_rrl / flux_rrl return, classes.py:1215-1229, :1340-1351) from the packed columns
 * of the jet-crossing rays -- so that only those columns have to cross PCIe:
 *   dst_host[c * plane + r] = fill                          for every ray r not listed
 *   dst_host[c * plane + ray_ids_host[k]] = cols_host[c * n_stride + k]
 * for c < nchan.  ray_ids_host must be strictly ascending (what rjp_ray_list produces);
 * cols_host is what rjp_pack_rays wrote, copied to the host.  The constants are written with
 * streaming stores by `nthreads` host threads (<= 0: all hardware threads).  All pointers are
 * HOST pointers; no CUDA call is made. */
int rjp_host_assemble(double* dst_host, int64_t nchan, int64_t plane,
                      const int32_t* ray_ids_host, int64_t n, const double* cols_host,
                      int64_t n_stride, double fill, int32_t nthreads);

/* Continuum epilogue (K5) for nfreq frequencies from one pass' kff/tsum/tcount:
 *   tau[f] = cff[f] * kff;  I[f] = iff[f] * Tmean * (1 - exp(-tau));  S[f] = I * omega_jy
 * (classes.py:1427-1432, :1484-1488, :1531-1533).  cff/iff are DEVICE arrays [nfreq];
 * any of tau/intensity/flux ([nfreq][npix]) may be NULL.                             */
int rjp_continuum_images(const double* kff, const double* tsum, const int32_t* tcount,
                         int64_t npix, const double* cff, const double* iff,
                         double omega_jy, int32_t nfreq, double* tau, double* intensity,
                         double* flux, void* stream);

/* Continuum sums for a BATCH of model times in one launch -- the variable-ejection time series
 * (BASELINE configs[3]): Pipeline repeats the continuum run once per run year with
 * JetModel.time advanced (classes.py:2347-2453), and the model time enters the integrals only
 * through the burst factor chi(time - travel time) of each cell (classes.py:861-870,
 * :399-463).  Every jet-crossing ray is walked once per block of 8 epochs; the cell load, the
 * travel time and the temperature power are shared by the epochs of a block.
 *   times  [n_epochs] DEVICE doubles, model times [s] (ep_host->time is ignored; the bursts
 *          are ep_host's)
 *   em (optional, may be NULL), kff: [n_epochs][nxs*nz];  tsum, tcount: [nxs*nz] (they do not
 *          depend on the epoch).  ONLY the pixels of the listed rays are written: the caller
 *          zero-fills the buffers (rays that miss the jet hold 0, classes.py:1120, :1427).
 *   travel_cells: optional user-assigned travel-time grid as in rjp_integrate.
 * Per epoch the result equals rjp_integrate's continuum sums for that model time.          */
int rjp_integrate_epochs(const rjp_model* m_host, const rjp_epoch* ep_host,
                         const rjp_continuum* cont_host, const rjp_cell* cells,
                         const int32_t* extents, const int32_t* ray_list,
                         const int32_t* n_active, int32_t n_active_hint, const double* times,
                         int32_t n_epochs, double* em, double* kff, double* tsum,
                         int32_t* tcount, const double* travel_cells, void* stream);

/* rjp_continuum_images for the K planes of a batch of epochs: kff [n_epochs][npix], tsum /
 * tcount [npix]; outputs [n_epochs][nfreq][npix] (any may be NULL).                        */
int rjp_continuum_images_epochs(const double* kff, int32_t n_epochs, const double* tsum,
                                const int32_t* tcount, int64_t npix, const double* cff,
                                const double* iff, double omega_jy, int32_t nfreq, double* tau,
                                double* intensity, double* flux, void* stream);

/* Line-of-sight means of the cell properties for the model plot (replaces the four
 * np.nanmean(<3-D grid>, axis=los) of plotting/functions.py:539-590 and the nanmin / nanmax
 * that scale its colour bars), from the ray walk: out is [7][nxs*nz] double =
 *   0 mean number density (with the burst factor, classes.py:872-899)   1 mean temperature
 *   2 mean ionisation fraction   3 mean (v_los - v_lsr) [km/s]
 *   4 min n, 5 max n, 6 max T along the ray;   NaN where the ray has no finite value.     */
int rjp_los_means(const rjp_model* m_host, const rjp_epoch* ep_host, const uint8_t* nverts,
                  const int32_t* extents, const int32_t* ray_list, const int32_t* n_active,
                  double* out, void* stream);

/* Voigt profile function of the channel loop, element-wise: out[i] = Re w(x[i] + i y[i]),
 * w = Faddeeva function, y > 0, evaluated by the same device routines the line-of-sight pass
 * uses (mixed fp64/fp32 split for RJP_VT_Y_MIN <= y <= 0.1, fp64 rational approximation
 * otherwise).  Replaces scipy.special.wofz at maths/rrls.py:353; exported so that the
 * parity tests can check the approximation itself against wofz on the device.
 * x, y, out: DEVICE arrays of n doubles. */
int rjp_voigt_profile(const double* x, const double* y, int64_t n, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAJEPY_B200_H */
