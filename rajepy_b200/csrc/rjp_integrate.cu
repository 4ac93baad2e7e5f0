// Line-of-sight integration for sm_100a: K3 (continuum sums), K4 (LTE recombination-
// line opacity over all velocity channels) and K5 (image / cube epilogue).
//
// The grid fill records, per ray, the y-extent [y_lo, y_hi) of its in-jet cells and the list
// of rays that cross the jet at all (0.4 % of the cells, 6 % of the rays of the BASELINE
// 1024^3 grid).  The pass walks exactly those extents and never touches the empty part of
// the 16 B/cell state:
//
// K4 `integrate_line_kernel` -- the channel loop, one CTA per jet-crossing ray: thread g
// owns channels g, g + NT, ... and keeps their tau_L in REGISTERS while the CTA walks the
// ray's extent: the threads each prepare one cell (burst factor, Doppler shift, widths,
// amplitude -> shared-memory entry), then every thread adds all prepared cells to its
// channels.  The Voigt function is the mixed fp64/fp32 split of rjp_device.cuh for cells
// with y <= 0.1 and the fp64 rational approximation otherwise.  Issue-bound.  The same
// walk yields the ray's continuum sums (EM, K, sum T, count) and the flux epilogue.
//
// K3 `continuum_rays_kernel` -- continuum-only passes: one warp per jet-crossing ray.
// `const_tiles_kernel` streams the constants (0 / NaN) of the rays that miss the jet into
// the images and cubes with TMA bulk stores -- the only HBM-bound part of the pass
// (write-only), run beside the channel loop on the caller's first stream.
//
// `integrate_continuum_kernel` -- the dense sweep for callers that have a cell state but no
// extents: every 16-byte cell read once, CTA = 32 adjacent rays, lanes along z so every
// warp-wide load is one contiguous 512-byte row, 8 rows in flight per warp; measured at the
// HBM copy bandwidth (profiles/README.md).
#include <stdlib.h>
#include <mutex>
#include "rjp_device.cuh"

namespace rjp {

#ifdef RJP_EXPERIMENT
__device__ unsigned long long g_stamp[4] = {~0ull, 0ull, ~0ull, 0ull};
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define RJP_STAMP_BEGIN(k) if (threadIdx.x == 0) atomicMin(&g_stamp[2 * (k)], gtime());
#define RJP_STAMP_END(k) if (threadIdx.x == 0) atomicMax(&g_stamp[2 * (k) + 1], gtime());
#else
#define RJP_STAMP_BEGIN(k)
#define RJP_STAMP_END(k)
#endif

constexpr int ZT = 32;          // rays per CTA of the dense sweep
constexpr int RPW = 8;          // rows in flight per warp
// tuning knobs of the channel loop (measured on B200, profiles/README.md: 8 channels per
// thread; 12 CTAs of 64 threads per SM at 80 registers -- 8 / 9 / 10 / 14 / 16 CTAs per SM:
// 4.79 / 4.62 / 4.60 / 4.43 / 4.51 ms against 4.39 -- since the per-cell preparation moved
// into its own kernel)
#ifndef RJP_GCH
#define RJP_GCH 8
#endif
#ifndef RJP_MINB64
#define RJP_MINB64 8
#endif
#ifndef RJP_MINB64U
#define RJP_MINB64U 12
#endif
#ifndef RJP_MINB128
#define RJP_MINB128 4
#endif
constexpr int GCH_MAX = RJP_GCH;  // channels per thread of the widest line-kernel variant
constexpr int LINE_THREADS = 256;

struct LineEntry {   // channel-independent factors of one in-jet cell (rrls.py:329-389)
  double xs;         // -(nu0_cell - nu0) / (sigma sqrt2)
  double inv_s2;     // 1 / (sigma sqrt2)
  double y;          // (dnu_L / 2) / (sigma sqrt2)
  double amp;        // kappa0 n_e^2 T^-1.5 exp(Z^2 E_n / kT) ff / (sigma sqrt2); 0 = skip
  double p0;         // 1 - exp(-h nu0 / kT)
  double hk;         // h / (k T); negative: |hk * dn| is small for every channel, use a1..a3
  double a1, a2;     // 1 - exp(-h nu/kT) = p0 + dn (a1 + dn (a2 + dn a3)) + O((hk dn)^4)
  double a3;
  double pad;        // 1: the quadratic term of that polynomial is negligible (fast class)
};

__device__ __forceinline__ double2 ld_cell(const double2* p) {
  double2 r;  // streamed once: do not keep in L1
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ bool empty_cell(const double2& c) {
  return c.x == 0.0 && c.y == 0.0;  // (-0.0 == 0.0): nothing to add either way
}

// Kernel parameters live in the constant bank; the sparse slow paths below are real
// (non-inlined) functions that take pointers, so each CTA stages the parameter blocks
// in shared memory once (1.2 KB) instead of every thread copying them to its
// local-memory stack.  Two derived travel-time constants are added on the way.
// User-assigned per-cell grids (the `ts` / `vel` setters of the reference, classes.py:857-859,
// :1097-1099): when given, the travel time [s] / line-of-sight velocity [km/s] of a cell is
// read from the slab-shaped array instead of being recomputed from the cell indices.
struct CellGrids {
  const double* travel;
  const double* vlos;
};

struct Params {
  rjp_model m;
  rjp_epoch ep;
  CellGrids ov;
  double tt_cst;   // MR0^q_v / (V0 (1 - q_v + eps q^d_v))          (geometry.py:154)
  double tt_f0;    // tt_cst * MR0^(1 - q_v): indefinite integral at r_0 when q^d_v = 0
};

__device__ __forceinline__ void stage_params(Params* s_p, const rjp_model& m,
                                             const rjp_epoch& ep,
                                             const CellGrids ov = CellGrids{nullptr, nullptr}) {
  const int nm = sizeof(rjp_model) / 4, ne = sizeof(rjp_epoch) / 4;
  const uint32_t* gm = reinterpret_cast<const uint32_t*>(&m);
  const uint32_t* ge = reinterpret_cast<const uint32_t*>(&ep);
  uint32_t* dm = reinterpret_cast<uint32_t*>(&s_p->m);
  uint32_t* de = reinterpret_cast<uint32_t*>(&s_p->ep);
  for (int i = threadIdx.x; i < nm; i += blockDim.x) dm[i] = gm[i];
  for (int i = threadIdx.x; i < ne; i += blockDim.x) de[i] = ge[i];
  if (threadIdx.x == 0) {
    s_p->ov = ov;
    const double MR0 = m.mr0 * m.au_m, V0 = m.v0 * 1e3;
    const double cst = powq(MR0, m.q_v) / (V0 * (1.0 - m.q_v + m.eps * m.qd_v));
    s_p->tt_cst = cst;
    s_p->tt_f0 = cst * powq(MR0, 1.0 - m.q_v);
  }
  __syncthreads();
}

// Ray constants of thread (ix, iz): the part of maths/geometry.py:249-255 that does not
// depend on y.  Same roundings as centroid_rw().
struct Ray { double x1, z1; };

__device__ __forceinline__ Ray ray_of(const rjp_model& m, int ix, int iz) {
  const double h = m.cs / 2.0;
  const double x = __dadd_rn(corner(m.cs, ix, m.nx), h);
  const double z = __dadd_rn(corner(m.cs, iz, m.nz), h);
  Ray r;
  r.x1 = __dadd_rn(__dmul_rn(m.cb, x), __dmul_rn(m.sb, z));
  r.z1 = __dsub_rn(__dmul_rn(m.cb, z), __dmul_rn(m.sb, x));
  return r;
}

// Not inlined: rare paths whose pow / 2F1 code would otherwise bloat the streaming loops.
__device__ __noinline__ double pow_call(double x, double q) { return pow(x, q); }

__device__ __noinline__ double travel_slow(const Params* P, int ix, int iy, int iz) {
  return travel_time(P->m, centroid_rw(P->m, ix, iy, iz));
}

struct Decoded {
  double ne;      // n_e = n_base * x * chi(t) [cm^-3], 0 if invalid
  double temp;    // K, 0 if invalid
  double ffw;     // 0.5 or 1
  bool ne_ok, t_ok;
};

// The burst factor needs the launch time of the cell's material: model time minus the
// travel time from the jet base (classes.py:845, :866-868), an analytic function of the
// cell indices that is recomputed here in fp64 (in-jet cells only).  With no
// cross-sectional velocity law (q^d_v = 0, every BASELINE configuration) the travel time
// is the closed form cst * (rad^(1-q_v) - MR0^(1-q_v)) of the axial coordinate alone.
__device__ __forceinline__ Decoded decode(const double2& c, const Params& P, const Ray& ray,
                                          int ix, int iy, int iz) {
  const rjp_model& m = P.m;
  Decoded d;
  d.ffw = signbit(c.y) ? 0.5 : 1.0;
  d.temp = fabs(c.y);
  d.t_ok = d.temp > 0.0;
  d.ne_ok = c.x > 0.0;
  d.ne = 0.0;
  if (d.ne_ok) {
    const double y = __dadd_rn(corner(m.cs, iy, m.ny), m.cs / 2.0);
    const double r = __dadd_rn(__dmul_rn(m.sa, y), __dmul_rn(m.ca, ray.z1));
    double travel;
    if (P.ov.travel != nullptr) {
      travel = P.ov.travel[((size_t)(ix - m.x_lo) * m.ny + iy) * m.nz + iz];
    } else if (m.qd_v == 0.0) {
      const double rad = (r_shifted(m, fabs(r)) + m.mr0 - m.r0) * m.au_m;
      const double e = 1.0 - m.q_v;
      travel = P.tt_cst * ((e == 1.0) ? rad : pow_call(rad, e)) - P.tt_f0;
    } else {
      travel = travel_slow(&P, ix, iy, iz);
    }
    const double tl = P.ep.time - travel;
    const double chi = (r < 0.0) ? burst_chi(P.ep.red, P.ep.n_red, tl)
                                 : burst_chi(P.ep.blue, P.ep.n_blue, tl);
    d.ne = c.x * chi;  // classes.py:875, :1375
    if (!(d.ne == d.ne)) { d.ne = 0.0; d.ne_ok = false; }  // NaN travel time -> NaN density
  }
  return d;
}

struct ContAcc { double em, kff, tsum; int cnt; };

// classes.py:1116-1120 (EM), :1395-1399 (tau_ff without nu^-2 g_ff), :1471-1472 (T sum)
__device__ __forceinline__ void accumulate(ContAcc& a, const Decoded& d, double t_exp) {
  const double ne2 = d.ne * d.ne * d.ffw;
  if (d.ne_ok) a.em += ne2;
  if (d.t_ok) {
    a.tsum += d.temp;
    a.cnt += 1;
    if (d.ne_ok) {
      const double tp = (t_exp == -1.5) ? 1.0 / (d.temp * sqrt(d.temp))
                                        : pow_call(d.temp, t_exp);
      a.kff += tp * ne2;
    }
  }
}

// Cross-warp reduction of the per-(warp, ray) partial sums; result valid in warp 0.
__device__ __forceinline__ void reduce_and_store(ContAcc a, const rjp_continuum& ct,
                                                 double* s_red, int* s_cnt, int nwarps,
                                                 double* em, double* kff, double* tsum,
                                                 int32_t* tcount, size_t pix, bool active,
                                                 double* s_out /* [4][ZT] */) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  s_red[(0 * nwarps + wrp) * ZT + lane] = a.em;
  s_red[(1 * nwarps + wrp) * ZT + lane] = a.kff;
  s_red[(2 * nwarps + wrp) * ZT + lane] = a.tsum;
  s_cnt[wrp * ZT + lane] = a.cnt;
  __syncthreads();
  if (wrp == 0) {
    double e = 0, k = 0, t = 0;
    int c = 0;
    for (int w = 0; w < nwarps; ++w) {
      e += s_red[(0 * nwarps + w) * ZT + lane];
      k += s_red[(1 * nwarps + w) * ZT + lane];
      t += s_red[(2 * nwarps + w) * ZT + lane];
      c += s_cnt[w * ZT + lane];
    }
    e *= ct.em_scale;
    k *= ct.tau_scale;
    if (active) {
      em[pix] = e;
      kff[pix] = k;
      tsum[pix] = t;
      tcount[pix] = c;
    }
    if (s_out) {
      s_out[0 * ZT + lane] = k;
      s_out[1 * ZT + lane] = t;
      s_out[2 * ZT + lane] = (double)c;
    }
  }
}

// ------------------------------------------------------------------ continuum only
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
integrate_continuum_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                           const CellGrids ov,
                           const double2* __restrict__ cells, double* __restrict__ em,
                           double* __restrict__ kff, double* __restrict__ tsum,
                           int32_t* __restrict__ tcount) {
  __shared__ double s_red[3 * 8 * ZT];
  __shared__ int s_cnt[8 * ZT];
  __shared__ Params s_p;
  stage_params(&s_p, m, ep, ov);
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ztiles = (m.nz + ZT - 1) / ZT;
  const int xl = blockIdx.x / ztiles;
  const int iz = (blockIdx.x % ztiles) * ZT + lane;
  const bool active = iz < m.nz;
  const double2* base = cells + (size_t)xl * m.ny * m.nz + (active ? iz : 0);
  const int ix = m.x_lo + xl;
  const Ray ray = ray_of(s_p.m, ix, iz);
  ContAcc a = {0.0, 0.0, 0.0, 0};
  for (int y0 = wrp; y0 < m.ny; y0 += nwarps * RPW) {
    double2 c[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      const int y = y0 + j * nwarps;
      c[j] = (active && y < m.ny) ? ld_cell(base + (size_t)y * m.nz) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      if (empty_cell(c[j])) continue;
      accumulate(a, decode(c[j], s_p, ray, ix, y0 + j * nwarps, iz), ct.t_exponent);
    }
  }
  reduce_and_store(a, ct, s_red, s_cnt, nwarps, em, kff, tsum, tcount,
                   (size_t)xl * m.nz + iz, active, nullptr);
}

// ------------------------------------------------------------------ line channels (K4)
__device__ __noinline__ LineEntry make_entry(const Decoded& d, const rjp_model& m,
                                             const rjp_line& ln, double dn_max, int ix,
                                             int iy, int iz, const double* vlos_grid) {
  LineEntry e;
  const double vlos = vlos_grid
      ? vlos_grid[((size_t)(ix - m.x_lo) * m.ny + iy) * m.nz + iz]
      : velocity_of(m, centroid_rw(m, ix, iy, iz)).vlos_rel + m.v_lsr;
  const double shift = -ln.nu0 * (vlos * ln.dopp);           // nu0_cell - nu0 (physics.py:558)
  const double nu0c = ln.nu0 + shift;
  // functions of the temperature alone: precomputed on the host for the temperature most
  // cells have (a jet with q_T = q^d_T = 0 is isothermal: every BASELINE configuration)
  double sqt, boltz;
  if (d.temp == ln.t_common) {
    sqt = ln.tc_sqrt;
    boltz = ln.tc_boltz;
    e.hk = ln.tc_hk;
    e.p0 = ln.tc_p0;
  } else {
    sqt = sqrt(d.temp);
    boltz = exp(ln.en_over_k / d.temp);
    e.hk = ln.h_over_k / d.temp;
    e.p0 = -expm1(-e.hk * ln.nu0);
  }
  const double s2 = ln.width_g * sqt * nu0c;                 // sigma*sqrt2 (rrls.py:104-118, :349)
  e.inv_s2 = 1.0 / s2;
  e.xs = -shift * e.inv_s2;
  e.y = ln.stark * d.ne * e.inv_s2;                          // rrls.py:101, :353
  // Taylor coefficients of (1 - p0)(1 - exp(-hk dn)) in dn; used when |hk dn| <= 1e-3
  // for all channels (error <= (1e-3)^4 / 24 relative), see planck_factor()
  e.a1 = (1.0 - e.p0) * e.hk;
  e.a2 = -0.5 * e.a1 * e.hk;
  e.a3 = (1.0 / 6.0) * e.a1 * e.hk * e.hk;
  // the fast class keeps only the linear term: its quadratic one must be < 1e-9 of p0
  e.pad = (fabs(e.a2) * dn_max * dn_max <= 1e-9 * e.p0) ? 1.0 : 0.0;
  if (e.hk * dn_max <= 1e-3) e.hk = -e.hk;
  // rrls.py:383-389 with n_i = (X mu'/m_amu) n_e, times path length and 1/(sigma sqrt(2 pi))
  e.amp = ln.kappa0 * d.ne * d.ne * d.ffw / (d.temp * sqt) * boltz * e.inv_s2;
  // NaN velocity / density / temperature: the reference's nansum drops the cell
  if (!(vlos == vlos) || !d.ne_ok || !d.t_ok || !(e.amp == e.amp)) e.amp = 0.0;
  return e;
}

// Fast class (rjp_device.cuh, "mixed-precision Voigt"): small Lorentz/Gauss ratio and a
// Planck factor that is a short polynomial in the channel offset.
__device__ __forceinline__ bool fast_class(const LineEntry& e) {
  return e.amp != 0.0 && e.hk < 0.0 && e.pad != 0.0 && e.y >= RJP_VT_Y_MIN &&
         e.y <= RJP_VT_Y_MAX;
}

__device__ __noinline__ FastEntry to_fast(const LineEntry& s) {
  FastEntry e;
  vt_cell_constants(s.y, e);
  e.xs = RJP_VT_KAPPA * s.xs;
  e.inv = RJP_VT_KAPPA * s.inv_s2;
  e.a0 = s.amp * s.p0;
  e.w0 = e.a0 * s.y * (RJP_VT_KAPPA * RJP_VT_KAPPA * RJP_VT_G10);
  e.b1 = (float)(s.a1 / s.p0);
  e.b2 = 0.0f;
  return e;
}

// 1 - exp(-h nu / kT) for nu = nu0 + dn from the cell's value at nu0 (rrls.py:387)
__device__ __forceinline__ double planck_factor(const LineEntry& e, double dn) {
  if (e.hk < 0.0) return fma(dn, fma(dn, fma(dn, e.a3, e.a2), e.a1), e.p0);
  return e.p0 + (1.0 - e.p0) * (-expm1(-e.hk * dn));
}

// Rays that miss the jet: EM = K = sum T = 0, count = 0, tau_L = 0 and flux = NaN in every
// channel (nansum / nanmean of an all-NaN column) -- 8 GB of constants at
// 1024^2 rays x 512 channels x 2 cubes, the only HBM-bound part of the pass.  The ray index
// space is cut into tiles of CT_TILE consecutive rays (one sky row at nz = 1024).  ~90 % of
// the tiles contain no jet-crossing ray at all: for those ONE elected thread streams the
// constant out of shared memory with TMA bulk stores (cp.async.bulk.global.shared::cta, SASS
// UBLKCP.G.S; 8 KB per instruction, the source tile never changes so nothing waits on the
// reads) -- practically no issue slots, which is what lets this kernel run beside the
// issue-bound channel loop.  Tiles that contain jet-crossing rays (or rays of the skip range)
// take predicated scalar stores; the ray kernels own the jet-crossing columns.
//   Ray i of `extents` is element offset + i of every cube plane (plane = elements per
// plane): a slab writes into its rows of a full-size cube this way.  Rays in
// [skip_lo, skip_hi) are left alone (multi-GPU: the own slab inside the global ray range).
constexpr int CT_TILE = 1024;    // rays per tile
constexpr int CT_SRC = 1024;     // doubles per shared-memory source tile (8 KB per bulk store)
#ifndef RJP_CT_CG
#define RJP_CT_CG 32             // channel planes per work item
#endif

__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
#ifndef RJP_NO_EVICT_FIRST
  // the constants are never read again on the device: first in line for eviction, so that
  // they do not push the ray kernels' partially written sectors out of L2
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               :: "l"(gdst), "r"(ssrc), "r"(bytes), "l"(pol) : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
#endif
}

// Column stores of the ray kernels: 8 bytes per (ray, channel), neighbouring rays (other
// CTAs) complete the 32-byte sector a little later
__device__ __forceinline__ void st_column(double* p, double v) { *p = v; }

struct ConstJob {      // what the constant writer needs (kernel-argument block)
  const int2* extents;
  size_t nray;
  int nchan;
  double *em, *kff, *tsum;
  int32_t* tcount;
  double *tau, *flux;
  size_t plane, offset, skip_lo, skip_hi;
  int use_bulk;
};

__device__ __forceinline__ size_t const_items(const ConstJob& j) {
  const size_t ntiles = (j.nray + CT_TILE - 1) / CT_TILE;
  return ntiles * (size_t)(j.nchan > 0 ? (j.nchan + RJP_CT_CG - 1) / RJP_CT_CG : 1);
}

// Fill the two source tiles (all threads of the CTA) and publish them to the async proxy.
__device__ __forceinline__ void const_sources(double* s_zero, double* s_nan) {
  for (int i = threadIdx.x; i < CT_SRC; i += blockDim.x) {
    s_zero[i] = 0.0;
    s_nan[i] = dnan();
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One work item = (tile of CT_TILE rays) x (group of RJP_CT_CG channel planes), done by the
// ONE warp of the calling CTA.
__device__ __forceinline__ void const_item(const ConstJob& j, size_t item, const double* s_zero,
                                           const double* s_nan) {
  const int NT = 32, g = threadIdx.x & 31;
  const int ncg = j.nchan > 0 ? (j.nchan + RJP_CT_CG - 1) / RJP_CT_CG : 1;
  const size_t t = item / ncg;
  const int cg = (int)(item - t * ncg);
  const size_t r0 = t * CT_TILE;
  const int per = CT_TILE / NT;                       // rays per thread (<= 32)
  // bit i of `miss`: ray r0 + NT i + g misses the jet (and is not in the skip range);
  // bit i of `valid`: that ray exists
  unsigned miss = 0u, valid = 0u;
  for (int i = 0; i < per; ++i) {
    const size_t ray = r0 + (size_t)NT * i + g;
    if (ray < j.nray) {
      valid |= 1u << i;
      if (!(ray >= j.skip_lo && ray < j.skip_hi)) {
        const int2 e = __ldg(j.extents + ray);
        if (e.x >= e.y) miss |= 1u << i;
      }
    }
  }
  if (cg == 0 && j.em != nullptr) {     // the four sky images, once per tile
    for (int i = 0; i < per; ++i) {
      if (miss >> i & 1u) {
        const size_t ray = r0 + (size_t)NT * i + g;
        j.em[ray] = 0.0;
        j.kff[ray] = 0.0;
        j.tsum[ray] = 0.0;
        j.tcount[ray] = 0;
      }
    }
  }
  if (j.nchan <= 0) return;
  const int c_lo = cg * RJP_CT_CG;
  const int c_hi = (c_lo + RJP_CT_CG < j.nchan) ? c_lo + RJP_CT_CG : j.nchan;
  const size_t n_in = (j.nray - r0 < (size_t)CT_TILE) ? j.nray - r0 : (size_t)CT_TILE;
  // whole tile constant?  (a short last tile counts if all of its rays miss and its byte
  // count is a multiple of 16)
  const int same = __all_sync(0xffffffffu, miss == valid);
  const int any = __any_sync(0xffffffffu, miss != 0u);
  if (same && (n_in & 1) == 0 && j.use_bulk) {
    if (g == 0) {
      const uint32_t a_zero = (uint32_t)__cvta_generic_to_shared(s_zero);
      const uint32_t a_nan = (uint32_t)__cvta_generic_to_shared(s_nan);
      for (int c = c_lo; c < c_hi; ++c) {
        const size_t o = (size_t)c * j.plane + j.offset + r0;
        for (size_t k = 0; k < n_in; k += CT_SRC) {
          const uint32_t bytes =
              (uint32_t)((n_in - k < (size_t)CT_SRC ? n_in - k : (size_t)CT_SRC) * sizeof(double));
          if (j.tau) bulk_store(j.tau + o + k, a_zero, bytes);
          if (j.flux) bulk_store(j.flux + o + k, a_nan, bytes);
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else if (any) {
    const double nanv = dnan();
    for (int c = c_lo; c < c_hi; ++c) {
      const size_t o = (size_t)c * j.plane + j.offset + r0 + g;
      for (int i = 0; i < per; ++i) {
        if (miss >> i & 1u) {
          if (j.tau) j.tau[o + (size_t)NT * i] = 0.0;
          if (j.flux) j.flux[o + (size_t)NT * i] = nanv;
        }
      }
    }
  }
}

// the shared-memory sources must outlive the reads, the writes must have landed at exit
__device__ __forceinline__ void const_drain() {
  if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
}

// The constant writer: every warp of the grid takes work items round-robin (consecutive items
// go to different CTAs); the warps of a CTA share the two source tiles.
#ifdef RJP_WRITER_MAXNREG
__global__ void __maxnreg__(RJP_WRITER_MAXNREG)
#else
__global__ void __launch_bounds__(256)
#endif
const_tiles_kernel(const ConstJob job) {
  __shared__ __align__(128) double s_zero[CT_SRC];
  __shared__ __align__(128) double s_nan[CT_SRC];
  RJP_STAMP_BEGIN(1)
  const_sources(s_zero, s_nan);
  const size_t nitems = const_items(job);
  const int nwarp = blockDim.x >> 5, wrp = threadIdx.x >> 5;
  for (size_t item = (size_t)wrp * gridDim.x + blockIdx.x; item < nitems;
       item += (size_t)gridDim.x * nwarp)
    const_item(job, item, s_zero, s_nan);
  const_drain();
  RJP_STAMP_END(1)
}

// Sparse tile exchange between slabs (multi-GPU): only the cube columns of jet-crossing rays
// travel.  pack: out[c][k] = cube[c][ray_ids[k]]; scatter: the inverse.  Lanes run along k;
// with a sorted ray list neighbouring k are neighbouring rays, so both sides coalesce.
__global__ void pack_rays_kernel(const double* __restrict__ cube, size_t plane,
                                 const int32_t* __restrict__ ray_ids, int n, int n_stride,
                                 int nchan, double* __restrict__ out) {
  const size_t total = (size_t)n * nchan;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / n), k = (int)(i - (size_t)c * n);
    out[(size_t)c * n_stride + k] = cube[(size_t)c * plane + ray_ids[k]];
  }
}

__global__ void scatter_rays_kernel(const double* __restrict__ in, int n_stride,
                                    const int32_t* __restrict__ ray_ids, int n, int nchan,
                                    double* __restrict__ cube, size_t plane) {
  const size_t total = (size_t)n * nchan;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / n), k = (int)(i - (size_t)c * n);
    cube[(size_t)c * plane + ray_ids[k]] = in[(size_t)c * n_stride + k];
  }
}

// Per-channel sum over the sky of a cube, NaN skipped (Pipeline's results['flux'] of a line
// run, classes.py:2468-2472): only the columns of the listed rays can hold anything but the
// constant 0 / NaN, so one CTA per channel gathers those (ordered list: neighbouring lanes read
// neighbouring rays) and reduces them in a fixed order.
__global__ void __launch_bounds__(256)
column_totals_kernel(const double* __restrict__ cube, size_t plane, size_t offset,
                     const int32_t* __restrict__ ray_list,
                     const int32_t* __restrict__ n_active_dev, double* __restrict__ totals) {
  __shared__ double s_w[8];
  const int n = *n_active_dev;
  const double* p = cube + (size_t)blockIdx.x * plane + offset;
  double sum = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double v = p[ray_list[k]];
    if (v == v) sum += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
    totals[blockIdx.x] = t;
  }
}

// Continuum-only walk: one warp per jet-crossing ray, lanes stride along the extent.  The
// number of listed rays is read on the device (the host never waits for it); warps stride
// over the list.
__global__ void __launch_bounds__(256)
continuum_rays_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                      const CellGrids ov,
                      const double2* __restrict__ cells, const int2* __restrict__ extents,
                      const int32_t* __restrict__ ray_list,
                      const int32_t* __restrict__ n_active_dev,
                      double* __restrict__ em, double* __restrict__ kff,
                      double* __restrict__ tsum, int32_t* __restrict__ tcount) {
  __shared__ Params s_p;
  stage_params(&s_p, m, ep, ov);
  const int lane = threadIdx.x & 31;
  const int n_active = *n_active_dev;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_active; w += nw) {
    const int ray = ray_list[w];
    const int xl = ray / m.nz, iz = ray - xl * m.nz;
    const int ix = m.x_lo + xl;
    const int2 ext = extents[ray];
    const Ray rc = ray_of(s_p.m, ix, iz);
    const double2* col = cells + (size_t)xl * m.ny * m.nz + iz;
    ContAcc a = {0.0, 0.0, 0.0, 0};
    for (int iy = ext.x + lane; iy < ext.y; iy += 32) {
      const double2 c = col[(size_t)iy * m.nz];
      if (empty_cell(c)) continue;
      accumulate(a, decode(c, s_p, rc, ix, iy, iz), ct.t_exponent);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.em += __shfl_xor_sync(0xffffffffu, a.em, o);
      a.kff += __shfl_xor_sync(0xffffffffu, a.kff, o);
      a.tsum += __shfl_xor_sync(0xffffffffu, a.tsum, o);
      a.cnt += __shfl_xor_sync(0xffffffffu, a.cnt, o);
    }
    if (lane == 0) {
      em[ray] = a.em * ct.em_scale;
      kff[ray] = a.kff * ct.tau_scale;
      tsum[ray] = a.tsum;
      tcount[ray] = a.cnt;
    }
  }
}

// The same walk for a BATCH of model times (a variable-ejection time series: Pipeline runs the
// continuum once per run year, classes.py:2347-2453; the model time enters only through the
// burst factor chi(time - travel time), classes.py:861-870).  blockIdx.y picks a block of
// EPOCH_BLOCK epochs, one warp walks a jet-crossing ray once for all of them: the cell load,
// the travel time and T^t_exponent are shared, each epoch costs its Gaussians and two FMAs.
// Per epoch the cells are added in the order continuum_rays_kernel adds them.  Only the
// pixels of the listed rays are written (the caller zero-fills: classes.py:1120, :1427 nansum).
#ifndef RJP_EPOCH_BLOCK
#define RJP_EPOCH_BLOCK 8
#endif
#ifndef RJP_EPOCH_MINB
#define RJP_EPOCH_MINB 2   // 128 registers, 16 warps per SM: 0.70 ms against 0.86 ms at 1 (236 registers)
#endif
constexpr int EPOCH_BLOCK = RJP_EPOCH_BLOCK;

__global__ void __launch_bounds__(256, RJP_EPOCH_MINB)
continuum_epochs_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                        const CellGrids ov, const double2* __restrict__ cells,
                        const int2* __restrict__ extents, const int32_t* __restrict__ ray_list,
                        const int32_t* __restrict__ n_active_dev,
                        const double* __restrict__ times, const int n_epochs,
                        double* __restrict__ em, double* __restrict__ kff,
                        double* __restrict__ tsum, int32_t* __restrict__ tcount) {
  __shared__ Params s_p;
  stage_params(&s_p, m, ep, ov);
  const rjp_model& M = s_p.m;
  const int lane = threadIdx.x & 31;
  const int n_active = *n_active_dev;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const size_t npix = (size_t)(M.x_hi - M.x_lo) * M.nz;
  const int e0 = blockIdx.y * EPOCH_BLOCK;
  double t_e[EPOCH_BLOCK];
#pragma unroll
  for (int j = 0; j < EPOCH_BLOCK; ++j) t_e[j] = times[min(e0 + j, n_epochs - 1)];
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_active; w += nw) {
    const int ray = ray_list[w];
    const int xl = ray / M.nz, iz = ray - xl * M.nz;
    const int ix = M.x_lo + xl;
    const int2 ext = extents[ray];
    const Ray rc = ray_of(M, ix, iz);
    const double2* col = cells + (size_t)xl * M.ny * M.nz + iz;
    double a_em[EPOCH_BLOCK], a_k[EPOCH_BLOCK];
#pragma unroll
    for (int j = 0; j < EPOCH_BLOCK; ++j) a_em[j] = a_k[j] = 0.0;
    double a_t = 0.0;
    int a_c = 0;
    for (int iy = ext.x + lane; iy < ext.y; iy += 32) {
      const double2 c = col[(size_t)iy * M.nz];
      if (empty_cell(c)) continue;
      const double ffw = signbit(c.y) ? 0.5 : 1.0;
      const double temp = fabs(c.y);
      const bool t_ok = temp > 0.0;
      if (t_ok) {
        a_t += temp;
        a_c += 1;
      }
      if (!(c.x > 0.0)) continue;
      // launch-time independent part of decode()
      const double y = __dadd_rn(corner(M.cs, iy, M.ny), M.cs / 2.0);
      const double r = __dadd_rn(__dmul_rn(M.sa, y), __dmul_rn(M.ca, rc.z1));
      double travel;
      if (s_p.ov.travel != nullptr) {
        travel = s_p.ov.travel[((size_t)xl * M.ny + iy) * M.nz + iz];
      } else if (M.qd_v == 0.0) {
        const double rad = (r_shifted(M, fabs(r)) + M.mr0 - M.r0) * M.au_m;
        const double e = 1.0 - M.q_v;
        travel = s_p.tt_cst * ((e == 1.0) ? rad : pow_call(rad, e)) - s_p.tt_f0;
      } else {
        travel = travel_slow(&s_p, ix, iy, iz);
      }
      double tp = 0.0;
      if (t_ok)
        tp = (ct.t_exponent == -1.5) ? 1.0 / (temp * sqrt(temp)) : pow_call(temp, ct.t_exponent);
      const rjp_burst* b = (r < 0.0) ? s_p.ep.red : s_p.ep.blue;
      const int nb = (r < 0.0) ? s_p.ep.n_red : s_p.ep.n_blue;
      double chi[EPOCH_BLOCK];
#pragma unroll
      for (int j = 0; j < EPOCH_BLOCK; ++j) chi[j] = 1.0;
      for (int i = 0; i < nb; ++i) {          // burst_chi() for the epochs of the block
        const double t0 = b[i].t0, amp = b[i].amp, q = b[i].inv2s2;
#pragma unroll
        for (int j = 0; j < EPOCH_BLOCK; ++j) {
          const double d = (t_e[j] - travel) - t0;
          // (skipping the Gaussians that are < 1e-26 was slower: divergence, 1.57 vs 1.36 ms)
          chi[j] += amp * exp(-(d * d) * q);
        }
      }
#pragma unroll
      for (int j = 0; j < EPOCH_BLOCK; ++j) {
        const double ne = c.x * chi[j];
        if (ne == ne) {                        // NaN travel time -> NaN density: dropped
          const double ne2 = ne * ne * ffw;
          a_em[j] += ne2;
          if (t_ok) a_k[j] += tp * ne2;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < EPOCH_BLOCK; ++j) {
        a_em[j] += __shfl_xor_sync(0xffffffffu, a_em[j], o);
        a_k[j] += __shfl_xor_sync(0xffffffffu, a_k[j], o);
      }
      a_t += __shfl_xor_sync(0xffffffffu, a_t, o);
      a_c += __shfl_xor_sync(0xffffffffu, a_c, o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < EPOCH_BLOCK; ++j) {
        if (e0 + j < n_epochs) {
          if (em) em[(size_t)(e0 + j) * npix + ray] = a_em[j] * ct.em_scale;
          kff[(size_t)(e0 + j) * npix + ray] = a_k[j] * ct.tau_scale;
        }
      }
      if (blockIdx.y == 0) {
        tsum[ray] = a_t;
        tcount[ray] = a_c;
      }
    }
  }
}

// Ordered compaction of the jet-crossing rays (extent non-empty) in two small launches:
// per-chunk counts, then every chunk re-derives its flags, adds the counts of the chunks
// before it and writes its rays in ascending order -- neighbouring CTAs of the ray kernels
// then write neighbouring cube columns, the sparse exchanges pack / scatter coalesced, and no
// sort is needed.  *n_active stays on the device: the ray kernels read it there.
constexpr int RL_CHUNK = 1024, RL_THREADS = 256;

__device__ __forceinline__ int block_sum_256(int v, int* s_w /* [8] */) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_w[wrp] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < RL_THREADS / 32; ++w) t += s_w[w];
  return t;
}

__global__ void __launch_bounds__(RL_THREADS)
ray_count_kernel(const int2* __restrict__ extents, int nray, int32_t* __restrict__ counts) {
  __shared__ int s_w[RL_THREADS / 32];
  const int base = blockIdx.x * RL_CHUNK;
  int n = 0;
#pragma unroll
  for (int k = 0; k < RL_CHUNK / RL_THREADS; ++k) {
    const int i = base + k * RL_THREADS + threadIdx.x;
    if (i < nray) {
      const int2 e = extents[i];
      n += e.x < e.y;
    }
  }
  n = block_sum_256(n, s_w);
  if (threadIdx.x == 0) counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(RL_THREADS)
ray_compact_kernel(const int2* __restrict__ extents, int nray,
                   const int32_t* __restrict__ counts, int32_t* __restrict__ list,
                   int32_t* __restrict__ n_active) {
  __shared__ int s_w[RL_THREADS / 32];
  __shared__ int s_run[RL_THREADS / 32];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  // rays listed by the chunks before this one (the last chunk also publishes the total)
  int before = 0;
  for (int c = threadIdx.x; c < (int)blockIdx.x; c += RL_THREADS) before += counts[c];
  before = block_sum_256(before, s_w);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_active = before + counts[blockIdx.x];
  const int base = blockIdx.x * RL_CHUNK;
  for (int k = 0; k < RL_CHUNK / RL_THREADS; ++k) {
    const int i = base + k * RL_THREADS + threadIdx.x;
    bool on = false;
    if (i < nray) {
      const int2 e = extents[i];
      on = e.x < e.y;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    __syncthreads();
    if (lane == 0) s_run[wrp] = __popc(bal);
    __syncthreads();
    int off = before, tot = 0;
#pragma unroll
    for (int w = 0; w < RL_THREADS / 32; ++w) {
      if (w < wrp) off += s_run[w];
      tot += s_run[w];
    }
    if (on) list[off + __popc(bal & ((1u << lane) - 1u))] = i;
    before += tot;
  }
}

// Phase 2a of the channel loop: the fast-class cells [i0, i1) of the batch added to this
// thread's GCH channels.  The wing term is evaluated branch-free for all channels (8
// independent chains per thread); channels inside the Gaussian core replace it by the table
// evaluation.  (A second copy of this loop specialised for y <= 0.03 -- no y^5 term, shorter
// cosine: 7 instructions less per core evaluation, 4.61 instead of 4.79 ms when used for every
// cell -- made the kernel SLOWER when both copies were present, 4.87 ms: the loop body is
// 10 KB of code and the instruction cache does not hold two of them.)
template <bool UNI, int GCH>
__device__ __forceinline__ void fast_cells(const FastEntry* __restrict__ s_fast, int i0, int i1,
                                           double (&acc)[GCH], const double* dn,
                                           const f32x2* dnf2, double dstep, float dstepf,
                                           uint32_t tab) {
  for (int i = i0; i < i1; ++i) {
    const FastEntry fe = s_fast[i];
    const WingCoef2 wc = vt_wing_coef2(fe);
    const f32x2 b1 = pk2(fe.b1, fe.b1);
    const double X0 = fma(dn[0], fe.inv, fe.xs), dX = dstep * fe.inv;   // UNI only
#pragma unroll
    for (int j = 0; j < GCH; j += 2) {
      // two channels per step: the fp32 work runs as packed FFMA2
      double Xa, Xb;
      if constexpr (UNI) {
        Xa = fma((double)j, dX, X0);
        Xb = fma((double)(j + 1), dX, X0);
      } else {
        Xa = fma(dn[j], fe.inv, fe.xs);
        Xb = fma(dn[j + 1], fe.inv, fe.xs);
      }
      const double X2a = Xa * Xa, X2b = Xb * Xb;
      const bool corea = __double2hiint(X2a) < fe.xc2_hi, coreb = __double2hiint(X2b) < fe.xc2_hi;
      const double ra = rcp_seed(X2a), rb = rcp_seed(X2b);
      double leada = fe.w0 * (ra * fma(-X2a, ra, 2.0));
      double leadb = fe.w0 * (rb * fma(-X2b, rb, 2.0));
      f32x2 k2 = vt_wing_poly2(wc, pk2(d2f_trunc_pos(ra), d2f_trunc_pos(rb)));
      if (corea || coreb) {
        float ka, kb;
        upk2(k2, ka, kb);
        if (corea) {
          ka = vt_core(fe, tab, Xa, X2a);
          leada = fe.a0;
        }
        if (coreb) {
          kb = vt_core(fe, tab, Xb, X2b);
          leadb = fe.a0;
        }
        k2 = pk2(ka, kb);
      }
      f32x2 d2;
      if constexpr (UNI)
        d2 = fma2(pk2((float)j, (float)(j + 1)), pk2(dstepf, dstepf), dnf2[0]);
      else
        d2 = dnf2[j >> 1];
      // (1 - exp(-h nu / kT)) / p0 = 1 + dn b1 (fast class: the next term is < 1e-9)
      float ka, kb;
      upk2(fma2(k2, mul2(d2, b1), k2), ka, kb);
      acc[j] = fma(leada, f2d_pos(ka), acc[j]);
      acc[j + 1] = fma(leadb, f2d_pos(kb), acc[j + 1]);
    }
  }
}

// ------------------------------------------------------------------ K4a: prepare the rays
// What the channel loop needs of a jet-crossing ray, written by ray_prepare_kernel: where its
// prepared cells sit in the entry buffer (fast-class entries from the front of the ray's
// segment, the others from its back) and the ray's continuum sums for the flux epilogue.
struct RayMeta {
  long long base;     // first entry of the ray's segment
  int len;            // length of the segment (= length of the ray's extent)
  int nf, ns;         // fast-class / fp64-class entries
  int cnt;            // cells with a valid temperature (classes.py:1471-1472)
  double kray, tsum;  // K (tau_ff = cff * K) and the sum of those temperatures
};
static_assert(sizeof(RayMeta) == 40, "RayMeta layout");

// One CTA (one warp) per jet-crossing ray: every thread prepares one cell of the extent per batch --
// burst factor, Doppler shift, widths, amplitude, class (decode / make_entry / to_fast: ~1500
// dependent fp64 instructions per cell, latency-bound) -- and stores the 80-byte entry; the same
// walk yields the ray's continuum sums (EM, K, sum T, count) and writes its pixels of the four
// sky images.  This used to be the first phase of every batch INSIDE the channel loop kernel,
// where its stalled warps took a third of that kernel's warp residency (ncu: 35 % of the stall
// samples on 16 % of the instructions); as a kernel of its own it is hidden behind thousands of
// independent rays, and the channel loop kernel is nothing but the channel loop.
// one warp per ray, 32 CTAs per SM: most rays have fewer than 64 cells, and what this kernel
// needs is many independent warps (measured: 64 threads x 8 CTAs 0.71 ms, x 16 0.57, this 0.5)
#ifndef RJP_PREP_MINB
#define RJP_PREP_MINB 32
#endif
#ifndef RJP_PREP_THREADS
#define RJP_PREP_THREADS 32
#endif
__global__ void __launch_bounds__(RJP_PREP_THREADS, RJP_PREP_MINB)
ray_prepare_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                   const CellGrids ov, const rjp_line ln, const double dn_max,
                   const double2* __restrict__ cells, const int2* __restrict__ extents,
                   const int32_t* __restrict__ ray_list,
                   const int32_t* __restrict__ n_active_dev, const int t0, const int t1,
                   unsigned long long* __restrict__ cursor, const long long capacity,
                   unsigned char* __restrict__ entries, RayMeta* __restrict__ meta,
                   double* __restrict__ em, double* __restrict__ kff,
                   double* __restrict__ tsum, int32_t* __restrict__ tcount) {
  __shared__ Params s_p;
  __shared__ rjp_line s_ln;
  __shared__ double s_part[3][2];
  __shared__ int s_pcnt[2];
  __shared__ int s_woff[2][3];
  __shared__ long long s_base;
  stage_params(&s_p, m, ep, ov);
  if (threadIdx.x == 0) s_ln = ln;
  __syncthreads();
  const int NT = blockDim.x;
  const int g = threadIdx.x, lane = g & 31, wrp = g >> 5, nwarps = NT >> 5;
  const int n_active = min(*n_active_dev, t1);     // this launch: rays [t0, t1) of the list
  for (int ticket = t0 + blockIdx.x; ticket < n_active; ticket += gridDim.x) {
    const int ray = ray_list[ticket];              // slab-local ray index = xl * nz + iz
    const int2 ext = extents[ray];
    const int len = ext.y - ext.x;
    if (g == 0) {
      long long b = (long long)atomicAdd(cursor, (unsigned long long)len);
      if (b + len > capacity) b = -1;              // (cannot happen with an exact capacity)
      s_base = b;
    }
    const int xl = ray / m.nz, iz = ray - xl * m.nz;
    const int ix = m.x_lo + xl;
    const Ray rc = ray_of(s_p.m, ix, iz);
    const double2* col = cells + (size_t)xl * m.ny * m.nz + iz;
    ContAcc ca = {0.0, 0.0, 0.0, 0};
    int nf_tot = 0, ns_tot = 0;
    __syncthreads();
    const long long base = s_base;
    for (int y0 = ext.x; y0 < ext.y; y0 += NT) {
      const int iy = y0 + g;
      LineEntry e;
      e.amp = 0.0;
      if (iy < ext.y) {
        const double2 c = col[(size_t)iy * m.nz];
        if (!empty_cell(c)) {
          const Decoded d = decode(c, s_p, rc, ix, iy, iz);
          accumulate(ca, d, ct.t_exponent);
          if (d.ne_ok && d.t_ok)
            e = make_entry(d, s_p.m, s_ln, dn_max, ix, iy, iz, s_p.ov.vlos);
        }
      }
      // cells that emit no line are dropped; the two classes are compacted in cell order
      // (deterministic summation order in the channel loop)
      const bool fast = fast_class(e);
      const bool slow = e.amp != 0.0 && !fast;
      const unsigned balf = __ballot_sync(0xffffffffu, fast);
      const unsigned bals = __ballot_sync(0xffffffffu, slow);
      if (lane == 0) {
        s_woff[0][wrp] = __popc(balf);
        s_woff[1][wrp] = __popc(bals);
      }
      __syncthreads();
      int basef = 0, bases = 0, nf = 0, ns = 0;
      for (int w = 0; w < nwarps; ++w) {
        if (w < wrp) {
          basef += s_woff[0][w];
          bases += s_woff[1][w];
        }
        nf += s_woff[0][w];
        ns += s_woff[1][w];
      }
      const unsigned below = (1u << lane) - 1u;
      if (base >= 0) {
        if (fast) {
          const FastEntry fe = to_fast(e);
          const long long k = base + nf_tot + basef + __popc(balf & below);
          reinterpret_cast<FastEntry*>(entries)[k] = fe;
        }
        if (slow) {
          const long long k = base + len - 1 - (ns_tot + bases + __popc(bals & below));
          reinterpret_cast<LineEntry*>(entries)[k] = e;
        }
      }
      nf_tot += nf;
      ns_tot += ns;
      __syncthreads();   // s_woff is rewritten by the next batch
    }
    // the ray's continuum sums: warp shuffles, then the warps' partials through shared memory
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ca.kff += __shfl_xor_sync(0xffffffffu, ca.kff, o);
      ca.tsum += __shfl_xor_sync(0xffffffffu, ca.tsum, o);
      ca.em += __shfl_xor_sync(0xffffffffu, ca.em, o);
      ca.cnt += __shfl_xor_sync(0xffffffffu, ca.cnt, o);
    }
    if (lane == 0) {
      s_part[0][wrp] = ca.kff;
      s_part[1][wrp] = ca.tsum;
      s_part[2][wrp] = ca.em;
      s_pcnt[wrp] = ca.cnt;
    }
    __syncthreads();
    if (g == 0) {
      double kray = 0.0, ts = 0.0, emr = 0.0;
      int cn = 0;
      for (int i = 0; i < nwarps; ++i) {
        kray += s_part[0][i];
        ts += s_part[1][i];
        emr += s_part[2][i];
        cn += s_pcnt[i];
      }
      kray *= ct.tau_scale;
      em[ray] = emr * ct.em_scale;
      kff[ray] = kray;
      tsum[ray] = ts;
      tcount[ray] = cn;
      RayMeta rm;
      rm.base = base;
      rm.len = len;
      rm.nf = base >= 0 ? nf_tot : 0;
      rm.ns = base >= 0 ? ns_tot : 0;
      rm.cnt = base >= 0 ? cn : -1;       // -1: entry buffer too small, the ray's columns = NaN
      rm.kray = kray;
      rm.tsum = ts;
      meta[ticket] = rm;
    }
    __syncthreads();   // s_base / s_part are rewritten by the next ray
  }
}

// ------------------------------------------------------------------ K4b: the channel loop
// UNI: the channels are equally spaced (rjp_line.chan_step != 0): the channel offsets are
// formed on the fly from two per-thread scalars instead of living in 24 registers (16 of
// which spilled), which is what lets the kernel fit more warps per SM.
// GCH = channels per thread (8 for whole cubes; 4 / 2 when a rank of a channel-sharded run
// owns only 128 / 64 channels, so that a one-warp CTA still covers them without idle slots).
// The list length is read on the device (the host never has to know how many rays cross the
// jet); see the ray loop for how the rays are dealt to the CTAs.
template <int MAXT, int MINB, bool UNI, int GCH>
#ifdef RJP_LINE_MAXNREG
__global__ void __maxnreg__(RJP_LINE_MAXNREG)
#else
__global__ void __launch_bounds__(MAXT, MINB)
#endif
integrate_line_kernel(const rjp_line ln, const rjp_channels ch, const int nchan,
                      const int c_first, const int contsub, const double dn_max,
                      const unsigned char* __restrict__ entries,
                      const RayMeta* __restrict__ meta,
                      const int32_t* __restrict__ ray_list,
                      const int32_t* __restrict__ n_active_dev, const int t0, const int t1,
                      double* __restrict__ tau_rrl, double* __restrict__ flux_rrl,
                      const size_t plane, const size_t cube_offset) {
  // one chunk of prepared cells (80 bytes each; +3 zero-amplitude pads for the 4-cell batches
  // of the fp64 path)
  __shared__ __align__(16) unsigned char s_raw[(MAXT + 4) * sizeof(LineEntry)];
  __shared__ float4 s_tab[VT_TAB_F4];
  static_assert(sizeof(LineEntry) == sizeof(FastEntry), "shared batch buffer");
  static_assert(sizeof(LineEntry) == 80, "entries are copied as five 16-byte words");
  RJP_STAMP_BEGIN(0)
  for (int i = threadIdx.x; i < VT_TAB_F4; i += blockDim.x)
    s_tab[i] = reinterpret_cast<const float4*>(g_vt_core)[i];

  uint32_t tab = (uint32_t)__cvta_generic_to_shared(s_tab);
  asm volatile("" : "+r"(tab));  // keep it in a register (rematerialising costs S2R + LEA)
  const int NT = blockDim.x;
  const int g = threadIdx.x;
  const int n_active = min(*n_active_dev, t1);     // this launch: rays [t0, t1) of the list

  // thread g owns channels g, g + NT, g + 2 NT, ...: at every step the lanes of a warp hold
  // 32 neighbouring channels, i.e. nearly the same point of the line profile, so the
  // core / wing branches below are (almost always) warp-uniform
  static_assert(GCH % 2 == 0, "channels are processed in pairs");
  // channel offsets nu_k - nu0 of this thread's channels g + j NT
  double dn[UNI ? 1 : GCH];
  f32x2 dnf2[UNI ? 1 : GCH / 2];
  const double dstep = ln.chan_step * (double)NT;        // UNI: dn_j = dn[0] + j dstep
  float dstepf = (float)dstep;
  asm volatile("" : "+f"(dstepf));   // keep it in a register (rematerialising costs 3 slots)
  if constexpr (UNI) {
    dn[0] = ln.chan_dnu0 + (double)(c_first + g) * ln.chan_step;
    dnf2[0] = pk2((float)dn[0], (float)dn[0]);
  } else {
#pragma unroll
    for (int j = 0; j < GCH; ++j)
      dn[j] = (g + j * NT < nchan) ? __ldg(ch.dnu + g + j * NT) : 0.0;
#pragma unroll
    for (int j = 0; j < GCH; j += 2) dnf2[j >> 1] = pk2((float)dn[j], (float)dn[j + 1]);
  }

  __syncthreads();
  // ray number blockIdx.x, + gridDim.x, ... of the ordered list.  The host sizes the grid to
  // one CTA per ray when it remembers the count of this geometry (rjp_integrate's
  // n_active_hint), else to a multiple of what fits on the device: CTAs beyond the list leave
  // at once, a long list gives every CTA a few rays.  Measured on B200: a grid of
  // exactly-resident CTAs pulling rays from a ticket counter is 7 % SLOWER (5.61 vs 5.25 ms at
  // 1024^2 rays x 512 channels) -- CTAs of one age advance in lockstep, CTAs of mixed ages
  // overlap their start-up and epilogue latencies with the others' channel loop.
  for (int ticket = t0 + blockIdx.x; ticket < n_active; ticket += gridDim.x) {
  const RayMeta rm = meta[ticket];
  const int ray = ray_list[ticket];                // slab-local ray index = xl * nz + iz
  double acc[GCH];
#pragma unroll
  for (int j = 0; j < GCH; ++j) acc[j] = 0.0;
  const uint4* seg = reinterpret_cast<const uint4*>(entries) + rm.base * 5;

  // fast class: chunks of NT prepared cells, one 80-byte entry per thread into shared memory,
  // then every thread adds each of them to its GCH channels
  for (int c0 = 0; c0 < rm.nf; c0 += NT) {
    const int cnt = (rm.nf - c0 < NT) ? rm.nf - c0 : NT;
    if (g < cnt) {
      const uint4* src = seg + (size_t)(c0 + g) * 5;
      uint4* dst = reinterpret_cast<uint4*>(s_raw) + g * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) dst[k] = __ldg(src + k);
    }
    __syncthreads();
    fast_cells<UNI, GCH>(reinterpret_cast<const FastEntry*>(s_raw), 0, cnt, acc, dn, dnf2,
                         dstep, dstepf, tab);
    __syncthreads();
  }

  // everything else in fp64 (large or tiny y, steep Planck factor), stored from the back of
  // the segment: 4 cells at a time, channel loop outermost
  for (int c0 = 0; c0 < rm.ns; c0 += NT) {
    const int cnt = (rm.ns - c0 < NT) ? rm.ns - c0 : NT;
    LineEntry* s_slow = reinterpret_cast<LineEntry*>(s_raw);
    if (g < cnt) {
      const uint4* src = seg + (size_t)(rm.len - 1 - (c0 + g)) * 5;
      uint4* dst = reinterpret_cast<uint4*>(s_raw) + g * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) dst[k] = __ldg(src + k);
    }
    __syncthreads();
    if (g < 3) {  // pad the last batch of 4 with zero-amplitude copies
      LineEntry pad = s_slow[cnt - 1];
      pad.amp = 0.0;
      s_slow[cnt + g] = pad;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < GCH; ++j) {
      if (g + j * NT >= nchan) break;
      const double dnj = UNI ? fma((double)j, dstep, dn[0]) : dn[j];
      double sum = 0.0;
      for (int i = 0; i < cnt; i += 4) {
        const LineEntry* eb = s_slow + i;
        double x[4], yy[4], w[4];
        bool wing = true;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          x[v] = fma(dnj, eb[v].inv_s2, eb[v].xs);
          yy[v] = eb[v].y;
          wing = wing && (fma(x[v], x[v], yy[v] * yy[v]) >= 36.0);
        }
        if (wing) faddeeva_wing_n<4>(x, yy, w);
        else faddeeva_re_n<4>(x, yy, w);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          sum = fma(eb[v].amp * w[v], planck_factor(eb[v], dnj), sum);
      }
      acc[j] += sum;
    }
    __syncthreads();
  }

  // K5 epilogue: rrls.py:444-447, physics.py:571-574, classes.py:1323-1328, :1484-1488
  const int cn = rm.cnt;
  const double kray = rm.kray;
  const double tmean = rm.tsum / (double)cn;
  // exp(h nu_c / k Tmean) = exp(h nu0 / k Tmean) exp(x), x = h (nu_c - nu0) / k Tmean: one exp
  // per ray and a cubic when |x| <= 1e-3 for every channel (x^4 / 24 <= 4e-14)
  const double hkm = ln.h_over_k / tmean;
  const bool bnu_taylor = hkm * dn_max <= 1e-3;
  const double e_nu0 = bnu_taylor ? exp(hkm * ln.nu0) : 0.0;
#pragma unroll
  for (int j = 0; j < GCH; ++j) {
    const int c = g + j * NT;
    if (c >= nchan) break;
    if (tau_rrl) st_column(tau_rrl + (size_t)c * plane + cube_offset + ray,
                           cn >= 0 ? acc[j] : dnan());
    if (flux_rrl) {
      double s = dnan();
      if (cn > 0) {
        const double tc = __ldg(ch.cff + c) * kray;
        const double ec = exp(-tc);
        double e_nu;
        if (bnu_taylor) {
          const double x = hkm * (UNI ? fma((double)j, dstep, dn[0]) : dn[j]);
          e_nu = e_nu0 * fma(x, fma(x, fma(x, 1.0 / 6.0, 0.5), 1.0), 1.0);
        } else {
          e_nu = exp(ln.h_over_k * __ldg(ch.nu + c) / tmean);
        }
        const double bnu = __ldg(ch.bnu + c) / (e_nu - 1.0);
        s = bnu * ec * (1.0 - exp(-acc[j]));
        if (!contsub) s += __ldg(ch.aff + c) * (tmean * (1.0 - ec));
      }
      st_column(flux_rrl + (size_t)c * plane + cube_offset + ray, s);
    }
  }
  }  // ray loop
  RJP_STAMP_END(0)
}

// Line-of-sight means of the per-cell properties (what the reference's model plot shows,
// plotting/functions.py:539-590: np.nanmean(field, axis=los) of number density, temperature,
// ionisation fraction and v_los - v_lsr) straight from the ray walk, so that the plot does not
// need four full 3-D grids on the host.  One warp per jet-crossing ray; out[q][ray] for
// q = 0..6: mean n, mean T, mean x, mean (v_los - v_lsr), min n, max n, max T; NaN where a
// ray has no finite value (and on every ray that misses the jet: prefilled by the caller).
__global__ void __launch_bounds__(256)
los_means_kernel(const rjp_model m, const rjp_epoch ep, const uint8_t* __restrict__ nverts,
                 const int2* __restrict__ extents, const int32_t* __restrict__ ray_list,
                 const int32_t* __restrict__ n_active_dev, const size_t nray,
                 double* __restrict__ out) {
  __shared__ Params s_p;
  stage_params(&s_p, m, ep);
  const int lane = threadIdx.x & 31;
  const int n_active = *n_active_dev;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_active; w += nw) {
  const int ray = ray_list[w];
  const int xl = ray / m.nz, iz = ray - xl * m.nz;
  const int ix = m.x_lo + xl;
  const int2 ext = extents[ray];
  const uint8_t* col = nverts + (size_t)xl * m.ny * m.nz + iz;
  double sum[4] = {0.0, 0.0, 0.0, 0.0};
  int cnt[4] = {0, 0, 0, 0};
  double nmin = CUDART_INF, nmax = -CUDART_INF, tmax = -CUDART_INF;
  for (int iy = ext.x + lane; iy < ext.y; iy += 32) {
    if (col[(size_t)iy * m.nz] == 0) continue;
    const Rw g = centroid_rw(s_p.m, ix, iy, iz);
    const Laws l = laws_of(s_p.m, g, false);
    const double tl = s_p.ep.time - travel_time(s_p.m, g);
    const double chi = (g.r < 0.0) ? burst_chi(s_p.ep.red, s_p.ep.n_red, tl)
                                   : burst_chi(s_p.ep.blue, s_p.ep.n_blue, tl);
    const double v[4] = {l.nd * chi, l.temp, l.xi, velocity_of(s_p.m, g).vlos_rel};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (v[q] == v[q]) {
        sum[q] += v[q];
        cnt[q] += 1;
      }
    }
    if (v[0] == v[0]) {
      nmin = fmin(nmin, v[0]);
      nmax = fmax(nmax, v[0]);
    }
    if (v[1] == v[1]) tmax = fmax(tmax, v[1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sum[q] += __shfl_xor_sync(0xffffffffu, sum[q], o);
      cnt[q] += __shfl_xor_sync(0xffffffffu, cnt[q], o);
    }
    nmin = fmin(nmin, __shfl_xor_sync(0xffffffffu, nmin, o));
    nmax = fmax(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  }
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q * nray + ray] = cnt[q] > 0 ? sum[q] / cnt[q] : dnan();
    out[4 * nray + ray] = cnt[0] > 0 ? nmin : dnan();
    out[5 * nray + ray] = cnt[0] > 0 ? nmax : dnan();
    out[6 * nray + ray] = cnt[1] > 0 ? tmax : dnan();
  }
  }
}

// Diagnostic: Re w(x + iy) with the channel loop's own device routines
__global__ void voigt_profile_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                     int64_t n, double* __restrict__ out) {
  __shared__ float4 s_tab[VT_TAB_F4];
  for (int i = threadIdx.x; i < VT_TAB_F4; i += blockDim.x)
    s_tab[i] = reinterpret_cast<const float4*>(g_vt_core)[i];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = voigt_any((uint32_t)__cvta_generic_to_shared(s_tab), x[i], y[i]);
}

// ------------------------------------------------------------------ continuum images
__global__ void continuum_images_kernel(const double* __restrict__ kff,
                                        const double* __restrict__ tsum,
                                        const int32_t* __restrict__ tcount, int64_t npix,
                                        const double* __restrict__ cff,
                                        const double* __restrict__ iff, double omega_jy,
                                        int nfreq, double* __restrict__ tau,
                                        double* __restrict__ inten, double* __restrict__ flux) {
  // blockIdx.y: one K plane of a batch of epochs (rjp_continuum_images_epochs), outputs
  // [plane][frequency][pixel]; tsum / tcount do not depend on the epoch
  kff += (size_t)blockIdx.y * npix;
  const size_t out0 = (size_t)blockIdx.y * nfreq * npix;
  if (tau) tau += out0;
  if (inten) inten += out0;
  if (flux) flux += out0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double k = kff[p];
    const int c = tcount[p];
    const double tmean = (c > 0) ? tsum[p] / (double)c : dnan();
    for (int f = 0; f < nfreq; ++f) {
      const double t = cff[f] * k;                               // classes.py:1427-1432
      if (tau) tau[(size_t)f * npix + p] = t;
      if (inten || flux) {
        const double tb = tmean * (1.0 - exp(-t));               // classes.py:1484-1486
        const double in = iff[f] * tb;                           // classes.py:1488
        if (inten) inten[(size_t)f * npix + p] = in;
        if (flux) flux[(size_t)f * npix + p] = in * omega_jy;    // classes.py:1531-1533
      }
    }
  }
}

}  // namespace rjp

using namespace rjp;

extern "C" int rjp_launch_ray_list(const int32_t* extents, int nray, int32_t* list,
                                   int32_t* chunk_counts, int32_t* n_active,
                                   cudaStream_t stream) {
  if (nray <= 0) {
    cudaMemsetAsync(n_active, 0, sizeof(int32_t), stream);
    return RJP_OK;
  }
  const int chunks = (nray + RL_CHUNK - 1) / RL_CHUNK;
  ray_count_kernel<<<chunks, RL_THREADS, 0, stream>>>(reinterpret_cast<const int2*>(extents),
                                                      nray, chunk_counts);
  ray_compact_kernel<<<chunks, RL_THREADS, 0, stream>>>(reinterpret_cast<const int2*>(extents),
                                                        nray, chunk_counts, list, n_active);
  return RJP_OK;
}

extern "C" int rjp_ray_list_chunk(void) { return RL_CHUNK; }

// ---------------------------------------------------------------- launch plumbing
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

struct DeviceInfo { int sms; bool carved; };

// Per-device facts and one-off function attributes, guarded for host threads that drive
// different devices at the same time.
static DeviceInfo& device_info() {
  static std::mutex mu;
  static DeviceInfo info[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  DeviceInfo& d = info[dev & 63];
  std::lock_guard<std::mutex> lock(mu);
  if (d.sms == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    d.sms = sms;
  }
  return d;
}

template <int T, int B, bool U, int G>
static void carve_line(int pct) {
  cudaFuncSetAttribute(integrate_line_kernel<T, B, U, G>,
                       cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// Same shared-memory carve-out for the kernels that are meant to be co-resident: an SM only
// switches its L1 / shared split when it is idle, so a kernel that asks for a different split
// waits until the other one has drained (measured: the channel loop started 0.9 ms late
// behind a constant writer with a different split).
static void set_carveouts() {
  static std::mutex mu;
  DeviceInfo& d = device_info();
  std::lock_guard<std::mutex> lock(mu);
  if (d.carved) return;
  const int pct = 75;
  cudaFuncSetAttribute(const_tiles_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(continuum_rays_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(ray_prepare_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  carve_line<32, 16, true, 2>(pct);
  carve_line<32, 16, true, 4>(pct);
  carve_line<32, 16, true, 8>(pct);
  carve_line<64, RJP_MINB64U, true, GCH_MAX>(pct);
  carve_line<128, RJP_MINB128, true, GCH_MAX>(pct);
  carve_line<LINE_THREADS, 2, true, GCH_MAX>(pct);
  carve_line<64, RJP_MINB64, false, GCH_MAX>(pct);
  carve_line<128, RJP_MINB128, false, GCH_MAX>(pct);
  carve_line<LINE_THREADS, 2, false, GCH_MAX>(pct);
  d.carved = true;
}

// fork / join events of the two-stream pass, created once per (host thread, device)
static int pass_events(cudaEvent_t* fork, cudaEvent_t* join) {
  thread_local cudaEvent_t ev[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaEvent_t* e = ev[dev & 63];
  for (int i = 0; i < 2; ++i)
    if (e[i] == nullptr &&
        cudaEventCreateWithFlags(&e[i], cudaEventDisableTiming) != cudaSuccess)
      return RJP_ERR_CUDA;
  *fork = e[0];
  *join = e[1];
  return RJP_OK;
}

// bulk stores need 16-byte aligned global addresses and sizes
static bool bulk_ok(const double* tau, const double* flux, size_t plane, size_t offset,
                    size_t nray) {
  if (env_int("RJP_NO_BULK", 0)) return false;
  return (plane % 2 == 0) && (offset % 2 == 0) && (nray % 2 == 0) &&
         (reinterpret_cast<uintptr_t>(tau) % 16 == 0) &&
         (reinterpret_cast<uintptr_t>(flux) % 16 == 0);
}

// byte offset of the entry buffer inside the line-pass scratch (cursor, RayMeta[nray], entries)
static long long line_entries_offset(long long nray) {
  return (64 + nray * (long long)sizeof(RayMeta) + 255) / 256 * 256;
}

extern "C" long long rjp_launch_line_scratch_bytes(long long nray, long long max_cells) {
  return line_entries_offset(nray) + max_cells * (long long)sizeof(LineEntry) + 256;
}

extern "C" int rjp_launch_integrate(const rjp_model* m, const rjp_epoch* ep,
                                    const rjp_continuum* ct, const rjp_cell* cells,
                                    const int32_t* extents, const int32_t* ray_list,
                                    const int32_t* n_active, int n_hint, double* em,
                                    double* kff, double* tsum, int32_t* tcount,
                                    const rjp_line* ln, const rjp_channels* ch, int nchan,
                                    int contsub, double dn_max, double* tau_rrl,
                                    double* flux_rrl, long long cube_plane,
                                    long long cube_offset, const double* travel_cells,
                                    const double* vlos_cells, void* scratch,
                                    long long max_cells, cudaStream_t stream,
                                    cudaStream_t stream2) {
  const CellGrids ov = {travel_cells, vlos_cells};
  const int nxs = m->x_hi - m->x_lo;
  const long long ctas = (long long)nxs * ((m->nz + ZT - 1) / ZT);
  if (ctas <= 0 || ctas > 2147483647LL) return RJP_ERR_ARG;
  const double2* c4 = reinterpret_cast<const double2*>(cells);
  const int2* ex2 = reinterpret_cast<const int2*>(extents);
  const bool lines = nchan > 0 && ln != nullptr;
  if (extents == nullptr || ray_list == nullptr || n_active == nullptr) {
    // no extents: dense sweep over the whole state (continuum only; the API demands extents
    // for line passes)
    if (lines) return RJP_ERR_ARG;
    integrate_continuum_kernel<2><<<(unsigned)ctas, 256, 0, stream>>>(*m, *ep, *ct, ov, c4, em,
                                                                     kff, tsum, tcount);
    return RJP_OK;
  }
  const size_t nray = (size_t)nxs * m->nz;
  const size_t plane = cube_plane > 0 ? (size_t)cube_plane : nray;
  const size_t coff = cube_plane > 0 ? (size_t)cube_offset : 0;
  if (plane < coff + nray) return RJP_ERR_ARG;
  set_carveouts();
  const int sms = device_info().sms;
  const bool bulk = bulk_ok(tau_rrl, flux_rrl, plane, coff, nray);
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaStream_t ls = stream;
  if (stream2 != nullptr && stream2 != stream) {
    // fork (before the constant fill is queued): the ray walk on stream2 runs beside the
    // write-bound constant fill on stream
    if (pass_events(&fork, &join) != RJP_OK) return RJP_ERR_CUDA;
    if (cudaEventRecord(fork, stream) != cudaSuccess ||
        cudaStreamWaitEvent(stream2, fork, 0) != cudaSuccess)
      return RJP_ERR_CUDA;
    ls = stream2;
  }
  ConstJob job = {ex2, nray, lines ? nchan : 0, em, kff, tsum, tcount, tau_rrl, flux_rrl,
                  plane, coff, 0, 0, bulk ? 1 : 0};
  // constants of the rays that miss the jet: one warp per SM streams whole constant tiles
  // with TMA bulk stores; launched first so that it is resident beside the ray kernels
  if (!env_int("RJP_SKIP_WRITER", 0)) {   // (debug knob: time the ray kernels alone)
    // beside a channel loop: few CTAs of several warps, so that only a few SMs give up a
    // slot of the (register-bound) channel loop; alone: one CTA of 4 warps on every SM.
    // The slabs of a sharded run differ: the outer ones hold mostly sky (the writer is the
    // longer of the two), the inner ones nearly none.  The writer is sized to end with the
    // ray kernels.  Measured (tools/slab_probe.py, tools/pass_probe.py): one writer CTA of 4
    // warps sustains ~57 GB/s; preparation + channel loop cost ~1.05e-6 ms per cell of the
    // listed rays' extents at 512 channels; a writer CTA on every SM slows the loop by ~1/3.
    int ctas_auto = sms / 4;
    if (lines) {
      // cells of the listed rays: the caller's count (line_max_cells is the summed length of
      // their extents), else ~63 per ray (BASELINE jet); n_hint = 0 is a count, not "unknown"
      const double rays = (double)(n_hint >= 0 ? n_hint : (int)(nray / 16));
      const double cells = (max_cells > 0 && n_hint >= 0) ? (double)max_cells : 63.0 * rays;
      const double t_loop_ms = 0.05 + 1.05e-6 * cells * (double)nchan / 512.0;
      const double cta_ms = 16.0 * (double)nray * (double)nchan / 57e6;   // writer work
      double want = cta_ms / t_loop_ms;
      for (int it = 0; it < 3; ++it)
        want = cta_ms / (t_loop_ms * (1.0 + 0.32 * (want < sms ? want : sms) / sms));
      if (want > ctas_auto) ctas_auto = want < sms ? (int)want : sms;
    }
    const int ctas = env_int("RJP_WRITER_CTAS", lines ? ctas_auto : sms);
    const int warps = env_int("RJP_WRITER_WARPS", 4);
    const size_t ntiles = (nray + CT_TILE - 1) / CT_TILE;
    const size_t items = ntiles * (size_t)(lines ? (nchan + RJP_CT_CG - 1) / RJP_CT_CG : 1);
    size_t grid = (size_t)(ctas > 0 ? ctas : 1);
    if (grid * warps > items) grid = (items + warps - 1) / warps;
    const_tiles_kernel<<<(unsigned)grid, 32 * warps, 0, stream>>>(job);
  }
  if (lines && !env_int("RJP_SKIP_LINES", 0)) {
    // scratch layout: [0, 64) cursor of the entry buffer, then one RayMeta per slab ray, then
    // max_cells entries of 80 bytes
    if (scratch == nullptr || max_cells < 0) return RJP_ERR_ARG;
    unsigned char* sc = static_cast<unsigned char*>(scratch);
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(sc);
    RayMeta* meta = reinterpret_cast<RayMeta*>(sc + 64);
    unsigned char* entries = sc + line_entries_offset((long long)nray);
    if (cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), ls) != cudaSuccess)
      return RJP_ERR_CUDA;
    // one CTA per listed ray when the caller knows their number, else a multiple of what fits
    const size_t grid_all = n_hint > 0 ? ((size_t)n_hint < nray ? (size_t)n_hint : nray)
                            : n_hint == 0 ? (size_t)sms
                                          : (size_t)sms * 8 * (size_t)env_int("RJP_GRID_FACTOR", 16);
    // K4b for the rays [t0, t1) of the list: channel blocks of at most 8 * 256 channels
    auto launch_line = [&](int t0, int t1, size_t grid_rays) {
      const int cblock = env_int("RJP_CHAN_BLOCK", GCH_MAX * LINE_THREADS);   // (experiments)
      for (int c0 = 0; c0 < nchan; c0 += cblock) {
        const int nc = (nchan - c0 < cblock) ? nchan - c0 : cblock;
        rjp_channels cb = *ch;
        cb.dnu += c0; cb.nu += c0; cb.cff += c0; cb.aff += c0; cb.bnu += c0;
        const size_t off = (size_t)c0 * plane;
        double* t_out = tau_rrl ? tau_rrl + off : nullptr;
        double* f_out = flux_rrl ? flux_rrl + off : nullptr;
        // equally spaced channels (the normal case) take the register-lean instantiation
        const bool uni = ln->chan_step != 0.0;
        // threads x channels-per-thread: few channels (a rank of a channel-sharded run) take a
        // one-warp CTA with 2 / 4 / 8 channels per thread
        int threads, gch = GCH_MAX;
        if (uni && nc <= 32 * GCH_MAX) {
          threads = 32;
          gch = nc <= 64 ? 2 : (nc <= 128 ? 4 : 8);
          if (GCH_MAX != 8) gch = GCH_MAX;
        } else {
          const int groups = (nc + GCH_MAX - 1) / GCH_MAX;
          threads = ((groups + 31) / 32) * 32;
        }
        const int force_t = env_int("RJP_LINE_THREADS", 0);
        if (force_t > 0 && uni && force_t * GCH_MAX >= nc) { threads = force_t; gch = GCH_MAX; }
#define RJP_LAUNCH_LINE(T, B, U, G)                                                           \
  integrate_line_kernel<T, B, U, G><<<(unsigned)grid_rays, threads, 0, ls>>>(                 \
      *ln, cb, nc, c0, contsub, dn_max, entries, meta, ray_list, n_active, t0, t1, t_out,     \
      f_out, plane, coff)
        if (threads <= 32 && uni) {
          if (gch == 2) RJP_LAUNCH_LINE(32, 16, true, 2);
          else if (gch == 4) RJP_LAUNCH_LINE(32, 16, true, 4);
          else RJP_LAUNCH_LINE(32, 16, true, 8);
        } else if (threads <= 64) {
          if (uni) RJP_LAUNCH_LINE(64, RJP_MINB64U, true, GCH_MAX);
          else RJP_LAUNCH_LINE(64, RJP_MINB64, false, GCH_MAX);
        } else if (threads <= 128) {
          if (uni) RJP_LAUNCH_LINE(128, RJP_MINB128, true, GCH_MAX);
          else RJP_LAUNCH_LINE(128, RJP_MINB128, false, GCH_MAX);
        } else {
          if (uni) RJP_LAUNCH_LINE(LINE_THREADS, 2, true, GCH_MAX);
          else RJP_LAUNCH_LINE(LINE_THREADS, 2, false, GCH_MAX);
        }
#undef RJP_LAUNCH_LINE
      }
    };
    // K4a for the rays [t0, t1): per-cell line constants, continuum sums, image pixels
    auto launch_prep = [&](int t0, int t1, size_t grid_rays, cudaStream_t st) {
      ray_prepare_kernel<<<(unsigned)grid_rays, RJP_PREP_THREADS, 0, st>>>(
          *m, *ep, *ct, ov, *ln, dn_max, c4, ex2, ray_list, n_active, t0, t1, cursor, max_cells,
          entries, meta, em, kff, tsum, tcount);
    };
    // (Cutting the list into chunks so that the preparation of chunk k+1 runs beside the
    // channel loop of chunk k was measured SLOWER: +0.05 ms per extra chunk at 67 048 rays x
    // 512 channels -- every chunk pays its own tail.)
    launch_prep(0, 2147483647, grid_all, ls);
    launch_line(0, 2147483647, grid_all);
  } else {
    size_t grid = (nray + 7) / 8;
    if (grid > (size_t)sms * 8) grid = (size_t)sms * 8;
    continuum_rays_kernel<<<(unsigned)grid, 256, 0, ls>>>(*m, *ep, *ct, ov, c4, ex2, ray_list,
                                                          n_active, em, kff, tsum, tcount);
  }
  if (fork) {
    if (cudaEventRecord(join, stream2) != cudaSuccess ||
        cudaStreamWaitEvent(stream, join, 0) != cudaSuccess)
      return RJP_ERR_CUDA;
  }
  return RJP_OK;
}

#ifdef RJP_EXPERIMENT
extern "C" int rjp_debug_stamps(unsigned long long* out4) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out4, g_stamp, sizeof(unsigned long long) * 4);
  unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull};
  cudaMemcpyToSymbol(g_stamp, init, sizeof(init));
  return 0;
}
#endif

extern "C" int rjp_launch_fill_missed(const int32_t* extents, long long nray, int nchan,
                                      long long plane, long long offset, long long skip_lo,
                                      long long skip_hi, double* tau, double* flux, int light,
                                      cudaStream_t stream) {
  if (nray <= 0 || nchan <= 0) return RJP_OK;
  set_carveouts();
  const int sms = device_info().sms;
  // light grid: meant to run on a side stream beside a long channel loop (multi-GPU: the
  // constants of the OTHER slabs' rays); otherwise enough warps to reach the HBM write peak
  const size_t ntiles = ((size_t)nray + CT_TILE - 1) / CT_TILE;
  const size_t items = ntiles * (size_t)((nchan + RJP_CT_CG - 1) / RJP_CT_CG);
  const int warps = 4;
  size_t grid = light ? (size_t)sms / 4 : (size_t)sms;
  if (grid * warps > items) grid = (items + warps - 1) / warps;
  ConstJob job = {reinterpret_cast<const int2*>(extents), (size_t)nray, nchan, nullptr, nullptr,
                  nullptr, nullptr, tau, flux, (size_t)plane, (size_t)offset, (size_t)skip_lo,
                  (size_t)skip_hi,
                  bulk_ok(tau, flux, (size_t)plane, (size_t)offset, (size_t)nray) ? 1 : 0};
  const_tiles_kernel<<<(unsigned)grid, 32 * warps, 0, stream>>>(job);
  return RJP_OK;
}

extern "C" int rjp_launch_column_totals(const double* cube, long long plane, long long offset,
                                        const int32_t* ray_list, const int32_t* n_active,
                                        int nchan, double* totals, cudaStream_t stream) {
  if (nchan <= 0) return RJP_OK;
  column_totals_kernel<<<nchan, 256, 0, stream>>>(cube, (size_t)plane, (size_t)offset, ray_list,
                                                  n_active, totals);
  return RJP_OK;
}

extern "C" int rjp_launch_pack_rays(const double* cube, long long plane, const int32_t* ray_ids,
                                    int n, int n_stride, int nchan, double* out,
                                    cudaStream_t stream) {
  if (n <= 0 || nchan <= 0) return RJP_OK;
  pack_rays_kernel<<<148 * 8, 256, 0, stream>>>(cube, (size_t)plane, ray_ids, n, n_stride, nchan,
                                                out);
  return RJP_OK;
}

extern "C" int rjp_launch_scatter_rays(const double* in, int n_stride, const int32_t* ray_ids,
                                       int n, int nchan, double* cube, long long plane,
                                       cudaStream_t stream) {
  if (n <= 0 || nchan <= 0) return RJP_OK;
  scatter_rays_kernel<<<148 * 8, 256, 0, stream>>>(in, n_stride, ray_ids, n, nchan, cube,
                                                   (size_t)plane);
  return RJP_OK;
}

extern "C" int rjp_launch_los_means(const rjp_model* m, const rjp_epoch* ep, const uint8_t* nverts,
                                    const int32_t* extents, const int32_t* ray_list,
                                    const int32_t* n_active, double* out, cudaStream_t stream) {
  const size_t nray = (size_t)(m->x_hi - m->x_lo) * m->nz;
  // NaN everywhere (all-ones bytes are a quiet NaN), then the jet-crossing rays
  cudaMemsetAsync(out, 0xFF, sizeof(double) * 7 * nray, stream);
  const size_t want = (nray + 7) / 8;
  los_means_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 256, 0, stream>>>(
      *m, *ep, nverts, reinterpret_cast<const int2*>(extents), ray_list, n_active, nray, out);
  return RJP_OK;
}

extern "C" int rjp_launch_voigt_profile(const double* x, const double* y, int64_t n,
                                        double* out, cudaStream_t stream) {
  if (n <= 0) return RJP_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  voigt_profile_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, y, n, out);
  return RJP_OK;
}

extern "C" int rjp_launch_continuum_images_epochs(const double*, int, const double*,
                                                  const int32_t*, int64_t, const double*,
                                                  const double*, double, int, double*, double*,
                                                  double*, cudaStream_t);

extern "C" int rjp_launch_continuum_images(const double* kff, const double* tsum,
                                           const int32_t* tcount, int64_t npix,
                                           const double* cff, const double* iff,
                                           double omega_jy, int nfreq, double* tau,
                                           double* inten, double* flux, cudaStream_t stream) {
  return rjp_launch_continuum_images_epochs(kff, 1, tsum, tcount, npix, cff, iff, omega_jy,
                                            nfreq, tau, inten, flux, stream);
}

extern "C" int rjp_launch_continuum_images_epochs(const double* kff, int n_epochs,
                                                  const double* tsum, const int32_t* tcount,
                                                  int64_t npix, const double* cff,
                                                  const double* iff, double omega_jy, int nfreq,
                                                  double* tau, double* inten, double* flux,
                                                  cudaStream_t stream) {
  if (npix <= 0 || nfreq <= 0 || n_epochs <= 0) return RJP_OK;
  if (n_epochs > 65535) return RJP_ERR_ARG;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const dim3 grid((unsigned)blocks, (unsigned)n_epochs);
  continuum_images_kernel<<<grid, 256, 0, stream>>>(kff, tsum, tcount, npix, cff, iff, omega_jy,
                                                    nfreq, tau, inten, flux);
  return RJP_OK;
}

// Continuum sums of a batch of model times (continuum_epochs_kernel).
extern "C" int rjp_launch_integrate_epochs(const rjp_model* m, const rjp_epoch* ep,
                                           const rjp_continuum* ct, const rjp_cell* cells,
                                           const int32_t* extents, const int32_t* ray_list,
                                           const int32_t* n_active, int n_hint,
                                           const double* times, int n_epochs, double* em,
                                           double* kff, double* tsum, int32_t* tcount,
                                           const double* travel_cells, cudaStream_t stream) {
  if (n_epochs <= 0) return RJP_OK;
  const int eblocks = (n_epochs + EPOCH_BLOCK - 1) / EPOCH_BLOCK;
  if (eblocks > 65535) return RJP_ERR_ARG;
  const CellGrids ov = {travel_cells, nullptr};
  const long long nray = (long long)(m->x_hi - m->x_lo) * m->nz;
  // one warp per listed ray and epoch block when the caller knows the count, else a grid
  // that fills the device a few times over (the warps stride over the list)
  long long ctas = n_hint > 0 ? ((long long)(n_hint < nray ? n_hint : nray) + 7) / 8
                              : (long long)device_info().sms * 8;
  if (ctas < 1) ctas = 1;
  const dim3 grid((unsigned)ctas, (unsigned)eblocks);
  continuum_epochs_kernel<<<grid, 256, 0, stream>>>(
      *m, *ep, *ct, ov, reinterpret_cast<const double2*>(cells),
      reinterpret_cast<const int2*>(extents), ray_list, n_active, times, n_epochs, em, kff, tsum,
      tcount);
  return RJP_OK;
}
