// Line-of-sight integration for sm_100a: K3 (continuum sums), K4 (LTE recombination-
// line opacity over all velocity channels) and K5 (image/cube epilogue) in ONE pass
// over the packed 16-byte cell state, so each cell is read from HBM exactly once.
//
// Mapping: a CTA owns ZT = 32 adjacent rays (fixed x, 32 consecutive z) over the whole
// line of sight (y).  Lanes run along z, so every warp-wide load is one contiguous
// 512-byte row of 16-byte cells; warps stride along y and keep RPW rows in flight each.
// Sums are accumulated in fp64 registers per (warp, ray) and reduced across the CTA's
// warps through shared memory at the end -- no atomics, deterministic order.
//
// The continuum part is HBM-bound (a handful of fp64 ops per in-jet cell, nothing for
// the empty ones).  The line part is fp64-pipe-bound: in-jet cells of a chunk of rows
// are compacted into a small shared-memory work list (channel-independent factors are
// computed once per cell), then the CTA switches to thread <-> channel and every
// thread accumulates its channel(s) over the list into a private row of the
// shared-memory tau_L[channel][ray] accumulator; the next chunk's rows are already in
// flight (register prefetch) while the Voigt profiles are evaluated.
#include "rjp_device.cuh"

namespace rjp {

constexpr int ZT = 32;          // rays per CTA
constexpr int RPW = 8;          // rows in flight per warp
constexpr int LCAP = 256;       // work-list capacity (entries)
constexpr int TAU_LD = ZT + 1;  // padded leading dimension of tau_s

struct LineEntry {   // channel-independent factors of one in-jet cell (rrls.py:329-389)
  double xs;         // -(nu0_cell - nu0) / (sigma sqrt2)
  double inv_s2;     // 1 / (sigma sqrt2)
  double y;          // (dnu_L / 2) / (sigma sqrt2)
  double amp;        // kappa0 n_e^2 T^-1.5 exp(Z^2 E_n / kT) ff / (sigma sqrt2)
  double p0;         // 1 - exp(-h nu0 / kT)
  double hk;         // h / (k T)
  int ray;
  int pad;
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
struct Params {
  rjp_model m;
  rjp_epoch ep;
  double tt_cst;   // MR0^q_v / (V0 (1 - q_v + eps q^d_v))          (geometry.py:154)
  double tt_f0;    // tt_cst * MR0^(1 - q_v): indefinite integral at r_0 when q^d_v = 0
};

__device__ __forceinline__ void stage_params(Params* s_p, const rjp_model& m,
                                             const rjp_epoch& ep) {
  const int nm = sizeof(rjp_model) / 4, ne = sizeof(rjp_epoch) / 4;
  const uint32_t* gm = reinterpret_cast<const uint32_t*>(&m);
  const uint32_t* ge = reinterpret_cast<const uint32_t*>(&ep);
  uint32_t* dm = reinterpret_cast<uint32_t*>(&s_p->m);
  uint32_t* de = reinterpret_cast<uint32_t*>(&s_p->ep);
  for (int i = threadIdx.x; i < nm; i += blockDim.x) dm[i] = gm[i];
  for (int i = threadIdx.x; i < ne; i += blockDim.x) de[i] = ge[i];
  if (threadIdx.x == 0) {
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
    if (m.qd_v == 0.0) {
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
__global__ void __launch_bounds__(256, 2)
integrate_continuum_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                           const double2* __restrict__ cells, double* __restrict__ em,
                           double* __restrict__ kff, double* __restrict__ tsum,
                           int32_t* __restrict__ tcount) {
  __shared__ double s_red[3 * 8 * ZT];
  __shared__ int s_cnt[8 * ZT];
  __shared__ Params s_p;
  stage_params(&s_p, m, ep);
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

// ------------------------------------------------------------------ continuum + line
__device__ __forceinline__ bool line_valid(const double2& c) {
  return c.x > 0.0 && fabs(c.y) > 0.0;
}

__device__ __noinline__ LineEntry make_entry(const Decoded& d, const rjp_model& m,
                                             const rjp_line& ln, int ix, int iy, int iz,
                                             int ray) {
  LineEntry e;
  const double vlos = velocity_of(m, centroid_rw(m, ix, iy, iz)).vlos_rel + m.v_lsr;
  const double shift = -ln.nu0 * (vlos * ln.dopp);           // nu0_cell - nu0 (physics.py:558)
  const double nu0c = ln.nu0 + shift;
  const double s2 = ln.width_g * sqrt(d.temp) * nu0c;        // sigma*sqrt2 (rrls.py:104-118, :349)
  e.inv_s2 = 1.0 / s2;
  e.xs = -shift * e.inv_s2;
  e.y = ln.stark * d.ne * e.inv_s2;                          // rrls.py:101, :353
  e.hk = ln.h_over_k / d.temp;
  e.p0 = -expm1(-e.hk * ln.nu0);
  // rrls.py:383-389 with n_i = (X mu'/m_amu) n_e, times path length and 1/(sigma sqrt(2 pi))
  e.amp = ln.kappa0 * d.ne * d.ne * d.ffw / (d.temp * sqrt(d.temp)) *
          exp(ln.en_over_k / d.temp) * e.inv_s2;
  if (!(vlos == vlos) || !d.ne_ok) {  // NaN velocity/density: nansum drops the cell
    e.amp = 0.0; e.xs = 0.0; e.inv_s2 = 0.0; e.y = 1.0;
  }
  e.ray = ray;
  e.pad = 0;
  return e;
}

__device__ __forceinline__ double line_term(const LineEntry& e, double dn) {
  const double x = fma(dn, e.inv_s2, e.xs);
  const double wr = faddeeva_re(x, e.y);
  const double d = e.hk * dn;
  double om;  // 1 - exp(-d)
  if (fabs(d) < 0.02)
    om = d * (1.0 + d * (-0.5 + d * (1.0 / 6.0 + d * (-1.0 / 24.0 + d * (1.0 / 120.0)))));
  else
    om = -expm1(-d);
  const double p4 = e.p0 + (1.0 - e.p0) * om;  // 1 - exp(-h nu / kT)  (rrls.py:387)
  return e.amp * wr * p4;
}

__global__ void __launch_bounds__(256)
integrate_line_kernel(const rjp_model m, const rjp_epoch ep, const rjp_continuum ct,
                      const rjp_line ln, const rjp_channels ch, const int nchan,
                      const int contsub, const double2* __restrict__ cells,
                      double* __restrict__ em, double* __restrict__ kff,
                      double* __restrict__ tsum, int32_t* __restrict__ tcount,
                      double* __restrict__ tau_rrl, double* __restrict__ flux_rrl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Params s_p;
  stage_params(&s_p, m, ep);
  const int nwarps = blockDim.x >> 5;
  double* tau_s = reinterpret_cast<double*>(smem_raw);                 // [nchan][TAU_LD]
  LineEntry* list = reinterpret_cast<LineEntry*>(tau_s + (size_t)nchan * TAU_LD);
  double* s_red = reinterpret_cast<double*>(list + LCAP);              // [3][nwarps][ZT]
  double* s_out = s_red + 3 * nwarps * ZT;                             // [3][ZT]
  int* s_cnt = reinterpret_cast<int*>(s_out + 3 * ZT);                 // [nwarps][ZT]
  int* s_scan = s_cnt + nwarps * ZT;                                   // [nwarps + 1]

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int ztiles = (m.nz + ZT - 1) / ZT;
  const int xl = blockIdx.x / ztiles;
  const int z0 = (blockIdx.x % ztiles) * ZT;
  const int iz = z0 + lane;
  const bool active = iz < m.nz;
  const double2* base = cells + (size_t)xl * m.ny * m.nz + (active ? iz : 0);
  const int nxs = m.x_hi - m.x_lo;
  const int ix = m.x_lo + xl;
  const Ray ray = ray_of(s_p.m, ix, iz);

  for (int i = tid; i < nchan * TAU_LD; i += blockDim.x) tau_s[i] = 0.0;

  ContAcc a = {0.0, 0.0, 0.0, 0};
  const int chunk = nwarps * RPW;
  double2 nxt[RPW];
#pragma unroll
  for (int j = 0; j < RPW; ++j) {
    const int y = wrp + j * nwarps;
    nxt[j] = (active && y < m.ny) ? ld_cell(base + (size_t)y * m.nz) : make_double2(0.0, 0.0);
  }
  __syncthreads();

  for (int yc = 0; yc < m.ny; yc += chunk) {
    double2 cur[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) cur[j] = nxt[j];
    // prefetch the next chunk: in flight while this chunk's profiles are evaluated
    if (yc + chunk < m.ny) {
#pragma unroll
      for (int j = 0; j < RPW; ++j) {
        const int y = yc + chunk + wrp + j * nwarps;
        nxt[j] = (active && y < m.ny) ? ld_cell(base + (size_t)y * m.nz)
                                      : make_double2(0.0, 0.0);
      }
    }
    int mine = 0;
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
      if (empty_cell(cur[j])) continue;
      accumulate(a, decode(cur[j], s_p, ray, ix, yc + wrp + j * nwarps, iz), ct.t_exponent);
      mine += line_valid(cur[j]) ? 1 : 0;
    }
    if (!__syncthreads_or(mine > 0)) continue;  // chunk has no line-emitting cell

    // block-wide exclusive scan of `mine`
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[wrp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
      const int v = s_scan[w];
      if (w < wrp) wbase += v;
      total += v;
    }
    const int my_first = wbase + incl - mine;

    for (int r0 = 0; r0 < total; r0 += LCAP) {
      int rank = my_first;
#pragma unroll
      for (int j = 0; j < RPW; ++j) {
        if (!line_valid(cur[j])) continue;
        if (rank >= r0 && rank < r0 + LCAP)
          list[rank - r0] = make_entry(decode(cur[j], s_p, ray, ix, yc + wrp + j * nwarps, iz),
                                       s_p.m, ln, ix, yc + wrp + j * nwarps, iz, lane);
        ++rank;
      }
      __syncthreads();
      const int nlist = min(LCAP, total - r0);
      for (int c = tid; c < nchan; c += blockDim.x) {
        const double dn = __ldg(ch.dnu + c);
        double* row = tau_s + (size_t)c * TAU_LD;
        int e = 0;
        for (; e + 1 < nlist; e += 2) {
          const LineEntry ea = list[e], eb = list[e + 1];
          const double va = line_term(ea, dn), vb = line_term(eb, dn);
          row[ea.ray] += va;
          row[eb.ray] += vb;
        }
        if (e < nlist) {
          const LineEntry ea = list[e];
          row[ea.ray] += line_term(ea, dn);
        }
      }
      __syncthreads();
    }
  }

  reduce_and_store(a, ct, s_red, s_cnt, nwarps, em, kff, tsum, tcount,
                   (size_t)xl * m.nz + iz, active, s_out);
  __syncthreads();

  // K5 epilogue: rrls.py:444-447, physics.py:571-574, classes.py:1323-1328, :1484-1488
  const double kray = s_out[0 * ZT + lane];
  const double cntr = s_out[2 * ZT + lane];
  const double tmean = s_out[1 * ZT + lane] / cntr;  // NaN for rays that miss the jet
  const size_t plane = (size_t)nxs * m.nz;
  const size_t pix = (size_t)xl * m.nz + iz;
  for (int c = wrp; c < nchan; c += nwarps) {
    const double tl = tau_s[(size_t)c * TAU_LD + lane];
    if (!active) continue;
    if (tau_rrl) tau_rrl[(size_t)c * plane + pix] = tl;
    if (flux_rrl) {
      double s = dnan();
      if (cntr > 0.0) {
        const double tc = __ldg(ch.cff + c) * kray;
        const double ec = exp(-tc);
        const double bnu = __ldg(ch.bnu + c) / (exp(ln.h_over_k * __ldg(ch.nu + c) / tmean) - 1.0);
        s = bnu * ec * (1.0 - exp(-tl));
        if (!contsub) s += __ldg(ch.aff + c) * (tmean * (1.0 - ec));
      }
      flux_rrl[(size_t)c * plane + pix] = s;
    }
  }
}

// ------------------------------------------------------------------ continuum images
__global__ void continuum_images_kernel(const double* __restrict__ kff,
                                        const double* __restrict__ tsum,
                                        const int32_t* __restrict__ tcount, int64_t npix,
                                        const double* __restrict__ cff,
                                        const double* __restrict__ iff, double omega_jy,
                                        int nfreq, double* __restrict__ tau,
                                        double* __restrict__ inten, double* __restrict__ flux) {
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

extern "C" size_t rjp_line_smem_bytes(int nchan, int nthreads) {
  const int nwarps = nthreads / 32;
  return (size_t)nchan * TAU_LD * 8 + (size_t)LCAP * sizeof(LineEntry) +
         (size_t)3 * nwarps * ZT * 8 + 3 * ZT * 8 + (size_t)nwarps * ZT * 4 +
         (size_t)(nwarps + 1) * 4 + 16;
}

extern "C" int rjp_launch_integrate(const rjp_model* m, const rjp_epoch* ep,
                                    const rjp_continuum* ct, const rjp_cell* cells,
                                    double* em, double* kff, double* tsum, int32_t* tcount,
                                    const rjp_line* ln, const rjp_channels* ch, int nchan,
                                    int contsub, double* tau_rrl, double* flux_rrl,
                                    cudaStream_t stream) {
  const int nxs = m->x_hi - m->x_lo;
  const long long ctas = (long long)nxs * ((m->nz + ZT - 1) / ZT);
  if (ctas <= 0 || ctas > 2147483647LL) return RJP_ERR_ARG;
  const double2* c4 = reinterpret_cast<const double2*>(cells);
  if (nchan <= 0 || ln == nullptr) {
    integrate_continuum_kernel<<<(unsigned)ctas, 256, 0, stream>>>(*m, *ep, *ct, c4, em, kff,
                                                                  tsum, tcount);
    return RJP_OK;
  }
  const int threads = 256;
  const size_t smem = rjp_line_smem_bytes(nchan, threads);
  if (smem > 227 * 1024) return RJP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(integrate_line_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return RJP_ERR_CUDA;
  integrate_line_kernel<<<(unsigned)ctas, threads, smem, stream>>>(
      *m, *ep, *ct, *ln, *ch, nchan, contsub, c4, em, kff, tsum, tcount, tau_rrl, flux_rrl);
  return RJP_OK;
}

extern "C" int rjp_launch_continuum_images(const double* kff, const double* tsum,
                                           const int32_t* tcount, int64_t npix,
                                           const double* cff, const double* iff,
                                           double omega_jy, int nfreq, double* tau,
                                           double* inten, double* flux, cudaStream_t stream) {
  if (npix <= 0 || nfreq <= 0) return RJP_OK;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  continuum_images_kernel<<<(unsigned)blocks, 256, 0, stream>>>(kff, tsum, tcount, npix, cff, iff,
                                                               omega_jy, nfreq, tau, inten, flux);
  return RJP_OK;
}
