// Device-side math of the RaJePy hot path for sm_100a.
//
// Everything here is fp64: the grid fill decides bit-exact integer vertex counts and
// feeds 1e-6-relative line-of-sight sums; B200's FP64 pipe (64 lanes/SM) is half the
// FP32 rate, and all heavy functions below run only for the few per cent of cells
// that lie inside the jet.  Reference citations are relative to the reference root.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "../../include/rajepy_b200.h"

namespace rjp {

__device__ __forceinline__ double dnan() { return CUDART_NAN; }

// ---------------------------------------------------------------- geometry
// maths/geometry.py:181-209 -> :212-260 ('yx') -> :266-302.  numpy never fuses a
// multiply with an add, so every product and sum is rounded separately (__dmul_rn /
// __dadd_rn are never contracted into FMAs).
struct Rw { double r, w, x1, y2; };

__device__ __forceinline__ Rw xyz_to_rw(const rjp_model& m, double x, double y, double z) {
  Rw o;
  o.x1 = __dadd_rn(__dmul_rn(m.cb, x), __dmul_rn(m.sb, z));
  double z1 = __dsub_rn(__dmul_rn(m.cb, z), __dmul_rn(m.sb, x));
  o.y2 = __dsub_rn(__dmul_rn(m.ca, y), __dmul_rn(m.sa, z1));
  o.r = __dadd_rn(__dmul_rn(m.sa, y), __dmul_rn(m.ca, z1));
  o.w = __dsqrt_rn(__dadd_rn(__dmul_rn(o.x1, o.x1), __dmul_rn(o.y2, o.y2)));
  return o;
}

// maths/geometry.py:34-61: ((|r| + mr0) - r0) / mr0, or |r| / r0 when mr0 is falsy.
__device__ __forceinline__ double rho_of(const rjp_model& m, double abs_r) {
  if (m.mr0 != 0.0) return __ddiv_rn(__dsub_rn(__dadd_rn(abs_r, m.mr0), m.r0), m.mr0);
  return __ddiv_rn(abs_r, m.r0);
}

// x ** q with numpy semantics (x ** 0 == 1 for every x, NaN for negative base and
// non-integer exponent); the q == 0 / q == 1 shortcuts only save time.
__device__ __forceinline__ double powq(double x, double q) {
  if (q == 0.0) return 1.0;
  if (q == 1.0) return x;
  return pow(x, q);
}

// Corner coordinate of lattice index i along an axis of n cells: cs * (i - n//2)
// (classes.py:497-499); exact integer -> one rounding.
__device__ __forceinline__ double corner(double cs, int i, int n) {
  return __dmul_rn(cs, (double)(i - n / 2));
}

// ---------------------------------------------------------------- vertex test
// classes.py:661-666: inside <=> w_0 * rho(r)**eps >= w  and  |r| >= r_0.
// Returns bit0 = decision in device arithmetic, bit1 = "too close to call": the
// decision could differ from numpy's (different pow implementation, or a vertex
// coordinate formed as cs*(i-h)+cs instead of cs*(i+1-h)); the host re-decides those
// with the reference's own numpy expression per (cell, corner).
__device__ __forceinline__ int vertex_inside(const rjp_model& m, double x, double y,
                                             double z) {
  Rw g = xyz_to_rw(m, x, y, z);
  const double ar = fabs(g.r);
  const double scale = fabs(x) + fabs(y) + fabs(z);
  const double tol = 3.5527136788005009e-15;  // 2^-48
  const bool in_r = ar >= m.r0;
  const bool r_band = fabs(ar - m.r0) <= tol * (scale + m.r0);
  if (!in_r && !r_band) return 0;
  const double rh = rho_of(m, ar);
  bool in_w, w_band = false;
  bool decided = false;
  if (rh > 0.0 && rh < 1e30 && g.w > 0.0) {
    // cheap certain decision from an fp32 estimate (relative error << 1e-3)
    float wa = (float)m.w0 * __powf((float)rh, (float)m.eps);
    float wf = (float)g.w;
    if (wa > wf * 1.001f) { in_w = true; decided = true; }
    else if (wa < wf * 0.999f) { in_w = false; decided = true; }
  }
  if (!decided) {
    double wr = __dmul_rn(m.w0, pow(rh, m.eps));
    in_w = wr >= g.w;
    w_band = fabs(wr - g.w) <= tol * (scale + g.w + fabs(wr));
  }
  int inside = (in_w && in_r) ? 1 : 0;
  int tie = ((w_band && (in_r || r_band)) || (r_band && (in_w || w_band))) ? 2 : 0;
  return inside | tie;
}

// ---------------------------------------------------------------- 2F1(a, b; b+1; z), z < 0
// scipy.special.hyp2f1 call of maths/geometry.py:168-171 (a = q^d_v,
// b = (1 - q_v + eps q^d_v)/eps).  Three regimes: Gauss series (|z| <= 1/2), Pfaff
// transformation z -> z/(z-1) in (1/3, 2/3] (|z| <= 2, or always when b - a is an
// integer), and the 1/z connection formula (A&S 15.3.7) whose second 2F1 collapses to
// 1 because c - b = 1.
__device__ inline double hyp2f1_bp1(const rjp_model& m, double z) {
  const double a = m.qd_v, b = m.hyp_b;
  if (a == 0.0 || z == 0.0) return 1.0;
  const double az = -z;
  if (!(az > 0.0)) return dnan();  // z > 0 needs R_2 < R_1: unsupported
  if (az <= 0.5) {
    double term = 1.0, sum = 1.0;
    for (int n = 0; n < 400; ++n) {
      term *= (a + n) / (n + 1.0) * z;
      double add = term * b / (b + n + 1.0);
      sum += add;
      if (fabs(add) <= 1e-17 * fabs(sum)) break;
    }
    return sum;
  }
  if (az <= 2.0 || m.hyp_degenerate) {
    const double zeta = az / (1.0 + az);
    double term = 1.0, sum = 1.0;
    for (int n = 0; n < 4000000; ++n) {
      term *= (a + n) / (b + 1.0 + n) * zeta;
      sum += term;
      if (fabs(term) <= 1e-17 * fabs(sum)) break;
    }
    return pow(1.0 + az, -a) * sum;
  }
  const double u = 1.0 / z;
  const double d = a - b;
  double term = 1.0, sum = 1.0;
  for (int n = 0; n < 400; ++n) {
    term *= (a + n) / (n + 1.0) * u;
    double add = term * d / (d + n + 1.0);
    sum += add;
    if (fabs(add) <= 1e-17 * fabs(sum)) break;
  }
  return m.hyp_c1 * pow(az, -a) * sum + m.hyp_c2 * pow(az, -b);
}

// ---------------------------------------------------------------- per-cell physics
// maths/geometry.py:150-173, SI units.  `reff_si` is only read when q^d_v != 0.
__device__ inline double travel_indef(const rjp_model& m, double r_si, double w_si) {
  const double W0 = m.w0 * m.au_m, R0 = m.r0 * m.au_m, MR0 = m.mr0 * m.au_m;
  const double V0 = m.v0 * 1e3;
  const double cst = powq(MR0, m.q_v) / (V0 * (1.0 - m.q_v + m.eps * m.qd_v));
  const double rad = r_si + MR0 - R0;
  const double p1 = powq(rad, 1.0 - m.q_v);
  if (m.qd_v == 0.0) return cst * p1;
  const double R1 = m.R1 * m.au_m, R2 = m.R2 * m.au_m;
  // r_eff with |r| in SI (geometry.py:157)
  const double rhos = (MR0 != 0.0) ? (fabs(r_si) + MR0 - R0) / MR0 : fabs(r_si) / R0;
  const double reff = R1 + ((R2 - R1) * w_si) / (W0 * pow(rhos, m.eps));
  const double p2 = pow(reff / R1, -m.qd_v);
  double p3, p4;
  if (w_si == 0.0) {
    p3 = 1.0;
    p4 = 1.0 + m.qd_v / (1.0 - m.q_v);
  } else {
    const double A = (R1 * W0 * pow(rad, m.eps)) / (w_si * pow(MR0, m.eps));
    p3 = pow(A / (R2 - R1) + 1.0, m.qd_v);
    p4 = hyp2f1_bp1(m, A / (R1 - R2));
  }
  return cst * p1 * p2 * p3 * p4;
}

__device__ __forceinline__ double clean(double v) {  // 0 -> NaN, +-inf -> NaN
  return (v == 0.0 || isinf(v)) ? dnan() : v;
}

// Centroid of cell (ix, iy, iz) in jet coordinates: corner + cs/2 (classes.py:521-523).
__device__ __forceinline__ Rw centroid_rw(const rjp_model& m, int ix, int iy, int iz) {
  const double h = m.cs / 2.0;
  return xyz_to_rw(m, __dadd_rn(corner(m.cs, ix, m.nx), h),
                   __dadd_rn(corner(m.cs, iy, m.ny), h),
                   __dadd_rn(corner(m.cs, iz, m.nz), h));
}

// |r| with the jet-base shift of classes.py:848-850 (= :884-886, :922-924, :1050-1052).
__device__ __forceinline__ double r_shifted(const rjp_model& m, double ar) {
  const double h = m.cs / 2.0;
  return (ar < m.r0 && (ar + h) >= m.r0) ? (m.r0 + ar + h) / 2.0 : ar;
}

// Travel time from the jet base to the cell [s]: classes.py:852 / geometry.py:177-178.
__device__ __forceinline__ double travel_time(const rjp_model& m, const Rw& g) {
  const double rt = r_shifted(m, fabs(g.r));
  const double F1 = travel_indef(m, rt * m.au_m, g.w * m.au_m);
  const double F0 = travel_indef(m, m.r0 * m.au_m, g.w * m.au_m);
  return (F1 - F0) / m.year_s * m.year_s;
}

// r_eff at the centroid: geometry.py:336 with |r| (classes.py:549-555).
__device__ __forceinline__ double reff_of(const rjp_model& m, const Rw& g, double rho_c_eps) {
  return m.R1 + ((m.R2 - m.R1) * g.w) / (m.w0 * rho_c_eps);
}

struct Velocity { double vx, vlos_rel, vz; };  // km/s, observer frame; vlos excludes v_lsr

// classes.py:1056-1093, physics.py:90, geometry.py:249-258 ('xy', 90 - inc, -pa).
__device__ inline Velocity velocity_of(const rjp_model& m, const Rw& g) {
  const double ar = fabs(g.r);
  const double rho_c = rho_of(m, ar);
  const double rce = pow(rho_c, m.eps);
  const double reff = reff_of(m, g, rce);
  const double rho_t = rho_of(m, r_shifted(m, ar));
  double vz = clean(m.v0 * powq(rho_t, m.q_v) * powq(reff / m.R1, m.qd_v));
  const double sgn = (g.r > 0.0) ? 1.0 : ((g.r < 0.0) ? -1.0 : 0.0);
  vz *= sgn;
  const double vrot = sqrt(m.gm_over_au / reff) * (1.0 / rce) / 1e3;
  const double q = g.y2 / g.w;  // sin(phi); NaN on the axis like arcsin(0/0)
  const double vxj = -vrot * q * m.rot_sign;
  const double vyj = vrot * (g.x1 / g.w) * m.rot_sign;
  const double y1 = m.cva * vyj - m.sva * vz;
  const double z1 = m.sva * vyj + m.cva * vz;
  Velocity v;
  v.vx = m.cvb * vxj + m.svb * z1;
  v.vlos_rel = y1;
  v.vz = m.cvb * z1 - m.svb * vxj;
  return v;
}

struct Laws { double nd, xi, temp, reff; };  // with the 0/inf -> NaN rule, not ff-masked

// classes.py:889-897 (n), :928-934 (x), :957-967 (T, incl. the cm-vs-au quirk).
__device__ inline Laws laws_of(const rjp_model& m, const Rw& g, bool force_reff) {
  Laws o;
  const double ar = fabs(g.r);
  double x_re = 1.0;
  o.reff = dnan();
  if (m.need_reff || force_reff) {
    o.reff = reff_of(m, g, pow(rho_of(m, ar), m.eps));
    x_re = o.reff / m.R1;
  }
  const double rho_t = rho_of(m, r_shifted(m, ar));
  double nd = clean(m.n0 * powq(rho_t, m.q_n) * powq(x_re, m.qd_n));
  if (g.r < 0.0) nd *= m.f_rb;
  o.nd = isinf(nd) ? dnan() : nd;
  o.xi = clean(m.x0 * powq(rho_t, m.q_x) * powq(x_re, m.qd_x));
  const double h = m.cs / 2.0;
  double rT = ar * m.au_cm;  // |r| in cm compared with r_0 in au (reference quirk, kept)
  if (rT < m.r0 && (rT + h) >= m.r0) rT = (m.r0 + rT + h) / 2.0;
  o.temp = clean(m.T0 * powq(rho_of(m, rT), m.q_T) * powq(x_re, m.qd_T));
  return o;
}

// Pack one in-jet cell (nverts > 0) into the 16-byte state.
__device__ inline rjp_cell pack_cell(const rjp_model& m, int ix, int iy, int iz, int nverts) {
  const Rw g = centroid_rw(m, ix, iy, iz);
  const Laws l = laws_of(m, g, false);
  rjp_cell c;
  const double ne0 = l.nd * l.xi;
  c.ne0 = (ne0 == ne0 && !isinf(ne0) && ne0 > 0.0) ? ne0 : 0.0;
  double t = (l.temp == l.temp && l.temp > 0.0) ? l.temp : 0.0;
  c.temp = (nverts < 8) ? -t : t;  // -0.0 keeps the flag when T is invalid
  return c;
}

// classes.py:442-448 divided by the steady-state rate: chi = 1 + sum amp_i exp(...)
__device__ __forceinline__ double burst_chi(const rjp_burst* b, int n, double t_launch) {
  double chi = 1.0;
  for (int i = 0; i < n; ++i) {
    const double d = t_launch - b[i].t0;
    chi += b[i].amp * exp(-(d * d) * b[i].inv2s2);
  }
  return chi;
}

// ---------------------------------------------------------------- Faddeeva (Voigt)
#include "rjp_weideman.inc"
__constant__ double c_weideman[RJP_WEIDEMAN_N] = {RJP_WEIDEMAN_COEFFS};

// Re w(x + iy), y > 0: Weideman's rational approximation; the degree-(N-1) polynomial
// with real coefficients is evaluated at the complex point Z with the real two-term
// recurrence b_k = a_k + 2Re(Z) b_{k+1} - |Z|^2 b_{k+2} (2 DFMA per coefficient).
// Replaces scipy.special.wofz at maths/rrls.py:353.
__device__ __forceinline__ double faddeeva_re(double x, double y) {
  const double L = RJP_WEIDEMAN_L;
  const double dr = L + y, nr = L - y;        // L - iz = dr - i x ; L + iz = nr + i x
  const double inv = 1.0 / (dr * dr + x * x);
  const double Zr = (nr * dr - x * x) * inv;
  const double Zi = (2.0 * L) * x * inv;
  const double s = 2.0 * Zr, q = Zr * Zr + Zi * Zi;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll
  for (int k = 0; k < RJP_WEIDEMAN_N - 1; ++k) {
    const double b0 = fma(s, b1, fma(-q, b2, c_weideman[k]));
    b2 = b1;
    b1 = b0;
  }
  const double pr = c_weideman[RJP_WEIDEMAN_N - 1] + Zr * b1 - q * b2;
  const double pi = Zi * b1;
  const double ur = dr * inv, ui = x * inv;   // 1/(L - iz)
  const double u2r = ur * ur - ui * ui, u2i = 2.0 * ur * ui;
  return 2.0 * (pr * u2r - pi * u2i) + ur * 0.56418958354775628695;  // 1/sqrt(pi)
}

// 1/d for finite positive d: MUFU.RCP64H seed (rcp.approx.ftz.f64, ~20 bits over the whole
// double range) + two Newton steps; within 1 ulp, no special cases needed here.
__device__ __forceinline__ double rcp_pos(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = r * fma(-d, r, 2.0);          // 2^-20 -> 2^-40
  r = fma(r, fma(-d, r, 1.0), r);   // -> rounding-limited
  return r;
}

// NV independent arguments (x_v, y_v): the NV recurrences are interleaved so that the
// fp64 pipe always has independent FMAs in flight (DFMA latency ~8.5 cycles on B200).
template <int NV>
__device__ __forceinline__ void faddeeva_re_n(const double* x, const double* y, double* out) {
  const double L = RJP_WEIDEMAN_L;
  double inv[NV], s[NV], q[NV], zr[NV], zi[NV], b1[NV], b2[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double dr = L + y[v], nr = L - y[v];
    inv[v] = rcp_pos(fma(x[v], x[v], dr * dr));
    zr[v] = (nr * dr - x[v] * x[v]) * inv[v];
    zi[v] = (2.0 * L) * x[v] * inv[v];
    s[v] = 2.0 * zr[v];
    q[v] = zr[v] * zr[v] + zi[v] * zi[v];
    b1[v] = 0.0;
    b2[v] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < RJP_WEIDEMAN_N - 1; ++k) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double b0 = fma(s[v], b1[v], fma(-q[v], b2[v], c_weideman[k]));
      b2[v] = b1[v];
      b1[v] = b0;
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double dr = L + y[v];
    const double pr = c_weideman[RJP_WEIDEMAN_N - 1] + zr[v] * b1[v] - q[v] * b2[v];
    const double pi = zi[v] * b1[v];
    const double ur = dr * inv[v], ui = x[v] * inv[v];
    const double u2r = ur * ur - ui * ui, u2i = 2.0 * ur * ui;
    out[v] = 2.0 * (pr * u2r - pi * u2i) + ur * 0.56418958354775628695;
  }
}

// Line wings, |z|^2 >= 36: six levels of the Laplace continued fraction
// w = (i/sqrt(pi)) / (z - (1/2)/(z - 1/(z - (3/2)/(z - ...)))) collapsed into
// (i/sqrt(pi)) p(u) / (z q(u)), u = z^2 (tools/gen_laplace_cf.py).  Re w to <= 2.1e-8 (|z| >= 6; 3e-10 for |z| >= 8)
// relative against wofz for |z| >= 8 at half the cost of the rational approximation above.
template <int NV>
__device__ __forceinline__ void faddeeva_wing_n(const double* x, const double* y, double* out) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double ur = fma(x[v], x[v], -y[v] * y[v]), ui = 2.0 * x[v] * y[v];
    // p(u) = ((u - 10) u + 21.75) u - 6
    double ar = ur - 10.0, ai = ui, t;
    t = fma(ar, ur, fma(-ai, ui, 21.75)); ai = fma(ar, ui, ai * ur); ar = t;
    t = fma(ar, ur, fma(-ai, ui, -6.0)); ai = fma(ar, ui, ai * ur); ar = t;
    // q(u) = ((u - 10.5) u + 26.25) u - 13.125
    double qr = ur - 10.5, qi = ui;
    t = fma(qr, ur, fma(-qi, ui, 26.25)); qi = fma(qr, ui, qi * ur); qr = t;
    t = fma(qr, ur, fma(-qi, ui, -13.125)); qi = fma(qr, ui, qi * ur); qr = t;
    // B = z q(u);  Re w = (A_r B_i - A_i B_r) / (sqrt(pi) |B|^2)
    const double br = fma(x[v], qr, -y[v] * qi), bi = fma(x[v], qi, y[v] * qr);
    const double num = fma(ar, bi, -ai * br);
    const double den = fma(br, br, bi * bi);
    out[v] = 0.56418958354775628695 * num * rcp_pos(den);
  }
}


// ---------------------------------------------------------------- mixed-precision Voigt
// Cells with a small Lorentz/Gauss ratio, RJP_VT_Y_MIN <= y <= RJP_VT_Y_MAX (97 % of the in-jet
// cells of the BASELINE jets), take the split K = Re w = 2^(Y^2 - X^2) cos(2xy) - H(x, y)
// derived in tools/gen_voigt_tables.py: the ill-conditioned Gaussian in fp64, the smooth
// part H (proportional to y) in fp32, products accumulated in fp64.  Measured against
// scipy.special.wofz (tests/test_faddeeva.py, same arithmetic emulated in numpy): <= 2.2e-7
// relative per evaluation, 5e-8 rms, against the 1e-6 bar on the line-of-sight sums.
// The fp64 <-> fp32 moves are exponent re-biases on the integer pipe: F2F conversions run
// at 16/clk/SM on B200 (profiles/r1_microbench_mix.txt), a quarter of the DFMA rate.
#include "rjp_voigt_tables.inc"
__device__ const float g_vt_core[RJP_VT_NI * RJP_VT_ROW] = {RJP_VT_CORE};
constexpr int VT_TAB_F4 = RJP_VT_NI * RJP_VT_ROW / 4;

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {  // 32-bit shared-window address
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

struct FastEntry {     // one in-jet cell of the fast class, X = kappa x (80 bytes)
  double xs;           // X at nu = nu0
  double inv;          // dX / dnu = kappa / (sigma sqrt2)
  double yy;           // Y^2 = (kappa y)^2
  double w0;           // wing lead factor: amp p0 y kappa^2 G1(0)
  double a0;           // core lead factor: amp p0
  float c1, c2, c3, c4;  // y-dependent coefficients of the wing polynomial P(U)
  float yf, y2f;       // y, y^2
  float ya;            // 2 y / kappa: cos argument per unit X
  float b1, b2;        // (1 - exp(-h nu / kT)) / p0 = 1 + dn (b1 + dn b2)
  int xc2_hi;          // high word of the X^2 below which the Gaussian matters
};
static_assert(sizeof(FastEntry) == 80, "FastEntry layout");

// (double)v of a positive normal float / its truncating inverse, without F2F
__device__ __forceinline__ double f2d_pos(float v) {
  const unsigned b = __float_as_uint(v);
  return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
__device__ __forceinline__ float d2f_trunc_pos(double v) {
  const unsigned hi = (unsigned)__double2hiint(v) - 0x38000000u;
  return __uint_as_float(__funnelshift_l((unsigned)__double2loint(v), hi, 3));
}
__device__ __forceinline__ double rcp_seed(double d) {  // MUFU.RCP64H, >= 20 bits
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  return r;
}

// y-dependent constants of a fast-class cell (once per cell, phase 1 of the channel loop)
__device__ inline void vt_cell_constants(double y, FastEntry& e) {
  const double K = RJP_VT_KAPPA;
  constexpr float g1[] = {RJP_VT_G1}, g3[] = {RJP_VT_G3}, g5[] = {RJP_VT_G5};
  e.yy = (K * y) * (K * y);
  const float yyf = (float)e.yy;
  e.c1 = fmaf(yyf, g3[0], g1[1]);
  e.c2 = fmaf(yyf, fmaf(yyf, g5[0], g3[1]), g1[2]);
  e.c3 = fmaf(yyf, fmaf(yyf, g5[1], g3[2]), g1[3]);
  e.c4 = fmaf(yyf, g3[3], g1[4]);
  e.yf = (float)y;
  e.y2f = (float)(y * y);
  e.ya = (float)(2.0 * y / K);
  // the Gaussian is dropped where exp(-x^2) < 1e-8 K ~ 1e-8 y / (sqrt(pi) x^2), but never
  // below the validity limit of the wing polynomials nor beyond the core table
  // (the switch point only has to be deterministic and on the safe side: fp32 logs)
  const float lcf = __logf(1e8f * 1.7724538509055159f / (float)y) + 1e-3f;
  float x2f = lcf + 3.3322045f;
  x2f = lcf + __logf(x2f);
  x2f = lcf + __logf(x2f);
  const double x2 = (double)x2f;
  e.xc2_hi = __double2hiint(fmin(fmax(x2 * K * K, RJP_VT_XWING2), RJP_VT_XCORE2 - 0.01));
}

// Packed fp32 pairs (sm_100 FFMA2: two FMAs per issue slot -- the channel loop is issue-bound)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// P(U) for two channels at once (same cell): same arithmetic as vt_wing_poly per half
struct WingCoef2 { f32x2 g6, g5, c4, c3, c2, c1, one; };
__device__ __forceinline__ WingCoef2 vt_wing_coef2(const FastEntry& e) {
  constexpr float g1[] = {RJP_VT_G1};
  WingCoef2 w;
  w.g6 = pk2(g1[6], g1[6]);
  w.g5 = pk2(g1[5], g1[5]);
  w.c4 = pk2(e.c4, e.c4);
  w.c3 = pk2(e.c3, e.c3);
  w.c2 = pk2(e.c2, e.c2);
  w.c1 = pk2(e.c1, e.c1);
  w.one = pk2(1.0f, 1.0f);
  return w;
}
__device__ __forceinline__ f32x2 vt_wing_poly2(const WingCoef2& w, f32x2 u) {
  f32x2 p = fma2(w.g6, u, w.g5);
  p = fma2(p, u, w.c4);
  p = fma2(p, u, w.c3);
  p = fma2(p, u, w.c2);
  p = fma2(p, u, w.c1);
  return fma2(p, u, w.one);
}

// wings: K = y kappa^2 G1(0) U P(U), U = 1 / X^2; this is P (fp32), U stays in fp64
__device__ __forceinline__ float vt_wing_poly(const FastEntry& e, float u) {
  constexpr float g1[] = {RJP_VT_G1};
  float p = fmaf(g1[6], u, g1[5]);
  p = fmaf(p, u, e.c4);
  p = fmaf(p, u, e.c3);
  p = fmaf(p, u, e.c2);
  p = fmaf(p, u, e.c1);
  return fmaf(p, u, 1.0f);
}

// core: K (fp32) for X^2 < 64; `tab` = shared-window address of g_vt_core staged in
// shared memory
__device__ __forceinline__ float vt_core(const FastEntry& e, uint32_t tab, double X,
                                         double X2) {
  // |X| 2^25 as an integer from the low mantissa word of |X| + 1.5 2^27
  const int q = __double2loint(fabs(X) + 201326592.0);
  const uint32_t row = tab + (uint32_t)(q >> 24) * (RJP_VT_ROW * 4);
  const float t = fmaf(__int2float_rn(q & 0xFFFFFF), 1.1920928955078125e-07f, -1.0f);
  const float4 r0 = lds_f4(row), r1 = lds_f4(row + 16), r2 = lds_f4(row + 32),
               r3 = lds_f4(row + 48);
  float a = fmaf(r1.w, t, r1.z);
  a = fmaf(a, t, r1.y);
  a = fmaf(a, t, r1.x);
  a = fmaf(a, t, r0.w);
  a = fmaf(a, t, r0.z);
  a = fmaf(a, t, r0.y);
  a = fmaf(a, t, r0.x);
  float b = fmaf(r3.y, t, r3.x);
  b = fmaf(b, t, r2.w);
  b = fmaf(b, t, r2.z);
  b = fmaf(b, t, r2.y);
  b = fmaf(b, t, r2.x);
  const float4 r4 = lds_f4(row + 64);
  float c = fmaf(r4.y, t, r4.x);
  c = fmaf(c, t, r3.w);
  c = fmaf(c, t, r3.z);
  const float hy = fmaf(e.y2f, fmaf(e.y2f, c, b), a);          // H / y
  // Gaussian 2^(Y^2 - X^2): the argument in fp64, split into integer and fraction by the
  // 1.5 2^52 shift
  const double T = e.yy - X2;
  const double M = T + 6755399441055744.0;
  const double f = T - (M - 6755399441055744.0);
  // 2^f on the SFU (MUFU.EX2: <= 2 ulp of fp32, measured 3.6e-8 rms on the whole profile,
  // tools/voigt_probe.py); the integer part goes into the exponent field.  A degree-8 fp64
  // polynomial here cost 9 more issue slots per evaluation (4.89 -> 5.13 ms per pass).
  float gf;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(gf) : "f"((float)f));
  const float g = __uint_as_float(__float_as_uint(gf) + ((unsigned)__double2loint(M) << 23));
  // cos(2xy) - 1, 2xy = X * ya <= 1.4
  const float w = (__int2float_rn(q) * 2.98023223876953125e-08f) * e.ya;
  const float w2 = w * w;
  float cs = fmaf(-2.755731922398589e-07f, w2, 2.48015873015873e-05f);
  cs = fmaf(cs, w2, -1.388888888888889e-03f);
  cs = fmaf(cs, w2, 4.1666666666666664e-02f);
  cs = fmaf(cs, w2, -0.5f);
  return fmaf(-e.yf, hy, fmaf(g, w2 * cs, g));
}

// Re w(x + iy) with the routines of the channel loop (diagnostic entry rjp_voigt_profile)
__device__ inline double voigt_any(uint32_t tab, double x, double y) {
  if (y >= RJP_VT_Y_MIN && y <= RJP_VT_Y_MAX) {
    FastEntry e;
    vt_cell_constants(y, e);
    const double X = RJP_VT_KAPPA * x, X2 = X * X;
    if (__double2hiint(X2) < e.xc2_hi) return f2d_pos(vt_core(e, tab, X, X2));
    const double r0 = rcp_seed(X2), U = r0 * fma(-X2, r0, 2.0);
    return (y * RJP_VT_KAPPA * RJP_VT_KAPPA * RJP_VT_G10) * U *
           f2d_pos(vt_wing_poly(e, d2f_trunc_pos(r0)));
  }
  double xv[1] = {x}, yv[1] = {y}, w[1];
  if (x * x + y * y >= 36.0) faddeeva_wing_n<1>(xv, yv, w);
  else faddeeva_re_n<1>(xv, yv, w);
  return w[0];
}

}  // namespace rjp
