// Grid fill for sm_100a: K1 (8-vertex inside test, classes.py:657-669) fused with K2
// (per-cell state, classes.py:838-1095).
//
// The grid is cut into bricks of TX x TY x TZ cells.  A brick that provably lies outside
// the jet (conservative bound on its circumscribed sphere) is all zeros; every other brick
// tests its (TX+1)(TY+1)(TZ+1) lattice vertices ONCE each (instead of 8 times, as the
// reference's eight grid passes do), stages the inside bits in shared memory, and every cell
// sums its eight corner bits.  Lanes run along z, the contiguous axis, so the 16-byte state
// stores and the 1-byte count stores of a warp are fully coalesced (512 B / 32 B).  No input
// is read from HBM.  One CTA owns a super-brick of SBX x SBY bricks (32^3 cells): its first
// warp decides all 32 bricks at once, then the CTA walks the bricks that need work.
//
// `brick_state` (optional, one byte per brick) makes the fill sparse: it records which bricks
// of the caller's nverts / cells buffers hold data.  A rejected brick whose state is 0 is
// already zero and is not written at all, so with a zero-initialised (or previously used)
// buffer the fill writes only the ~3 % of bricks around the jet instead of 17 B for every
// cell of the grid.  Without it every brick is written (dense behaviour).
#include "rjp_device.cuh"

namespace rjp {

constexpr int TX = 4, TY = 8, TZ = 32;
#ifndef RJP_SBX
#define RJP_SBX 8
#endif
#ifndef RJP_SBY
#define RJP_SBY 4
#endif
constexpr int SBX = RJP_SBX, SBY = RJP_SBY;  // bricks per super-brick along x and y
constexpr int FILL_THREADS = 256;
constexpr int NVERT = (TX + 1) * (TY + 1) * (TZ + 1);

// Conservative whole-brick rejection (the jet fills < 1-15 % of the grid).  Every vertex v of
// the brick lies within R_b of the brick centre c, so w_v >= w_c - R_b and
// |r_v| <= |r_c| + R_b; the jet width w_0 rho(|r|)^eps grows with |r| (eps > 0), hence no
// vertex can pass the inside test if w_c - R_b > w_jet(|r_c| + R_b) (or if the whole brick is
// below the launch radius).  The 1e-6 margin dwarfs rounding, so the integer counts are
// unchanged: such bricks are all zeros.
__device__ inline bool brick_outside(const rjp_model& m, int tx0, int ty0, int tz0) {
  if (!(m.eps > 0.0 && m.w0 > 0.0)) return false;
  const double hx = 0.5 * TX * m.cs, hy = 0.5 * TY * m.cs, hz = 0.5 * TZ * m.cs;
  const Rw c = xyz_to_rw(m, corner(m.cs, tx0, m.nx) + hx, corner(m.cs, ty0, m.ny) + hy,
                         corner(m.cs, tz0, m.nz) + hz);
  const double rb = sqrt(hx * hx + hy * hy + hz * hz) * (1.0 + 1e-9);
  const double rmax = fabs(c.r) + rb;
  if (rmax < m.r0 * (1.0 - 1e-9)) return true;
  const double rh = rho_of(m, rmax);
  return rh > 0.0 && (c.w - rb) > m.w0 * pow(rh, m.eps) * (1.0 + 1e-6);
}

// All-zero brick (outside the jet) into a buffer that holds other data there.
__device__ __forceinline__ void brick_zero(const rjp_model& m, int tx0, int ty0, int tz0,
                                           uint8_t* __restrict__ nverts,
                                           rjp_cell* __restrict__ cells) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int lx = 0; lx < TX; ++lx) {
    const int ix = tx0 + lx, iy = ty0 + wrp, iz = tz0 + lane;
    if (ix >= m.x_hi || iy >= m.ny || iz >= m.nz) continue;
    const size_t idx = ((size_t)(ix - m.x_lo) * m.ny + iy) * m.nz + iz;
    nverts[idx] = 0;
    reinterpret_cast<double2*>(cells)[idx] = make_double2(0.0, 0.0);
  }
}

// One brick that may intersect the jet: vertex tests into s_in, then the cells.  Called by
// all FILL_THREADS threads of the CTA; ends with a barrier (s_in can be reused).
__device__ __forceinline__ void brick_compute(const rjp_model& m, int tx0, int ty0, int tz0,
                                              uint8_t* s_in, uint8_t* __restrict__ nverts,
                                              rjp_cell* __restrict__ cells,
                                              int32_t* __restrict__ ties, int32_t tie_capacity,
                                              int32_t* __restrict__ n_ties,
                                              int32_t* __restrict__ extents) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int v = threadIdx.x; v < NVERT; v += FILL_THREADS) {
    const int lz = v % (TZ + 1);
    const int ly = (v / (TZ + 1)) % (TY + 1);
    const int lx = v / ((TZ + 1) * (TY + 1));
    const int I = tx0 + lx, J = ty0 + ly, K = tz0 + lz;
    int res = 0;
    if (I <= m.x_hi && J <= m.ny && K <= m.nz) {
      res = vertex_inside(m, corner(m.cs, I, m.nx), corner(m.cs, J, m.ny),
                          corner(m.cs, K, m.nz));
      // report each near-tie once: by the brick that owns the vertex (lower faces),
      // or by the last brick at the slab / grid upper faces
      const bool own = (lx < TX || I == m.x_hi) && (ly < TY || J == m.ny) &&
                       (lz < TZ || K == m.nz);
      if ((res & 2) && own) {
        const int slot = atomicAdd(n_ties, 1);
        if (slot < tie_capacity) {
          ties[4 * slot + 0] = I;
          ties[4 * slot + 1] = J;
          ties[4 * slot + 2] = K;
          ties[4 * slot + 3] = res & 1;
        }
      }
    }
    s_in[v] = (uint8_t)(res & 1);
  }
  __syncthreads();

#pragma unroll
  for (int lx = 0; lx < TX; ++lx) {
    const int ix = tx0 + lx, iy = ty0 + wrp, iz = tz0 + lane;  // warp index == local y
    if (ix >= m.x_hi || iy >= m.ny || iz >= m.nz) continue;
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int vx = lx + (c & 1), vy = wrp + ((c >> 1) & 1), vz = lane + (c >> 2);
      cnt += s_in[(vx * (TY + 1) + vy) * (TZ + 1) + vz];
    }
    const size_t idx = ((size_t)(ix - m.x_lo) * m.ny + iy) * m.nz + iz;
    nverts[idx] = (uint8_t)cnt;
    rjp_cell c = {0.0, 0.0};
    if (cnt > 0) {
      c = pack_cell(m, ix, iy, iz, cnt);
      // y-extent of the ray's in-jet cells, for the ray kernels of the integration pass
      int32_t* e = extents + 2 * ((size_t)(ix - m.x_lo) * m.nz + iz);
      atomicMin(e, iy);
      atomicMax(e + 1, iy + 1);
    }
    reinterpret_cast<double2*>(cells)[idx] = make_double2(c.ne0, c.temp);
  }
  __syncthreads();   // s_in is reused by the next brick
}

__global__ void __launch_bounds__(FILL_THREADS)
fill_grid_kernel(const rjp_model m, uint8_t* __restrict__ nverts,
                 rjp_cell* __restrict__ cells, uint8_t* __restrict__ brick_state,
                 int32_t* __restrict__ ties, int32_t tie_capacity,
                 int32_t* __restrict__ n_ties, int32_t* __restrict__ extents) {
  __shared__ uint8_t s_in[NVERT];
  __shared__ int s_todo[SBX * SBY];   // 0 nothing, 1 write zeros, 2 compute
  const int nxs = m.x_hi - m.x_lo;
  const int tiles_x = (nxs + TX - 1) / TX;
  const int tiles_y = (m.ny + TY - 1) / TY;
  const int tiles_z = (m.nz + TZ - 1) / TZ;
  const int sup_y = (tiles_y + SBY - 1) / SBY;
  int t = blockIdx.x;
  const int bz = t % tiles_z; t /= tiles_z;
  const int sy = t % sup_y;
  const int sx = t / sup_y;

  if (threadIdx.x < SBX * SBY) {
    const int bx = sx * SBX + threadIdx.x / SBY, by = sy * SBY + threadIdx.x % SBY;
    int todo = 0;
    if (bx < tiles_x && by < tiles_y) {
      const bool out = brick_outside(m, m.x_lo + bx * TX, by * TY, bz * TZ);
      const size_t bid = ((size_t)bx * tiles_y + by) * tiles_z + bz;
      const bool holds_data = brick_state ? brick_state[bid] != 0 : true;
      todo = out ? (holds_data ? 1 : 0) : 2;
      if (brick_state) brick_state[bid] = out ? 0 : 1;
    }
    s_todo[threadIdx.x] = todo;
  }
  __syncthreads();

  for (int k = 0; k < SBX * SBY; ++k) {
    const int todo = s_todo[k];            // uniform over the CTA
    if (todo == 0) continue;
    const int tx0 = m.x_lo + (sx * SBX + k / SBY) * TX;
    const int ty0 = (sy * SBY + k % SBY) * TY;
    const int tz0 = bz * TZ;
    if (todo == 1) brick_zero(m, tx0, ty0, tz0, nverts, cells);
    else brick_compute(m, tx0, ty0, tz0, s_in, nverts, cells, ties, tie_capacity, n_ties, extents);
  }
}

// Two-level sparse fill (needs the occupancy map and a work list from the caller): one thread
// per brick classifies it and appends the bricks that need work to a list; a persistent grid
// then pulls bricks from the list one at a time, so the ~3 % of bricks around the jet are
// spread evenly over the SMs (the one-kernel fill above leaves them to ~3 % of its CTAs).
// work[0] = list length, work[1] = cursor, work[4 + k] = 2 * brick id + (1 if compute).
__global__ void __launch_bounds__(256)
fill_classify_kernel(const rjp_model m, uint8_t* __restrict__ brick_state,
                     int32_t* __restrict__ work, long long nbricks) {
  const long long bid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int tiles_y = (m.ny + TY - 1) / TY;
  const int tiles_z = (m.nz + TZ - 1) / TZ;
  int todo = 0;
  if (bid < nbricks) {
    const int bz = (int)(bid % tiles_z);
    const int by = (int)((bid / tiles_z) % tiles_y);
    const int bx = (int)(bid / ((long long)tiles_z * tiles_y));
    const bool out = brick_outside(m, m.x_lo + bx * TX, by * TY, bz * TZ);
    const bool holds_data = brick_state[bid] != 0;
    todo = out ? (holds_data ? 1 : 0) : 2;
    brick_state[bid] = out ? 0 : 1;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, todo != 0);
  if (bal == 0u) return;
  const int lane = threadIdx.x & 31, leader = __ffs(bal) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(work, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (todo != 0)
    work[4 + base + __popc(bal & ((1u << lane) - 1u))] = (int32_t)(bid * 2 + (todo == 2 ? 1 : 0));
}

__global__ void __launch_bounds__(FILL_THREADS)
fill_bricks_kernel(const rjp_model m, uint8_t* __restrict__ nverts, rjp_cell* __restrict__ cells,
                   int32_t* __restrict__ ties, int32_t tie_capacity,
                   int32_t* __restrict__ n_ties, int32_t* __restrict__ extents,
                   int32_t* __restrict__ work) {
  __shared__ uint8_t s_in[NVERT];
  __shared__ int s_next;
  const int tiles_y = (m.ny + TY - 1) / TY;
  const int tiles_z = (m.nz + TZ - 1) / TZ;
  const int n = work[0];
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(work + 1, 1);
    __syncthreads();
    const int k = s_next;
    __syncthreads();
    if (k >= n) break;
    const int entry = work[4 + k];
    const long long bid = entry >> 1;
    const int tz0 = (int)(bid % tiles_z) * TZ;
    const int ty0 = (int)((bid / tiles_z) % tiles_y) * TY;
    const int tx0 = m.x_lo + (int)(bid / ((long long)tiles_z * tiles_y)) * TX;
    if (entry & 1)
      brick_compute(m, tx0, ty0, tz0, s_in, nverts, cells, ties, tie_capacity, n_ties, extents);
    else
      brick_zero(m, tx0, ty0, tz0, nverts, cells);
  }
}

__global__ void init_extents_kernel(int2* __restrict__ extents, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) extents[i] = make_int2(2147483647, 0);
}

__global__ void patch_cells_kernel(const rjp_model m, const int64_t* __restrict__ cell_idx,
                                   const uint8_t* __restrict__ new_count, int32_t n,
                                   uint8_t* __restrict__ nverts,
                                   rjp_cell* __restrict__ cells,
                                   uint8_t* __restrict__ brick_state,
                                   int32_t* __restrict__ extents) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t idx = cell_idx[i];
  const int iz = (int)(idx % m.nz);
  const int iy = (int)((idx / m.nz) % m.ny);
  const int ix = m.x_lo + (int)(idx / ((int64_t)m.nz * m.ny));
  const int cnt = new_count[i];
  nverts[idx] = (uint8_t)cnt;
  rjp_cell c = {0.0, 0.0};
  if (cnt > 0) {
    c = pack_cell(m, ix, iy, iz, cnt);
    int32_t* e = extents + 2 * ((size_t)(ix - m.x_lo) * m.nz + iz);
    atomicMin(e, iy);
    atomicMax(e + 1, iy + 1);
    if (brick_state) {
      const int tiles_y = (m.ny + TY - 1) / TY, tiles_z = (m.nz + TZ - 1) / TZ;
      brick_state[((size_t)((ix - m.x_lo) / TX) * tiles_y + iy / TY) * tiles_z + iz / TZ] = 1;
    }
  }
  cells[idx] = c;
}

// User-assigned temperature / ionisation-fraction grids (the setters of classes.py:936-940,
// :994-1000): re-pack the state of the in-jet cells from the caller's 3-D array.
__global__ void __launch_bounds__(256)
override_cells_kernel(const rjp_model m, const uint8_t* __restrict__ nverts, int field,
                      const double* __restrict__ values, rjp_cell* __restrict__ cells) {
  const size_t ncell = (size_t)(m.x_hi - m.x_lo) * m.ny * m.nz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ncell;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int cnt = nverts[idx];
    if (cnt == 0) continue;
    const double v = values[idx];
    rjp_cell c = cells[idx];
    if (field == RJP_FIELD_TEMP) {
      const double t = (v == v && v > 0.0 && !isinf(v)) ? v : 0.0;
      c.temp = (cnt < 8) ? -t : t;
    } else {
      const int iz = (int)(idx % m.nz);
      const int iy = (int)((idx / m.nz) % m.ny);
      const int ix = m.x_lo + (int)(idx / ((size_t)m.nz * m.ny));
      const double ne0 = laws_of(m, centroid_rw(m, ix, iy, iz), false).nd * v;
      c.ne0 = (ne0 == ne0 && !isinf(ne0) && ne0 > 0.0) ? ne0 : 0.0;
    }
    cells[idx] = c;
  }
}

__global__ void __launch_bounds__(256)
cell_field_kernel(const rjp_model m, const rjp_epoch ep, const uint8_t* __restrict__ nverts,
                  int field, double* __restrict__ out) {
  const size_t ncell = (size_t)(m.x_hi - m.x_lo) * m.ny * m.nz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < ncell;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int cnt = nverts[idx];
    double v;
    if (field == RJP_FIELD_FILL_FACTOR) {
      v = (cnt == 8) ? 1.0 : (cnt > 0 ? 0.5 : dnan());
    } else if (field == RJP_FIELD_AREAS) {
      v = cnt > 0 ? 1.0 : dnan();
    } else {
      const bool masked = field >= RJP_FIELD_ND_BASE && field <= RJP_FIELD_VZ;
      if (masked && cnt == 0) {
        v = dnan();
      } else {
        const int iz = (int)(idx % m.nz);
        const int iy = (int)((idx / m.nz) % m.ny);
        const int ix = m.x_lo + (int)(idx / ((size_t)m.nz * m.ny));
        const Rw g = centroid_rw(m, ix, iy, iz);
        switch (field) {
          case RJP_FIELD_R: v = g.r; break;
          case RJP_FIELD_W: v = g.w; break;
          case RJP_FIELD_PHI: {  // geometry.py:291-294
            const double p = asin(g.y2 / g.w);
            v = (g.x1 < 0.0) ? (CUDART_PI - p) : p;
            break;
          }
          case RJP_FIELD_REFF: v = laws_of(m, g, true).reff; break;
          case RJP_FIELD_TRAVEL: v = travel_time(m, g); break;
          case RJP_FIELD_ND_BASE: v = laws_of(m, g, false).nd; break;
          case RJP_FIELD_XI: v = laws_of(m, g, false).xi; break;
          case RJP_FIELD_TEMP: v = laws_of(m, g, false).temp; break;
          case RJP_FIELD_VX: v = velocity_of(m, g).vx; break;
          case RJP_FIELD_VLOS: v = velocity_of(m, g).vlos_rel + m.v_lsr; break;
          case RJP_FIELD_VZ: v = velocity_of(m, g).vz; break;
          case RJP_FIELD_CHI: {
            const double tl = ep.time - travel_time(m, g);
            v = (g.r < 0.0) ? burst_chi(ep.red, ep.n_red, tl)
                            : burst_chi(ep.blue, ep.n_blue, tl);
            break;
          }
          default: v = dnan();
        }
      }
    }
    out[idx] = v;
  }
}

}  // namespace rjp

using namespace rjp;

extern "C" long long rjp_launch_brick_count(const rjp_model* m) {
  const int nxs = m->x_hi - m->x_lo;
  return (long long)((nxs + TX - 1) / TX) * ((m->ny + TY - 1) / TY) * ((m->nz + TZ - 1) / TZ);
}

extern "C" int rjp_launch_fill(const rjp_model* m, uint8_t* nverts, rjp_cell* cells,
                               uint8_t* brick_state, int32_t* brick_work, int32_t* ties,
                               int32_t tie_capacity, int32_t* n_ties, int32_t* extents,
                               cudaStream_t stream) {
  const int nxs = m->x_hi - m->x_lo;
  const size_t nray = (size_t)nxs * m->nz;
  init_extents_kernel<<<(unsigned)((nray + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<int2*>(extents), nray);
  if (brick_state != nullptr && brick_work != nullptr) {
    const long long nb = rjp_launch_brick_count(m);
    if (nb <= 0 || nb >= (1LL << 30)) return RJP_ERR_ARG;
    cudaMemsetAsync(brick_work, 0, 4 * sizeof(int32_t), stream);
    fill_classify_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, stream>>>(*m, brick_state,
                                                                          brick_work, nb);
    fill_bricks_kernel<<<148 * 6, FILL_THREADS, 0, stream>>>(*m, nverts, cells, ties,
                                                             tie_capacity, n_ties, extents,
                                                             brick_work);
    return RJP_OK;
  }
  const int tiles_x = (nxs + TX - 1) / TX, tiles_y = (m->ny + TY - 1) / TY;
  const long long sup = (long long)((tiles_x + SBX - 1) / SBX) * ((tiles_y + SBY - 1) / SBY) *
                        ((m->nz + TZ - 1) / TZ);
  if (sup <= 0 || sup > 2147483647LL) return RJP_ERR_ARG;
  fill_grid_kernel<<<(unsigned)sup, FILL_THREADS, 0, stream>>>(
      *m, nverts, cells, brick_state, ties, tie_capacity, n_ties, extents);
  return RJP_OK;
}

extern "C" int rjp_launch_patch(const rjp_model* m, const int64_t* cell_idx,
                                const uint8_t* new_count, int32_t n, uint8_t* nverts,
                                rjp_cell* cells, uint8_t* brick_state, int32_t* extents,
                                cudaStream_t stream) {
  if (n <= 0) return RJP_OK;
  patch_cells_kernel<<<(n + 127) / 128, 128, 0, stream>>>(*m, cell_idx, new_count, n, nverts,
                                                         cells, brick_state, extents);
  return RJP_OK;
}

extern "C" int rjp_launch_override(const rjp_model* m, const uint8_t* nverts, int32_t field,
                                   const double* values, rjp_cell* cells, cudaStream_t stream) {
  const size_t ncell = (size_t)(m->x_hi - m->x_lo) * m->ny * m->nz;
  size_t blocks = (ncell + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (blocks == 0) return RJP_OK;
  override_cells_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*m, nverts, field, values, cells);
  return RJP_OK;
}

extern "C" int rjp_launch_field(const rjp_model* m, const rjp_epoch* ep,
                                const uint8_t* nverts, int32_t field, double* out,
                                cudaStream_t stream) {
  const size_t ncell = (size_t)(m->x_hi - m->x_lo) * m->ny * m->nz;
  size_t blocks = (ncell + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (blocks == 0) return RJP_OK;
  cell_field_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*m, *ep, nverts, field, out);
  return RJP_OK;
}
