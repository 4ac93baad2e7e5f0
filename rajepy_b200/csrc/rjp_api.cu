// extern "C" entry points declared in include/rajepy_b200.h: argument validation,
// launch, CUDA error capture.  No state besides the thread-local last-error string.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "../../include/rajepy_b200.h"

extern "C" {
int rjp_launch_fill(const rjp_model*, uint8_t*, rjp_cell*, uint8_t*, int32_t*, int32_t*, int32_t,
                    int32_t*, int32_t*, cudaStream_t);
long long rjp_launch_brick_count(const rjp_model*);
int rjp_launch_patch(const rjp_model*, const int64_t*, const uint8_t*, int32_t, uint8_t*,
                     rjp_cell*, uint8_t*, int32_t*, cudaStream_t);
int rjp_launch_field(const rjp_model*, const rjp_epoch*, const uint8_t*, int32_t, double*,
                     cudaStream_t);
int rjp_launch_integrate(const rjp_model*, const rjp_epoch*, const rjp_continuum*,
                         const rjp_cell*, const int32_t*, const int32_t*, const int32_t*,
                         int, double*, double*, double*, int32_t*, const rjp_line*,
                         const rjp_channels*, int, int, double, double*, double*, long long,
                         long long, const double*, const double*, void*, long long, cudaStream_t,
                         cudaStream_t);
long long rjp_launch_line_scratch_bytes(long long, long long);
int rjp_launch_fill_missed(const int32_t*, long long, int, long long, long long, long long,
                           long long, double*, double*, int, cudaStream_t);
int rjp_launch_pack_rays(const double*, long long, const int32_t*, int, int, int, double*,
                         cudaStream_t);
int rjp_launch_scatter_rays(const double*, int, const int32_t*, int, int, double*, long long,
                            cudaStream_t);
int rjp_launch_column_totals(const double*, long long, long long, const int32_t*, const int32_t*,
                             int, double*, cudaStream_t);
int rjp_launch_ray_list(const int32_t*, int, int32_t*, int32_t*, int32_t*, cudaStream_t);
int rjp_ray_list_chunk(void);
int rjp_launch_continuum_images(const double*, const double*, const int32_t*, int64_t,
                                const double*, const double*, double, int, double*, double*,
                                double*, cudaStream_t);
int rjp_launch_continuum_images_epochs(const double*, int, const double*, const int32_t*, int64_t,
                                       const double*, const double*, double, int, double*,
                                       double*, double*, cudaStream_t);
int rjp_launch_integrate_epochs(const rjp_model*, const rjp_epoch*, const rjp_continuum*,
                                const rjp_cell*, const int32_t*, const int32_t*, const int32_t*,
                                int, const double*, int, double*, double*, double*, int32_t*,
                                const double*, cudaStream_t);
int rjp_launch_voigt_profile(const double*, const double*, int64_t, double*, cudaStream_t);
int rjp_launch_override(const rjp_model*, const uint8_t*, int32_t, const double*, rjp_cell*,
                        cudaStream_t);
int rjp_launch_los_means(const rjp_model*, const rjp_epoch*, const uint8_t*, const int32_t*,
                         const int32_t*, const int32_t*, double*, cudaStream_t);
}

static thread_local char g_cuda_err[256] = "";

static int check_launch(int st) {
  if (st != RJP_OK) return st;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e),
             cudaGetErrorString(e));
    return RJP_ERR_CUDA;
  }
  return RJP_OK;
}

static bool model_ok(const rjp_model* m) {
  return m && m->nx > 0 && m->ny > 0 && m->nz > 0 && m->x_lo >= 0 && m->x_hi <= m->nx &&
         m->x_lo < m->x_hi && m->cs > 0.0;
}

extern "C" const char* rjp_strerror(int status) {
  switch (status) {
    case RJP_OK: return "ok";
    case RJP_ERR_ARG: return "invalid argument";
    case RJP_ERR_CUDA: return "CUDA runtime error";
    case RJP_ERR_CAPACITY: return "caller-provided list too small";
    case RJP_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

extern "C" const char* rjp_last_cuda_error(void) { return g_cuda_err; }

extern "C" int rjp_abi_version(void) { return RJP_ABI_VERSION; }

extern "C" int rjp_struct_sizes(int32_t* model, int32_t* epoch, int32_t* continuum,
                                int32_t* line, int32_t* channels, int32_t* cell) {
  if (model) *model = (int32_t)sizeof(rjp_model);
  if (epoch) *epoch = (int32_t)sizeof(rjp_epoch);
  if (continuum) *continuum = (int32_t)sizeof(rjp_continuum);
  if (line) *line = (int32_t)sizeof(rjp_line);
  if (channels) *channels = (int32_t)sizeof(rjp_channels);
  if (cell) *cell = (int32_t)sizeof(rjp_cell);
  return RJP_OK;
}

extern "C" int64_t rjp_brick_count(const rjp_model* m) {
  if (!model_ok(m)) return RJP_ERR_ARG;
  return (int64_t)rjp_launch_brick_count(m);
}

extern "C" int rjp_fill_grid(const rjp_model* m, uint8_t* nverts, rjp_cell* cells,
                             uint8_t* brick_state, int32_t* brick_work, int32_t* ties,
                             int32_t tie_capacity, int32_t* n_ties, int32_t* extents,
                             void* stream) {
  if (!model_ok(m) || !nverts || !cells || !n_ties || !extents || tie_capacity < 0 ||
      (tie_capacity > 0 && !ties))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_fill(m, nverts, cells, brick_state, brick_work, ties,
                                      tie_capacity, n_ties, extents, (cudaStream_t)stream));
}

extern "C" int rjp_patch_cells(const rjp_model* m, const int64_t* cell_idx,
                               const uint8_t* new_count, int32_t n, uint8_t* nverts,
                               rjp_cell* cells, uint8_t* brick_state, int32_t* extents,
                               void* stream) {
  if (!model_ok(m) || n < 0 || (n > 0 && (!cell_idx || !new_count)) || !nverts || !cells ||
      !extents)
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_patch(m, cell_idx, new_count, n, nverts, cells, brick_state,
                                       extents, (cudaStream_t)stream));
}

extern "C" int rjp_cell_field(const rjp_model* m, const rjp_epoch* ep, const uint8_t* nverts,
                              int32_t field, double* out, void* stream) {
  if (!model_ok(m) || !ep || !nverts || !out || field < 0 || field >= RJP_FIELD_COUNT)
    return RJP_ERR_ARG;
  if (ep->n_blue < 0 || ep->n_blue > RJP_MAX_BURSTS || ep->n_red < 0 ||
      ep->n_red > RJP_MAX_BURSTS)
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_field(m, ep, nverts, field, out, (cudaStream_t)stream));
}

extern "C" int64_t rjp_ray_list_chunks(int64_t nray) {
  if (nray < 0) return RJP_ERR_ARG;
  const int64_t c = rjp_ray_list_chunk();
  return (nray + c - 1) / c;
}

extern "C" int rjp_ray_list(const int32_t* extents, int64_t nray, int32_t* list,
                            int32_t* chunk_counts, int32_t* n_active, void* stream) {
  if (!extents || !list || !n_active || nray < 0 || nray > 2147483647LL ||
      (nray > 0 && !chunk_counts))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_ray_list(extents, (int)nray, list, chunk_counts, n_active,
                                          (cudaStream_t)stream));
}

extern "C" int rjp_integrate(const rjp_model* m, const rjp_epoch* ep, const rjp_continuum* ct,
                             const rjp_cell* cells, const int32_t* extents,
                             const int32_t* ray_list, const int32_t* n_active,
                             int32_t n_active_hint, double* em,
                             double* kff, double* tsum, int32_t* tcount, const rjp_line* ln,
                             const rjp_channels* ch, int32_t nchan, int32_t contsub,
                             double* tau_rrl, double* flux_rrl, int64_t cube_plane,
                             int64_t cube_offset, const double* travel_cells,
                             const double* vlos_cells, void* line_scratch,
                             int64_t line_max_cells, void* stream, void* stream2) {
  if (!model_ok(m) || !ep || !ct || !cells || !em || !kff || !tsum || !tcount || nchan < 0)
    return RJP_ERR_ARG;
  if (ep->n_blue < 0 || ep->n_blue > RJP_MAX_BURSTS || ep->n_red < 0 ||
      ep->n_red > RJP_MAX_BURSTS)
    return RJP_ERR_ARG;
  if (nchan > 0) {
    if (!ln || !ch || !ch->dnu || !ch->nu || !ch->cff || !ch->aff || !ch->bnu)
      return RJP_ERR_ARG;
    if (!tau_rrl && !flux_rrl) return RJP_ERR_ARG;
    if (!extents || !ray_list || !n_active) return RJP_ERR_ARG;
    if (cube_plane < 0 || cube_offset < 0) return RJP_ERR_ARG;
    if (!line_scratch || line_max_cells < 0) return RJP_ERR_ARG;
  }
  return check_launch(rjp_launch_integrate(m, ep, ct, cells, extents, ray_list, n_active,
                                           n_active_hint,
                                           em, kff, tsum, tcount,
                                           ln, ch, nchan, contsub, ln ? ln->dn_max : 0.0,
                                           tau_rrl, flux_rrl, cube_plane, cube_offset,
                                           travel_cells, vlos_cells, line_scratch,
                                           line_max_cells,
                                           (cudaStream_t)stream, (cudaStream_t)stream2));
}

extern "C" int64_t rjp_line_scratch_bytes(const rjp_model* m, int64_t max_cells) {
  if (!model_ok(m) || max_cells < 0) return RJP_ERR_ARG;
  return (int64_t)rjp_launch_line_scratch_bytes((long long)(m->x_hi - m->x_lo) * m->nz,
                                                (long long)max_cells);
}

extern "C" int rjp_continuum_images(const double* kff, const double* tsum,
                                    const int32_t* tcount, int64_t npix, const double* cff,
                                    const double* iff, double omega_jy, int32_t nfreq,
                                    double* tau, double* intensity, double* flux,
                                    void* stream) {
  if (!kff || !tsum || !tcount || npix < 0 || nfreq < 0 || (nfreq > 0 && (!cff || !iff)))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_continuum_images(kff, tsum, tcount, npix, cff, iff, omega_jy,
                                                  nfreq, tau, intensity, flux,
                                                  (cudaStream_t)stream));
}

extern "C" int rjp_continuum_images_epochs(const double* kff, int32_t n_epochs,
                                           const double* tsum, const int32_t* tcount,
                                           int64_t npix, const double* cff, const double* iff,
                                           double omega_jy, int32_t nfreq, double* tau,
                                           double* intensity, double* flux, void* stream) {
  if (!kff || !tsum || !tcount || npix < 0 || nfreq < 0 || n_epochs < 0 ||
      (nfreq > 0 && (!cff || !iff)))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_continuum_images_epochs(kff, n_epochs, tsum, tcount, npix, cff,
                                                         iff, omega_jy, nfreq, tau, intensity,
                                                         flux, (cudaStream_t)stream));
}

extern "C" int rjp_integrate_epochs(const rjp_model* m, const rjp_epoch* ep,
                                    const rjp_continuum* ct, const rjp_cell* cells,
                                    const int32_t* extents, const int32_t* ray_list,
                                    const int32_t* n_active, int32_t n_active_hint,
                                    const double* times, int32_t n_epochs, double* em,
                                    double* kff, double* tsum, int32_t* tcount,
                                    const double* travel_cells, void* stream) {
  if (!model_ok(m) || !ep || !ct || !cells || !extents || !ray_list || !n_active ||
      n_epochs < 0 || (n_epochs > 0 && (!times || !kff || !tsum || !tcount)))
    return RJP_ERR_ARG;
  if (ep->n_blue < 0 || ep->n_blue > RJP_MAX_BURSTS || ep->n_red < 0 ||
      ep->n_red > RJP_MAX_BURSTS)
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_integrate_epochs(m, ep, ct, cells, extents, ray_list, n_active,
                                                  n_active_hint, times, n_epochs, em, kff, tsum,
                                                  tcount, travel_cells, (cudaStream_t)stream));
}

extern "C" int rjp_voigt_profile(const double* x, const double* y, int64_t n, double* out,
                                 void* stream) {
  if (n < 0 || (n > 0 && (!x || !y || !out))) return RJP_ERR_ARG;
  return check_launch(rjp_launch_voigt_profile(x, y, n, out, (cudaStream_t)stream));
}

extern "C" int rjp_fill_missed(const int32_t* extents, int64_t nray, int32_t nchan,
                               int64_t cube_plane, int64_t cube_offset, int64_t skip_lo,
                               int64_t skip_hi, double* tau, double* flux, int32_t light,
                               void* stream) {
  if (nray < 0 || nchan < 0 || (nray > 0 && !extents) || cube_plane < cube_offset + nray ||
      cube_offset < 0 || (!tau && !flux) || skip_lo < 0 || skip_hi < skip_lo)
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_fill_missed(extents, nray, nchan, cube_plane, cube_offset,
                                             skip_lo, skip_hi, tau, flux, light,
                                             (cudaStream_t)stream));
}

extern "C" int rjp_column_totals(const double* cube, int64_t cube_plane, int64_t cube_offset,
                                 const int32_t* ray_list, const int32_t* n_active,
                                 int32_t nchan, double* totals, void* stream) {
  if (nchan < 0 || cube_plane <= 0 || cube_offset < 0 ||
      (nchan > 0 && (!cube || !ray_list || !n_active || !totals)))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_column_totals(cube, cube_plane, cube_offset, ray_list, n_active,
                                               nchan, totals, (cudaStream_t)stream));
}

extern "C" int rjp_pack_rays(const double* cube, int64_t cube_plane, const int32_t* ray_ids,
                             int32_t n, int32_t n_stride, int32_t nchan, double* out,
                             void* stream) {
  if (n < 0 || nchan < 0 || n_stride < n || (n > 0 && nchan > 0 && (!cube || !ray_ids || !out)))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_pack_rays(cube, cube_plane, ray_ids, n, n_stride, nchan, out,
                                           (cudaStream_t)stream));
}

extern "C" int rjp_scatter_rays(const double* in, int32_t n_stride, const int32_t* ray_ids,
                                int32_t n, int32_t nchan, double* cube, int64_t cube_plane,
                                void* stream) {
  if (n < 0 || nchan < 0 || n_stride < n || (n > 0 && nchan > 0 && (!cube || !ray_ids || !in)))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_scatter_rays(in, n_stride, ray_ids, n, nchan, cube, cube_plane,
                                              (cudaStream_t)stream));
}

extern "C" int rjp_los_means(const rjp_model* m, const rjp_epoch* ep, const uint8_t* nverts,
                             const int32_t* extents, const int32_t* ray_list,
                             const int32_t* n_active, double* out, void* stream) {
  if (!model_ok(m) || !ep || !nverts || !extents || !out || !n_active || !ray_list)
    return RJP_ERR_ARG;
  if (ep->n_blue < 0 || ep->n_blue > RJP_MAX_BURSTS || ep->n_red < 0 ||
      ep->n_red > RJP_MAX_BURSTS)
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_los_means(m, ep, nverts, extents, ray_list, n_active, out,
                                           (cudaStream_t)stream));
}

extern "C" int rjp_override_cells(const rjp_model* m, const uint8_t* nverts, int32_t field,
                                  const double* values, rjp_cell* cells, void* stream) {
  if (!model_ok(m) || !nverts || !values || !cells ||
      (field != RJP_FIELD_TEMP && field != RJP_FIELD_XI))
    return RJP_ERR_ARG;
  return check_launch(rjp_launch_override(m, nverts, field, values, cells,
                                          (cudaStream_t)stream));
}
