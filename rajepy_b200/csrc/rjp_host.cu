// Host side of the product hand-over: dense (nchan, nx, nz) cubes in HOST memory assembled
// from what actually differs between rays.  94 % of the rays of the BASELINE jet miss the jet
// and carry the constants 0 (tau) / NaN (flux, classes.py:1323-1328 through nanmean), so only
// the packed columns of the jet-crossing rays (rjp_pack_rays) cross PCIe; this routine writes
// the constants with streaming stores from several host threads and drops the columns in.
// No CUDA calls here: plain C++ threads, compiled into the same library.
#include <immintrin.h>
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>
#include "../../include/rajepy_b200.h"

namespace {

struct Run { int64_t ray, k, len; };   // ray_ids[k .. k+len) = ray, ray+1, ...

// 32-byte streaming stores where the CPU has AVX (checked once at run time)
__attribute__((target("avx"))) void fill_stream_avx(double* p, int64_t n, double v) {
  int64_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(p + i) & 31)) p[i++] = v;
  const __m256d vv = _mm256_set1_pd(v);
  for (; i + 16 <= n; i += 16) {
    _mm256_stream_pd(p + i, vv);
    _mm256_stream_pd(p + i + 4, vv);
    _mm256_stream_pd(p + i + 8, vv);
    _mm256_stream_pd(p + i + 12, vv);
  }
  for (; i + 4 <= n; i += 4) _mm256_stream_pd(p + i, vv);
  for (; i < n; ++i) p[i] = v;
}

inline bool have_avx() {
  static const bool v = (__builtin_cpu_init(), __builtin_cpu_supports("avx") != 0);
  return v;
}

// n doubles of `v` at p with non-temporal stores (no read-for-ownership of the destination)
inline void fill_stream(double* p, int64_t n, double v) {
  if (n >= 64 && have_avx()) {
    fill_stream_avx(p, n, v);
    return;
  }
  if (n <= 0) return;
  if (n < 32) {
    for (int64_t i = 0; i < n; ++i) p[i] = v;
    return;
  }
  int64_t i = 0;
  if (reinterpret_cast<uintptr_t>(p) & 15) p[i++] = v;
  const __m128d vv = _mm_set1_pd(v);
  for (; i + 8 <= n; i += 8) {
    _mm_stream_pd(p + i, vv);
    _mm_stream_pd(p + i + 2, vv);
    _mm_stream_pd(p + i + 4, vv);
    _mm_stream_pd(p + i + 6, vv);
  }
  for (; i + 2 <= n; i += 2) _mm_stream_pd(p + i, vv);
  for (; i < n; ++i) p[i] = v;
}

void assemble_planes(double* dst, int64_t c_lo, int64_t c_hi, int64_t c_step, int64_t plane,
                     const std::vector<Run>* runs, const double* cols, int64_t n_stride,
                     double fill) {
  for (int64_t c = c_lo; c < c_hi; c += c_step) {
    double* row = dst + c * plane;
    const double* col = cols ? cols + c * n_stride : nullptr;
    int64_t pos = 0;
    for (const Run& r : *runs) {
      fill_stream(row + pos, r.ray - pos, fill);
      memcpy(row + r.ray, col + r.k, (size_t)r.len * sizeof(double));
      pos = r.ray + r.len;
    }
    fill_stream(row + pos, plane - pos, fill);
  }
  _mm_sfence();
}

}  // namespace

extern "C" int rjp_host_assemble(double* dst_host, int64_t nchan, int64_t plane,
                                 const int32_t* ray_ids_host, int64_t n,
                                 const double* cols_host, int64_t n_stride, double fill,
                                 int32_t nthreads) {
  if (!dst_host || nchan < 0 || plane <= 0 || n < 0 || n > plane || n_stride < n ||
      (n > 0 && (!ray_ids_host || !cols_host)))
    return RJP_ERR_ARG;
  std::vector<Run> runs;
  for (int64_t k = 0; k < n;) {
    const int64_t r0 = ray_ids_host[k];
    if (r0 < 0 || r0 >= plane || (k > 0 && r0 <= ray_ids_host[k - 1])) return RJP_ERR_ARG;
    int64_t j = k + 1;
    while (j < n && ray_ids_host[j] == r0 + (j - k)) ++j;
    if (r0 + (j - k) > plane) return RJP_ERR_ARG;
    runs.push_back({r0, k, j - k});
    k = j;
  }
  int t = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  if (t > nchan) t = (int)(nchan > 0 ? nchan : 1);
  if (t == 1) {
    assemble_planes(dst_host, 0, nchan, 1, plane, &runs, cols_host, n_stride, fill);
    return RJP_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(t);
  for (int i = 0; i < t; ++i)     // interleaved planes: even load whatever nchan is
    pool.emplace_back(assemble_planes, dst_host, (int64_t)i, nchan, (int64_t)t, plane, &runs,
                      cols_host, n_stride, fill);
  for (auto& th : pool) th.join();
  return RJP_OK;
}
