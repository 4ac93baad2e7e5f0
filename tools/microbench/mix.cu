// Issue / pipe micro-benchmarks for the mixed-precision channel loop (B200, sm_100a):
// conversion throughput (F2F 64<->32, I2F), shared-memory LDS.128 gathers, and how FFMA,
// integer and LDS instructions co-issue beside DFMA.  Throughput = thread-instructions
// per clock per SM, from CUDA-event time at the measured SM clock (clock64 delta).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu && ./mix
#include <cstdio>
#include <cuda_runtime.h>

#define REP8(X) X X X X X X X X

enum { T_DFMA, T_FFMA, T_F2F_DS, T_F2F_SD, T_I2F, T_F2I, T_D2I, T_DFMA_FFMA1, T_DFMA_FFMA2,
       T_DFMA_FFMA4, T_LDS128, T_LDS128_FFMA4, T_DFMA_INT2, T_MUFU_RCP, T_MUFU_EX2, T_DADD,
       T_DFMA_LDS, T_IMAD, T_MIXLOOP, T_COUNT };
const char* names[] = {"DFMA", "FFMA", "F2F.F32.F64+FADD+DADD", "F2F.F64.F32+DADD+FADD", "I2F+FADD+IADD", "F2I+IADD+FADD",
                       "D2I+IADD+DADD", "DFMA+1FFMA", "DFMA+2FFMA", "DFMA+4FFMA", "LDS.128",
                       "LDS.128+4FFMA", "DFMA+2INT", "MUFU.RCP", "MUFU.EX2", "DADD", "DFMA+LDS.128",
                       "IMAD", "mix(2D+14F+3LDS+4I+1cvt)"};
const int per_iter[] = {8, 8, 8, 8, 8, 8, 8, 16, 24, 40, 8, 40, 24, 8, 8, 8, 16, 8, 8 * 24};

template <int T>
__global__ void k(double* out, int iters, long long* cycles, const float4* gtab) {
  __shared__ float4 tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = gtab[i];
  __syncthreads();
  double d0 = 1.0 + threadIdx.x * 1e-9, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3, d4 = d0 + 4,
         d5 = d0 + 5, d6 = d0 + 6, d7 = d0 + 7;
  float f0 = 1.0f + threadIdx.x * 1e-6f, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4,
        f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
  int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6,
      i7 = i0 + 7;
  const double db = 1.0000001, dc = 1e-9;
  const float fb = 1.0001f, fc = 1e-6f;
  float4 acc4 = {0, 0, 0, 0};
  long long t0 = clock64();
#define DF(n) d##n = fma(d##n, db, dc);
#define FF(n) f##n = fmaf(f##n, fb, fc);
#define ALL8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)
  for (int it = 0; it < iters; ++it) {
    if (T == T_DFMA) { ALL8(DF) }
    if (T == T_DADD) {
#define DA(n) d##n = d##n + dc;
      ALL8(DA)
    }
    if (T == T_FFMA) { ALL8(FF) }
    if (T == T_F2F_DS) {
#define C1(n) f##n += (float)d##n; d##n += dc;
      ALL8(C1)
    }
    if (T == T_F2F_SD) {
#define C2(n) d##n += (double)f##n; f##n += fc;
      ALL8(C2)
    }
    if (T == T_I2F) {
#define C3(n) f##n += (float)i##n; i##n += 3;
      ALL8(C3)
    }
    if (T == T_F2I) {
#define C4(n) i##n += __float2int_rn(f##n); f##n += fc;
      ALL8(C4)
    }
    if (T == T_D2I) {
#define C5(n) i##n += __double2int_rn(d##n); d##n += dc;
      ALL8(C5)
    }
    if (T == T_DFMA_FFMA1) { ALL8(DF) ALL8(FF) }
    if (T == T_DFMA_FFMA2) { ALL8(DF) ALL8(FF) ALL8(FF) }
    if (T == T_DFMA_FFMA4) { ALL8(DF) ALL8(FF) ALL8(FF) ALL8(FF) ALL8(FF) }
    if (T == T_LDS128 || T == T_LDS128_FFMA4 || T == T_DFMA_LDS) {
      // neighbouring lanes hit the same or the next 16-byte entry (like neighbouring channels)
#define LD(n) { const float4 v = tab[(i##n >> 2) & 1023]; acc4.x += v.x; i##n += (int)v.y; }
      // cheaper: pure loads, accumulate one component
#define LD2(n) { const float4 v = tab[((i0 >> 2) + n * 37 + it) & 1023]; f##n += v.x; }
      ALL8(LD2)
      if (T == T_LDS128_FFMA4) { ALL8(FF) ALL8(FF) ALL8(FF) }  // + the FADD of LD2 = 4
      if (T == T_DFMA_LDS) { ALL8(DF) }
    }
    if (T == T_DFMA_INT2) {
#define II(n) i##n = (i##n >> 3) ^ (i##n + 0x1234567);
      ALL8(DF) ALL8(II)
    }
    if (T == T_IMAD) {
#define IM(n) i##n = i##n * 3 + 7;
      ALL8(IM)
    }
    if (T == T_MUFU_RCP) {
#define M1(n) f##n = __frcp_rn(f##n) + fc;
      ALL8(M1)
    }
    if (T == T_MUFU_EX2) {
#define M2(n) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f##n) : "f"(f##n));
      ALL8(M2)
    }
    if (T == T_MIXLOOP) {
      // the planned wing evaluation: 2 fp64, 14 fp32, 3 LDS.128, 4 int, 1 cvt per item, x8
#define MX(n) { d##n = fma(d##n, db, dc); const double m = d##n + 103079215104.0;                \
        int q = __double2loint(m); q = abs(q); const int idx = (q >> 23) & 255;                 \
        const float t = __int_as_float(0x4B000000 | (q & 0x7FFFFF)) - 8388608.0f;               \
        const float4 a = tab[idx * 3], b = tab[idx * 3 + 1], c = tab[idx * 3 + 2];              \
        float h = fmaf(a.x, t, a.y); h = fmaf(h, t, a.z); h = fmaf(h, t, a.w);                  \
        h = fmaf(h, t, b.x); h = fmaf(h, t, b.y);                                               \
        float h3 = fmaf(b.z, t, b.w); h3 = fmaf(h3, t, c.x); h3 = fmaf(h3, t, c.y);             \
        float h5 = fmaf(c.z, t, c.w); h = fmaf(fb, fmaf(fb, h5, h3), h);                         \
        float e = f##n * fmaf(f##n, fc, fb); h *= fc; h = fmaf(h, e, h);                         \
        double hd; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(hd) : "f"(h));                      \
        d##n = fma(db, hd, d##n); }
      ALL8(MX)
    }
  }
  long long t1 = clock64();
  double s = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + i0 +
             i1 + i2 + i3 + i4 + i5 + i6 + i7 + acc4.x;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int T>
void run(double* out, long long* dcyc, const float4* gtab, int blocks, int threads, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<T><<<blocks, threads>>>(out, iters, dcyc, gtab);
  cudaEventRecord(e0);
  k<T><<<blocks, threads>>>(out, iters, dcyc, gtab);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc;
  cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double sm_cycles = (double)cyc;  // cycles block 0 spent in the loop
  const double ops_per_sm = (double)blocks / 148.0 * threads * iters * per_iter[T];
  printf("%-28s blocks=%4d threads=%4d : %8.3f ms  %7.1f counted-instr/clk/SM (by time @1.965GHz)  (block0 %.1f clk/iter) err=%s\n",
         names[T], blocks, threads, ms, (double)blocks*threads*iters*per_iter[T]/(ms*1e-3*1.965e9*148), (double)cyc / iters,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out;
  long long* dcyc;
  float4* gtab;
  cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
  cudaMalloc(&dcyc, 8);
  cudaMalloc(&gtab, sizeof(float4) * 1024);
  cudaMemset(gtab, 0, sizeof(float4) * 1024);
  const int B = 148 * 2, TH = 512, IT = 4000;
#define R(T) run<T>(out, dcyc, gtab, B, TH, IT);
  R(T_DFMA) R(T_DADD) R(T_FFMA) R(T_F2F_DS) R(T_F2F_SD) R(T_I2F) R(T_F2I) R(T_D2I) R(T_DFMA_FFMA1)
  R(T_DFMA_FFMA2) R(T_DFMA_FFMA4) R(T_LDS128) R(T_LDS128_FFMA4) R(T_DFMA_LDS) R(T_DFMA_INT2) R(T_IMAD)
  R(T_MUFU_RCP) R(T_MUFU_EX2) R(T_MIXLOOP)
  return 0;
}
