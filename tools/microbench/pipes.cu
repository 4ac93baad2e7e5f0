// Pipe micro-benchmarks for B200 (sm_100a): dependent-chain latency and saturated
// throughput of DFMA, FFMA and MUFU.EX2, measured with clock64 / CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, long long* cycles) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
__global__ void ffma_kernel(float* out, int iters, long long* cycles) {
  float a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
  const float b = 1.0001f, c = 1e-6f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fmaf(a[i], b, c);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <typename K, typename T>
void run(const char* name, K kern, T* out, int ilp, int blocks, int threads, int iters,
         long long* dcyc, double flop_per_op) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, iters, dcyc);
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(out, iters, dcyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long cyc;
  cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  double ops = (double)blocks * threads * iters * ilp;
  printf("%-10s ILP=%d blocks=%5d threads=%4d : %8.3f ms  %8.2f Gop/s (%7.2f TFLOP/s)  "
         "cycles/iter(block0)=%.2f\n", name, ilp, blocks, threads, ms, ops / ms / 1e6,
         ops * flop_per_op / ms / 1e9, (double)cyc / iters);
}

int main() {
  double* dout;
  float* fout;
  long long* dcyc;
  cudaMalloc(&dout, sizeof(double) * 148 * 64 * 1024);
  cudaMalloc(&fout, sizeof(float) * 148 * 64 * 1024);
  cudaMalloc(&dcyc, sizeof(long long));
  const int iters = 20000;
  // latency: one warp, one chain
  run("DFMA", dfma_kernel<1>, dout, 1, 1, 32, iters, dcyc, 2);
  run("DFMA", dfma_kernel<2>, dout, 2, 1, 32, iters, dcyc, 2);
  run("DFMA", dfma_kernel<4>, dout, 4, 1, 32, iters, dcyc, 2);
  run("DFMA", dfma_kernel<8>, dout, 8, 1, 32, iters, dcyc, 2);
  run("FFMA", ffma_kernel<1>, fout, 1, 1, 32, iters, dcyc, 2);
  // one SM sub-partition's worth of warps, varying occupancy (threads per block on 1 SM)
  for (int th : {128, 256, 512, 1024}) run("DFMA 1SM", dfma_kernel<1>, dout, 1, 1, th, iters, dcyc, 2);
  for (int th : {128, 256, 512, 1024}) run("DFMA 1SM", dfma_kernel<4>, dout, 4, 1, th, iters, dcyc, 2);
  // whole chip throughput
  run("DFMA chip", dfma_kernel<4>, dout, 4, 148 * 4, 512, iters, dcyc, 2);
  run("DFMA chip", dfma_kernel<8>, dout, 8, 148 * 2, 1024, iters, dcyc, 2);
  run("FFMA chip", ffma_kernel<8>, fout, 8, 148 * 2, 1024, iters, dcyc, 2);
  return 0;
}
