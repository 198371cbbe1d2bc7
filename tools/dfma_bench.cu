#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < CH; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int blocks_per_sm, int threads) {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<148 * blocks_per_sm, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<CH><<<148 * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double dfma = (double)148 * blocks_per_sm * threads * iters * 8.0 * CH;
  printf("chains=%d blocks/SM=%d threads=%d: %.1f DFMA/clk/SM (assuming 1.965 GHz), %.2f TFLOPS\n", CH, blocks_per_sm, threads,
         dfma / (ms * 1e-3) / 1.965e9 / 148, 2 * dfma / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  run<1>(1, 256); run<4>(1, 256); run<8>(1, 256); run<8>(2, 256); run<8>(4, 256); run<16>(2, 256); run<4>(8, 256);
  return 0;
}
