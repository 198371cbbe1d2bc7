// DMMA.8x8x4 (mma.sync m8n8k4 f64) issue rate on one SM as a function of resident warps and of independent
// accumulator chains per warp: how many warps does it take to keep the FP64 tensor pipe busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate_bench.bin tools/dmma_rate_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k(double* out, int iters, long long* cycles) {
  double c[CHAINS][2];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) c[j][0] = c[j][1] = threadIdx.x * 1e-9 + j;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) dmma(c[j], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(int warps, double* out, long long* cyc) {
  const int iters = 2000;
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  k<CHAINS><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double dmma_per_smsp = (double)iters * CHAINS * warps / 4.0;
  printf("warps/SM %2d chains %d: %.1f cycles per DMMA per sub-partition  (%.1f FMA/clk/SM), per-warp issue interval %.1f cycles\n", warps, CHAINS,
         avg / dmma_per_smsp, 256.0 * 4.0 * dmma_per_smsp / avg, avg / ((double)iters * CHAINS));
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int warps : {4, 8, 12, 16, 24, 32}) {
    run<1>(warps, out, cyc);
    run<2>(warps, out, cyc);
    run<4>(warps, out, cyc);
    run<8>(warps, out, cyc);
  }
  return 0;
}
