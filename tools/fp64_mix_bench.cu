// fp64_mix_bench.cu -- what limits a DFMA stream on B200?  (diagnostic, not product)
//   mode 0: DFMA only, operands in registers
//   mode 1: DFMA + one broadcast LDS.128 per 4 DFMA (matrix entry from shared memory)
//   mode 2: DFMA + one LDCU.128 (uniform constant load) per 4 DFMA
//   mode 3: mma.sync m8n8k4 f64 (DMMA)
//   mode 4: mma.sync m16n8k16 f64
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_bench fp64_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Params { double2 m[1024]; };

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, const __grid_constant__ Params P) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = P.m[i];
  __syncthreads();
  double ax[8], ay[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ax[i] = threadIdx.x * 1e-3 + i; ay[i] = 0.5 * i; }
  const double vx = 1.0 + threadIdx.x * 1e-9, vy = 0.25;
  if (MODE <= 2) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          double2 m;
          if (MODE == 0) m = make_double2(1.0000001, 1e-9);
          else if (MODE == 1) m = sm[(it & 3) * 128 + c * 8 + r];
          else m = P.m[(it & 3) * 128 + c * 8 + r];
          ax[r] = fma(m.x, vx, ax[r]);
          ax[r] = fma(-m.y, vy, ax[r]);
          ay[r] = fma(m.x, vy, ay[r]);
          ay[r] = fma(m.y, vx, ay[r]);
        }
      }
    }
  } else if (MODE == 3) {
    double a = vx, b = vy;
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = ax[i]; c1[i] = ay[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { ax[i] = c0[i]; ay[i] = c1[i]; }
  } else {
    double a[8], b[4], c[4] = {ax[0], ax[1], ay[0], ay[1]}, d2[4] = {ax[2], ax[3], ay[2], ay[3]};
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = vx + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = vy + i;
    for (int it = 0; it < iters; ++it) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(d2[0]), "+d"(d2[1]), "+d"(d2[2]), "+d"(d2[3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    ax[0] = c[0] + c[1] + d2[0] + d2[1]; ay[0] = c[2] + c[3] + d2[2] + d2[3];
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += ax[i] + ay[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int threads, double fma_per_thread_iter) {
  double* out; cudaMalloc(&out, 148 * 1024 * sizeof(double));
  Params P; for (int i = 0; i < 1024; ++i) P.m[i] = make_double2(1.0 + 1e-9 * i, 1e-9);
  int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 50, P);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, P);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = 148.0 * threads * iters * fma_per_thread_iter;
  printf("%-44s threads/SM=%4d: %6.1f FMA/clk/SM (at 1.965 GHz)  %.2f ms  err=%s\n", name, threads, fma / (ms * 1e-3) / 1.965e9 / 148, ms,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int th : {128, 256, 512}) {
    run<0>("DFMA, register operands", th, 16 * 8 * 4);
    run<1>("DFMA + 1 broadcast LDS.128 per 4", th, 16 * 8 * 4);
    run<2>("DFMA + 1 LDCU.128 per 4", th, 16 * 8 * 4);
    run<3>("DMMA m8n8k4 (8 independent accumulators)", th, 8 * 256.0 / 32);
    run<4>("DMMA m16n8k16 (2 independent accumulators)", th, 2 * 2048.0 / 32);
  }
  return 0;
}
