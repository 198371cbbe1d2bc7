// reduce_kernels.cuh -- probability reductions, the measurement scan and its collapse.
//
// Replaces the reference's serial running sums (QubitRegister.h:169-195, 619-642;
// QubitRegisterCalculator.h:948-1254).  The outcome of a measurement is defined there as the
// first index i with  prob <= sum_{j<=i} |a_j|^2.  Here:
//   1. k_chunk_sums   : every chunk of kChunk amplitudes -> its probability mass as an error-free
//                       double-double (warp shuffle + shared-memory block reduction);
//   2. k_find_chunk   : one block scans the chunk masses (double-double prefix) and locates the
//                       chunk where the prefix first reaches prob;
//   3. k_find_in_chunk: one block scans that chunk and returns the index.
// |a|^2 is rounded exactly like the reference's -msse2 build (norm_rn), and the prefixes are
// exact to ~2^-100, so the outcome equals the reference's unless prob lies within the rounding
// error of the reference's own sequential fp64 sum (~sqrt(i) * 1e-16) of a bin edge.
// Bound: HBM, 16 B per amplitude read once.
#pragma once

#include "common.cuh"

namespace qcsim {

constexpr int kChunkLog2 = 12;
constexpr uint64_t kChunk = 1ULL << kChunkLog2;  // amplitudes per scan chunk (64 KiB)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ dd warp_sum_dd(dd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dd t;
    t.hi = __shfl_down_sync(0xffffffffu, v.hi, o);
    t.lo = __shfl_down_sync(0xffffffffu, v.lo, o);
    v = dd_add(v, t);
  }
  return v;
}

// block-wide sum; result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[kThreads / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
__device__ __forceinline__ dd block_sum_dd(dd v) {
  __shared__ dd shd[kThreads / 32];
  v = warp_sum_dd(v);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = v;
  __syncthreads();
  dd r = dd_make(0, 0);
  if (threadIdx.x < 32) {
    if (threadIdx.x < (blockDim.x >> 5)) r = shd[threadIdx.x];
    r = warp_sum_dd(r);
  }
  __syncthreads();
  return r;
}

// sum of |a|^2 over local indices i with ((base + i) & mask) == want -> partials[blockIdx.x]
// mask == 0: squared norm; mask = bit, want = bit: GetQubitProbability (Calculator :1088-1122);
// mask = measured part: the collapse norm of Measure (:980-987, 1151-1158).
static __global__ void __launch_bounds__(kThreads)
k_masked_norm2(const amp* __restrict__ psi, uint64_t n, uint64_t base, uint64_t mask, uint64_t want,
               double* __restrict__ partials) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double acc0 = 0, acc1 = 0;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n; i += 2 * stride) {
    const amp a = psi[i], b = psi[i + stride];
    if (((base + i) & mask) == want) acc0 += a.x * a.x + a.y * a.y;
    if (((base + i + stride) & mask) == want) acc1 += b.x * b.x + b.y * b.y;
  }
  if (i < n) {
    const amp a = psi[i];
    if (((base + i) & mask) == want) acc0 += a.x * a.x + a.y * a.y;
  }
  const double s = block_sum(acc0 + acc1);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// conj(a) . b partials (stateFidelity / ExpectationValue, QubitRegister.h:527-534, 646-660)
static __global__ void __launch_bounds__(kThreads)
k_inner_product(const amp* __restrict__ a, const amp* __restrict__ b, uint64_t n, double* __restrict__ partials) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double re = 0, im = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const amp x = a[i], y = b[i];
    re += x.x * y.x + x.y * y.y;
    im += x.x * y.y - x.y * y.x;
  }
  const double sr = block_sum(re);
  const double si = block_sum(im);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = sr;
    partials[2 * blockIdx.x + 1] = si;
  }
}

// deterministic final reduction of `count` partials (x `width` interleaved components)
static __global__ void __launch_bounds__(kThreads) k_final_sum(const double* __restrict__ partials, int count, int width,
                                                        double* __restrict__ out) {
  for (int c = 0; c < width; ++c) {
    double acc = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partials[i * width + c];
    const double s = block_sum(acc);
    if (threadIdx.x == 0) out[c] = s;
  }
}

// ---- measurement scan --------------------------------------------------------------------------

// chunk c = amplitudes [c*kChunk, (c+1)*kChunk) (clipped to n): error-free probability mass
static __global__ void __launch_bounds__(kThreads)
k_chunk_sums(const amp* __restrict__ psi, uint64_t n, uint64_t n_chunks, dd* __restrict__ sums) {
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t lo = c << kChunkLog2;
    const uint64_t hi = (lo + kChunk < n) ? lo + kChunk : n;
    dd acc = dd_make(0, 0);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = dd_add_d(acc, norm_rn(psi[i]));
    const dd s = block_sum_dd(acc);
    if (threadIdx.x == 0) sums[c] = s;
  }
}

struct ScanResult {
  uint64_t chunk;    // chunk where the prefix first reaches prob (valid if found)
  uint64_t index;    // local index of the outcome (valid if found, after k_find_in_chunk)
  dd before;         // exact mass of everything before `chunk`, including `offset`
  dd total;          // offset + exact mass of the whole local slice
  double margin;     // min distance of prob to the two edges of the selected bin
  int found;
};

// Single block.  `offset` = exact mass owned by lower ranks (0 on one GPU).
static __global__ void __launch_bounds__(1024)
k_find_chunk(const dd* __restrict__ sums, uint64_t n_chunks, dd offset, double prob, ScanResult* __restrict__ res) {
  __shared__ dd tot[1024];
  __shared__ unsigned long long first_chunk;
  const int t = threadIdx.x, T = blockDim.x;
  const uint64_t per = (n_chunks + T - 1) / T;
  const uint64_t lo = (uint64_t)t * per;
  const uint64_t hi = (lo + per < n_chunks) ? lo + per : n_chunks;
  dd acc = dd_make(0, 0);
  for (uint64_t c = lo; c < hi; ++c) acc = dd_add(acc, sums[c]);
  tot[t] = acc;
  if (t == 0) first_chunk = ~0ULL;
  __syncthreads();
  // exclusive prefix of the per-thread strips (serial over <= 1024 entries, one thread: cheap and exact-ordered)
  if (t == 0) {
    dd run = offset;
    for (int k = 0; k < T; ++k) {
      const dd v = tot[k];
      tot[k] = run;
      run = dd_add(run, v);
    }
    res->total = run;
  }
  __syncthreads();
  dd run = tot[t];
  for (uint64_t c = lo; c < hi; ++c) {
    const dd nxt = dd_add(run, sums[c]);
    if (dd_reaches(prob, nxt)) {
      atomicMin(&first_chunk, (unsigned long long)c);
      break;
    }
    run = nxt;
  }
  __syncthreads();
  // the winning strip re-derives the exact prefix before its chunk
  if (first_chunk != ~0ULL && first_chunk >= lo && first_chunk < hi) {
    dd r2 = tot[t];
    for (uint64_t c = lo; c < first_chunk; ++c) r2 = dd_add(r2, sums[c]);
    res->chunk = first_chunk;
    res->before = r2;
    res->found = 1;
  }
  if (t == 0 && first_chunk == ~0ULL) {
    res->found = 0;
    res->chunk = 0;
    res->index = 0;
    res->margin = 0;
  }
}

// Single block of kThreads: locate the outcome inside res->chunk.
static __global__ void __launch_bounds__(kThreads)
k_find_in_chunk(const amp* __restrict__ psi, uint64_t n, double prob, ScanResult* __restrict__ res) {
  if (!res->found) return;
  constexpr int PER = (int)(kChunk / kThreads);
  __shared__ dd tot[kThreads];
  __shared__ unsigned long long first_idx;
  const uint64_t lo = res->chunk << kChunkLog2;
  const int t = threadIdx.x;
  double p[PER];
  dd acc = dd_make(0, 0);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint64_t i = lo + (uint64_t)t * PER + k;
    p[k] = (i < n) ? norm_rn(psi[i]) : 0.0;
    acc = dd_add_d(acc, p[k]);
  }
  tot[t] = acc;
  if (t == 0) first_idx = ~0ULL;
  __syncthreads();
  if (t == 0) {
    dd run = res->before;
    for (int k = 0; k < kThreads; ++k) {
      const dd v = tot[k];
      tot[k] = run;
      run = dd_add(run, v);
    }
  }
  __syncthreads();
  dd run = tot[t];
  int hit = -1;
  dd prev = run, at = run;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const dd nxt = dd_add_d(run, p[k]);
    if (hit < 0 && dd_reaches(prob, nxt)) {
      hit = k;
      prev = run;
      at = nxt;
    }
    run = nxt;
  }
  if (hit >= 0) atomicMin(&first_idx, (unsigned long long)(lo + (uint64_t)t * PER + hit));
  __syncthreads();
  if (hit >= 0 && first_idx == lo + (uint64_t)t * PER + hit) {
    res->index = first_idx;
    const double up = (at.hi - prob) + at.lo;    // distance to the edge that was reached
    const double dn = (prob - prev.hi) - prev.lo; // distance to the previous edge
    res->margin = up < dn ? up : dn;
  }
  if (t == 0 && first_idx == ~0ULL) res->found = 0;  // cannot happen when k_find_chunk found it
}

// Strict replay of the reference's sequential fp64 running sum (QubitRegister.h:172-190) over
// local indices [0, upto]: acc_{i} = fl(acc_{i-1} + p_i), starting from `start`.  One block;
// the adds are one dependent chain on a single thread, the loads are cooperative.
// Writes the first index with prob <= acc (or ~0) and the final acc.
static __global__ void __launch_bounds__(kThreads)
k_sequential_scan(const amp* __restrict__ psi, uint64_t n, double start, double prob, unsigned long long* __restrict__ out_idx,
                  double* __restrict__ out_acc) {
  __shared__ double buf[2][kThreads * 8];
  constexpr int TILE = kThreads * 8;
  double acc = start;
  unsigned long long found = ~0ULL;
  const uint64_t n_tiles = (n + TILE - 1) / TILE;
  // prefetch tile 0
  for (int k = threadIdx.x; k < TILE; k += blockDim.x) {
    const uint64_t i = k;
    buf[0][k] = (i < n) ? norm_rn(psi[i]) : 0.0;
  }
  __syncthreads();
  __shared__ int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  for (uint64_t tile = 0; tile < n_tiles && !stop; ++tile) {
    const int cur = (int)(tile & 1);
    if (threadIdx.x == 0) {
      const uint64_t lo = tile * TILE;
      const int lim = (int)((n - lo < (uint64_t)TILE) ? (n - lo) : (uint64_t)TILE);
      for (int k = 0; k < lim; ++k) {
        acc = __dadd_rn(acc, buf[cur][k]);
        if (prob <= acc) {
          found = lo + k;
          stop = 1;
          break;
        }
      }
    } else if (tile + 1 < n_tiles) {
      const uint64_t lo = (tile + 1) * TILE;
      for (int k = threadIdx.x - 1; k < TILE; k += blockDim.x - 1) {
        const uint64_t i = lo + k;
        buf[cur ^ 1][k] = (i < n) ? norm_rn(psi[i]) : 0.0;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out_idx = found;
    *out_acc = acc;
  }
}

}  // namespace qcsim
