// reduce_kernels.cuh -- probability reductions, the measurement scan and its collapse.
//
// Replaces the reference's serial running sums (QubitRegister.h:169-195, 619-642;
// QubitRegisterCalculator.h:948-1254).  The outcome of a measurement is defined there as the
// first index i with  prob <= acc_i,  acc_i = fl(acc_{i-1} + |a_i|^2)  in strictly sequential fp64.
// The kernels below reproduce that running sum bit for bit, in parallel (see "the reference's
// sequential running sum" further down); |a|^2 is rounded exactly like the reference's -msse2 build
// (norm_rn).  Outcomes are therefore identical to the reference's for every draw.
// Bound: HBM, 16 B per amplitude, read twice (chunk masses, chunk increments).
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

// the same sum over a CONTIGUOUS range of selector bits [first, first + width) fixed to `want` (already shifted into
// place), visiting only the 2^(n_bits - width) amplitudes of that subspace: GetQubitProbability reads half of the
// state (8 B per amplitude of the state, its algorithmic figure), the collapse norm of Measure 2^-width of it
static __global__ void __launch_bounds__(kThreads)
k_subspace_norm2(const amp* __restrict__ psi, uint64_t n_sub, int first, int width, uint64_t want, double* __restrict__ partials) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t low = (1ULL << first) - 1ULL;
  double acc0 = 0, acc1 = 0;
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; w + stride < n_sub; w += 2 * stride) {
    const uint64_t w1 = w + stride;
    const amp a = psi[((w >> first) << (first + width)) | want | (w & low)];
    const amp b = psi[((w1 >> first) << (first + width)) | want | (w1 & low)];
    acc0 += a.x * a.x + a.y * a.y;
    acc1 += b.x * b.x + b.y * b.y;
  }
  if (w < n_sub) {
    const amp a = psi[((w >> first) << (first + width)) | want | (w & low)];
    acc0 += a.x * a.x + a.y * a.y;
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

// ---- the reference's sequential running sum, reproduced exactly and in parallel ----------------------
// The reference measures with  acc_i = fl(acc_{i-1} + p_i),  p_i = |a_i|^2  (QubitRegister.h:172-190, 240-258,
// 619-642; QubitRegisterCalculator.h:948-1254): a strictly sequential fp64 sum, and the outcome of a draw is the
// first i with  prob <= acc_i.  A bit-identical outcome for EVERY draw needs the bit-identical acc_i -- an exact
// (double-double) prefix is not enough when the draw lands within the rounding error of a bin edge, which at 30
// qubits is most of the time.  The sum is nevertheless parallel: while acc stays inside one binade
// [2^e, 2^(e+1)), it is an integer multiple of u = 2^(e-52) and  fl(acc + p) = acc + u * rint(p / u)  unless p/u
// falls exactly on a half (round-half-even then depends on the running parity).  So
//   1. k_chunk_sums        exact (double-double) mass per chunk of 4096 amplitudes -> the binade each chunk starts in
//   2. k_chunk_increments  per chunk: K = sum rint(p_i / u) as an exact integer, flags for ties / overflow
//   3. k_sequential_walk   one block walks the chunks: a run of chunks inside the binade of the running sum is an
//                          integer prefix sum of their K (1024 chunks per step); the chunk that ends a run (tie, or
//                          the sum crosses a power of two: ~once per binade) has its 4096 additions replayed one by
//                          one: acc at every chunk start, exact
//   4. k_resolve_draws     per draw: binary search over the chunk starts, then the reference's own loop inside one chunk.
// Cost: two passes over the state + O(chunks) sequential work, for any number of draws.
static constexpr unsigned long long kNoOutcome = ~0ULL;

// exclusive prefix (hi word of the double-double) of the chunk masses, `offset` added: predicts the binade
static __global__ void __launch_bounds__(1024)
k_chunk_prefix(const dd* __restrict__ sums, uint64_t n_chunks, dd offset, double* __restrict__ prefix_hi, dd* __restrict__ total) {
  __shared__ dd tot[1024];
  const int t = threadIdx.x, T = blockDim.x;
  const uint64_t per = (n_chunks + T - 1) / T;
  const uint64_t lo = (uint64_t)t * per;
  const uint64_t hi = (lo + per < n_chunks) ? lo + per : n_chunks;
  dd acc = dd_make(0, 0);
  for (uint64_t c = lo; c < hi; ++c) acc = dd_add(acc, sums[c]);
  tot[t] = acc;
  __syncthreads();
  if (t == 0) {
    dd run = offset;
    for (int k = 0; k < T; ++k) {
      const dd v = tot[k];
      tot[k] = run;
      run = dd_add(run, v);
    }
    *total = run;
  }
  __syncthreads();
  dd run = tot[t];
  for (uint64_t c = lo; c < hi; ++c) {
    prefix_hi[c] = run.hi;
    run = dd_add(run, sums[c]);
  }
}

// flag bits of a chunk
constexpr int kChunkZero = 1;   // every |a|^2 of the chunk is 0: acc does not move
constexpr int kChunkSlow = 2;   // tie, overflow, or no usable binade: replay one by one

static __global__ void __launch_bounds__(kThreads)
k_chunk_increments(const amp* __restrict__ psi, uint64_t n, uint64_t n_chunks, const double* __restrict__ prefix_hi,
                   unsigned long long* __restrict__ K, int* __restrict__ flags) {
  __shared__ unsigned long long shk[kThreads / 32];
  __shared__ int shf[kThreads / 32];
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t lo = c << kChunkLog2;
    const uint64_t hi = (lo + kChunk < n) ? lo + kChunk : n;
    const double start = prefix_hi[c];
    const int e = start > 0.0 ? ilogb(start) : -5000;
    const bool usable = e > -900;           // p * 2^(52 - e) must not overflow; subnormal sums take the slow path
    unsigned long long k = 0;
    int slow = usable ? 0 : 1, nonzero = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const double p = norm_rn(psi[i]);
      if (p != 0.0) nonzero = 1;
      if (usable) {
        const double x = scalbn(p, 52 - e);  // exact: a power-of-two scale
        if (!(x < 4503599627370496.0)) {     // >= 2^52: this element alone leaves the binade
          slow = 1;
        } else {
          const double f = floor(x);
          if (x - f == 0.5) slow = 1;        // round-half-even depends on the running parity
          k += (unsigned long long)rint(x);
        }
      }
    }
    // block reduction (exact integers; 4096 terms < 2^52 each cannot overflow 64 bits)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      k += __shfl_down_sync(0xffffffffu, k, o);
      slow |= __shfl_down_sync(0xffffffffu, slow, o);
      nonzero |= __shfl_down_sync(0xffffffffu, nonzero, o);
    }
    if ((threadIdx.x & 31) == 0) {
      shk[threadIdx.x >> 5] = k;
      shf[threadIdx.x >> 5] = slow | (nonzero << 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long kk = 0;
      int ff = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        kk += shk[w];
        ff |= shf[w];
      }
      K[c] = kk;
      flags[c] = ((ff & 2) ? 0 : kChunkZero) | ((ff & 1) ? kChunkSlow : 0);
    }
    __syncthreads();
  }
}

// One block.  acc_start[c] = the reference's running sum before element c * 4096 (acc_start[n_chunks] = after the
// last element), starting from `start` (the sum the lower ranks ended with on a sharded register).
// A run of chunks that stays inside the binade of the running sum is resolved by the whole block at once: with
// e0 = ilogb(carry), every chunk whose predicted binade is e0 has its K in units of u = 2^(e0-52), so the sum before
// chunk j of the run is  carry + u * (K_0 + ... + K_{j-1})  exactly (an integer prefix sum), as long as that value is
// still below 2^(e0+1).  The first chunk that breaks the run -- tie / overflow flag, a predicted binade other than
// e0, or a sum that leaves the binade -- is replayed element by element by thread 0 and the next run starts behind it.
constexpr int kWalkThreads = 1024;
static __global__ void __launch_bounds__(kWalkThreads)
k_sequential_walk(const amp* __restrict__ psi, uint64_t n, uint64_t n_chunks, const double* __restrict__ prefix_hi,
                  const unsigned long long* __restrict__ K, const int* __restrict__ flags, double start, double* __restrict__ acc_start,
                  unsigned long long* __restrict__ n_slow_out) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ int warp_first[32];
  __shared__ double sP[kChunk];
  __shared__ double carry_s;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, n_warps = blockDim.x >> 5;
  constexpr unsigned long long kClamp = 1ULL << 53;  // a chunk this heavy leaves any binade; clamping keeps the prefix sum inside 64 bits
  unsigned long long n_slow = 0;
  double carry = start;  // uniform across the block
  uint64_t c = 0;
  while (c < n_chunks) {
    const int cnt = (int)((n_chunks - c < (uint64_t)blockDim.x) ? (n_chunks - c) : (uint64_t)blockDim.x);
    const bool have = t < cnt;
    const int e0 = carry > 0.0 ? ilogb(carry) : -6000;  // -6000: no binade yet, the first non-zero chunk is replayed
    int f = kChunkZero, e = -5000;
    unsigned long long inc = 0;
    if (have) {
      f = flags[c + t];
      const double ph = prefix_hi[c + t];
      e = ph > 0.0 ? ilogb(ph) : -5000;
      if (!(f & kChunkZero)) {
        const unsigned long long k = K[c + t];
        inc = k < kClamp ? k : kClamp;
      }
    }
    const bool moves = have && !(f & kChunkZero);
    bool fails = moves && ((f & kChunkSlow) || e != e0);
    // inclusive prefix sum of the increments over the block
    unsigned long long incl = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const unsigned long long mine = lane < n_warps ? warp_tot[lane] : 0ULL;
      unsigned long long run = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += v;
      }
      warp_tot[lane] = run - mine;  // exclusive
    }
    __syncthreads();
    incl += warp_tot[wid];
    // (double)incl is exact below 2^53; above, the sum is outside the binade whatever the rounding
    const double before = carry + scalbn((double)(incl - inc), e0 - 52);
    const double after = carry + scalbn((double)incl, e0 - 52);
    if (moves && !fails && ilogb(after) != e0) fails = true;
    // first chunk of the tile that breaks the run
    const unsigned ballot = __ballot_sync(0xffffffffu, fails);
    if (lane == 0) warp_first[wid] = ballot ? (wid * 32 + __ffs((int)ballot) - 1) : (1 << 30);
    __syncthreads();
    if (wid == 0) {
      int v = lane < n_warps ? warp_first[lane] : (1 << 30);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0) warp_first[0] = v;
    }
    __syncthreads();
    const int first = warp_first[0] < cnt ? warp_first[0] : cnt;  // chunks [0, first) of the tile are resolved
    if (have && t <= first) acc_start[c + t] = before;             // (t == first: the replayed chunk starts here)
    if (first < cnt) {
      if (t == first) carry_s = before;
      // cooperative load of the chunk's probabilities, then thread 0 adds them in order like the reference
      const uint64_t lo = (c + (uint64_t)first) << kChunkLog2;
      for (int k2 = t; k2 < (int)kChunk; k2 += blockDim.x) {
        const uint64_t i = lo + k2;
        sP[k2] = (i < n) ? norm_rn(psi[i]) : 0.0;
      }
      __syncthreads();
      if (t == 0) {
        double acc = carry_s;
        for (int k2 = 0; k2 < (int)kChunk; ++k2) acc = __dadd_rn(acc, sP[k2]);
        carry_s = acc;
      }
      ++n_slow;
      c += (uint64_t)first + 1;
    } else {
      if (t == cnt - 1) carry_s = after;
      c += (uint64_t)cnt;
    }
    __syncthreads();
    carry = carry_s;
    __syncthreads();  // carry_s, warp_tot and warp_first are rewritten by the next tile
  }
  if (t == 0) {
    acc_start[n_chunks] = carry;
    if (n_slow_out) *n_slow_out = n_slow;
  }
}

// One thread per draw.  outcome = first local index i with probs[d] <= acc_i, or kNoOutcome when the draw lies
// outside (acc_start[0], acc_start[n_chunks]] (another rank's range, or beyond the total).
static __global__ void __launch_bounds__(128)
k_resolve_draws(const amp* __restrict__ psi, uint64_t n, uint64_t n_chunks, const double* __restrict__ acc_start,
                const double* __restrict__ probs, uint64_t count, unsigned long long* __restrict__ outcomes) {
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= count) return;
  const double prob = probs[d];
  unsigned long long out = kNoOutcome;
  if (prob > acc_start[0] && prob <= acc_start[n_chunks]) {
    // last chunk c with acc_start[c] < prob (acc_start is non-decreasing): the hit is inside it
    uint64_t lo = 0, hi = n_chunks;  // invariant: acc_start[lo] < prob <= acc_start[hi]
    while (hi - lo > 1) {
      const uint64_t mid = (lo + hi) >> 1;
      if (acc_start[mid] < prob) lo = mid;
      else hi = mid;
    }
    double acc = acc_start[lo];
    const uint64_t b = lo << kChunkLog2;
    const uint64_t e = (b + kChunk < n) ? b + kChunk : n;
    for (uint64_t i = b; i < e; ++i) {
      acc = __dadd_rn(acc, norm_rn(psi[i]));
      if (prob <= acc) {
        out = i;
        break;
      }
    }
  }
  outcomes[d] = out;
}

}  // namespace qcsim
