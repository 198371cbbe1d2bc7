// common.cuh -- shared device/host helpers for the statevector engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qcsim {

typedef double2 amp;  // one complex<double> amplitude = one 128-bit load/store

constexpr int kThreads = 256;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

__host__ __device__ __forceinline__ amp make_amp(double re, double im) {
  amp r;
  r.x = re;
  r.y = im;
  return r;
}

// complex product, same operand order as the reference's `matrix_entry * amplitude`
__device__ __forceinline__ amp cmul(amp a, amp b) { return make_amp(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ amp cadd(amp a, amp b) { return make_amp(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ amp cmad(amp m, amp a, amp acc) { return cadd(acc, cmul(m, a)); }

// |a|^2 exactly as the reference's -msse2 build rounds std::norm: two rounded products, one
// rounded sum, no FMA contraction (QubitRegister.h:177; SURVEY 8c "low-order-bit caveats").
__device__ __forceinline__ double norm_rn(amp a) { return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); }

// Insert a zero bit at position p (bits >= p shift up by one).
__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t x, int p) {
  const uint64_t low = x & ((1ULL << p) - 1ULL);
  return ((x >> p) << (p + 1)) | low;
}

// positions sorted ascending; NF <= 3
struct FixedBits {
  int n;
  int pos[3];
};

__host__ __device__ __forceinline__ uint64_t scatter_index(uint64_t w, const FixedBits& f) {
  uint64_t x = w;
  if (f.n > 0) x = insert_zero(x, f.pos[0]);
  if (f.n > 1) x = insert_zero(x, f.pos[1]);
  if (f.n > 2) x = insert_zero(x, f.pos[2]);
  return x;
}

// 128-bit streaming accesses.  The state is touched once per pass, so keep it out of L1
// (L2 still merges the two half-sector accesses of a low-qubit pair).
__device__ __forceinline__ amp ld_amp(const amp* p) {
  amp r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_amp(amp* p, amp v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// two adjacent amplitudes moved by one 256-bit LDG/STG (sm_100: LDG.E.ENL2.256)
struct amp2 {
  amp a, b;
};
__device__ __forceinline__ amp2 ld_amp2(const amp* p) {
  amp2 r;
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.a.x), "=d"(r.a.y), "=d"(r.b.x), "=d"(r.b.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_amp2(amp* p, amp2 v) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.a.x), "d"(v.a.y), "d"(v.b.x), "d"(v.b.y)
               : "memory");
}

// ---- shared-memory tiles (tile_kernels.cuh, qft_kernels.cuh) ---------------------------------------
constexpr int kMaxTileBits = 12;   // 2^12 x 16 B = 64 KiB of shared memory per tile
constexpr int kTileThreads = 256;

// 16-byte slot swizzle, linear over XOR: folds every 3-bit group of the index onto the low 3
// bits, so any 8 indices that differ in 3 bits with distinct (position mod 3) hit 8 distinct
// 16-byte bank groups.  The host picks which tile bits the low 3 item-index bits walk over.
__host__ __device__ __forceinline__ uint32_t swz(uint32_t j) { return j ^ ((j >> 3) & 7u) ^ ((j >> 6) & 7u) ^ ((j >> 9) & 7u); }

// ---- double-double (error-free) accumulation, used by the measurement scan -----------------
struct dd {
  double hi, lo;
};
__host__ __device__ __forceinline__ dd dd_make(double hi, double lo) {
  dd r;
  r.hi = hi;
  r.lo = lo;
  return r;
}
#ifdef __CUDA_ARCH__
#define QCSIM_ADD(a, b) __dadd_rn((a), (b))
#define QCSIM_SUB(a, b) __dadd_rn((a), -(b))
#else
#define QCSIM_ADD(a, b) ((a) + (b))
#define QCSIM_SUB(a, b) ((a) - (b))
#endif
// Knuth TwoSum: s + e == a + b exactly
__host__ __device__ __forceinline__ dd two_sum(double a, double b) {
  const double s = QCSIM_ADD(a, b);
  const double bb = QCSIM_SUB(s, a);
  const double e = QCSIM_ADD(QCSIM_SUB(a, QCSIM_SUB(s, bb)), QCSIM_SUB(b, bb));
  return dd_make(s, e);
}
__host__ __device__ __forceinline__ dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  const double lo = QCSIM_ADD(s.lo, a.lo);
  return two_sum(s.hi, lo);
}
__host__ __device__ __forceinline__ dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  dd t = two_sum(a.lo, b.lo);
  double c = QCSIM_ADD(s.lo, t.hi);
  dd v = two_sum(s.hi, c);
  double w = QCSIM_ADD(t.lo, v.lo);
  return two_sum(v.hi, w);
}
// prob <= value(x) for a normalised double-double x
__host__ __device__ __forceinline__ bool dd_reaches(double prob, dd x) {
  return (prob < x.hi) || (prob == x.hi && x.lo >= 0.0);
}

}  // namespace qcsim
