// gate_kernels.cuh -- single-pass, in-place gate kernels (one gate application per HBM pass).
//
// Every 1/2/3-qubit gate of the reference (QubitRegisterCalculator.h:39-939) is lowered by
// classify.h to one of four shapes, each touching only the amplitudes that actually change:
//
//   PAIR  : a 2x2 matrix on the amplitude pair (base|or_lo, base|or_hi); covers dense/anti-
//           diagonal 1q gates, every controlled-U / CC-U (CNOT, CH, CRx, Toffoli ...) and the
//           swap family (SWAP, iSWAP, iSWAPdg, Fredkin) where the pair is (|01>, |10>).
//   DENSE : a 4x4 / 8x8 matrix on 2 / 3 target qubits (+ optional control).
//   DIAG  : in-place multiply by a <=8-entry table selected by <=3 qubits, on the subspace where
//           the control qubits are 1 (Rz, S, T, CZ, CPhaseShift, CCZ ...).
//
// Unlike the reference (out-of-place into resultsStorage, QubitRegister.h:451-479) all shapes
// are in place: a work item loads its whole 2/4/8-amplitude group before storing it, and groups
// are disjoint.  Bound: HBM.  Algorithmic bytes = 32 B x (amplitudes touched).
//
// Memory access: each thread moves two adjacent amplitudes with one 256-bit LDG/STG whenever
// the lowest fixed bit is >= 1 (or the pair itself is adjacent, target qubit 0); consecutive
// threads cover consecutive 32-byte words, so a warp request is 1 KiB contiguous.
#pragma once

#include "common.cuh"

namespace qcsim {

// ------------------------------------------------------------------------------------------------
// PAIR
// ------------------------------------------------------------------------------------------------
struct PairArgs {
  FixedBits fix;         // bit positions removed from the work-item index (targets + controls)
  uint64_t or_lo, or_hi; // patterns OR-ed into the scattered index (controls set in both)
  amp m00, m01, m10, m11;
  uint64_t n_items;      // 2^(n_local - fix.n)
};

__device__ __forceinline__ void pair_apply(const PairArgs& A, amp a, amp b, amp& oa, amp& ob) {
  oa = cadd(cmul(A.m00, a), cmul(A.m01, b));
  ob = cadd(cmul(A.m10, a), cmul(A.m11, b));
}

// generic 128-bit version: any fixed-bit layout, any size
__global__ void __launch_bounds__(kThreads) k_pair_v1(amp* __restrict__ psi, const __grid_constant__ PairArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < A.n_items; w += stride) {
    const uint64_t base = scatter_index(w, A.fix);
    amp* pa = psi + (base | A.or_lo);
    amp* pb = psi + (base | A.or_hi);
    const amp a = *pa, b = *pb;
    amp oa, ob;
    pair_apply(A, a, b, oa, ob);
    *pa = oa;
    *pb = ob;
  }
}

// all fixed bits >= 1: two adjacent work items per thread, 256-bit accesses, 2x unrolled
__global__ void __launch_bounds__(kThreads) k_pair_v2(amp* __restrict__ psi, const __grid_constant__ PairArgs A) {
  const uint64_t n2 = A.n_items >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; w + stride < n2; w += 2 * stride) {
    const uint64_t b0 = scatter_index(2 * w, A.fix), b1 = scatter_index(2 * (w + stride), A.fix);
    amp* pa0 = psi + (b0 | A.or_lo);
    amp* pb0 = psi + (b0 | A.or_hi);
    amp* pa1 = psi + (b1 | A.or_lo);
    amp* pb1 = psi + (b1 | A.or_hi);
    const amp2 x0 = ld_amp2(pa0), y0 = ld_amp2(pb0), x1 = ld_amp2(pa1), y1 = ld_amp2(pb1);
    amp2 ox, oy;
    pair_apply(A, x0.a, y0.a, ox.a, oy.a);
    pair_apply(A, x0.b, y0.b, ox.b, oy.b);
    st_amp2(pa0, ox);
    st_amp2(pb0, oy);
    pair_apply(A, x1.a, y1.a, ox.a, oy.a);
    pair_apply(A, x1.b, y1.b, ox.b, oy.b);
    st_amp2(pa1, ox);
    st_amp2(pb1, oy);
  }
  if (w < n2) {
    const uint64_t b0 = scatter_index(2 * w, A.fix);
    amp* pa0 = psi + (b0 | A.or_lo);
    amp* pb0 = psi + (b0 | A.or_hi);
    const amp2 x0 = ld_amp2(pa0), y0 = ld_amp2(pb0);
    amp2 ox, oy;
    pair_apply(A, x0.a, y0.a, ox.a, oy.a);
    pair_apply(A, x0.b, y0.b, ox.b, oy.b);
    st_amp2(pa0, ox);
    st_amp2(pb0, oy);
  }
}

// the pair itself is adjacent (or_hi == or_lo + 1, i.e. 2x2 on qubit 0): one 256-bit access per pair
__global__ void __launch_bounds__(kThreads) k_pair_q0(amp* __restrict__ psi, const __grid_constant__ PairArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; w + stride < A.n_items; w += 2 * stride) {
    amp* p0 = psi + (scatter_index(w, A.fix) | A.or_lo);
    amp* p1 = psi + (scatter_index(w + stride, A.fix) | A.or_lo);
    const amp2 x0 = ld_amp2(p0), x1 = ld_amp2(p1);
    amp2 o;
    pair_apply(A, x0.a, x0.b, o.a, o.b);
    st_amp2(p0, o);
    pair_apply(A, x1.a, x1.b, o.a, o.b);
    st_amp2(p1, o);
  }
  if (w < A.n_items) {
    amp* p0 = psi + (scatter_index(w, A.fix) | A.or_lo);
    const amp2 x0 = ld_amp2(p0);
    amp2 o;
    pair_apply(A, x0.a, x0.b, o.a, o.b);
    st_amp2(p0, o);
  }
}

// ------------------------------------------------------------------------------------------------
// DENSE (K = 2 or 3 target qubits)
// ------------------------------------------------------------------------------------------------
template <int K> struct DenseArgs {
  FixedBits fix;             // targets + control, sorted
  uint64_t or_ctrl;          // control bits (set in every index)
  uint64_t off[1 << K];      // off[j] = OR of target bits selected by matrix index j
  amp m[(1 << K) * (1 << K)]; // row-major
  uint64_t n_items;
};

template <int K>
__global__ void __launch_bounds__(kThreads) k_dense_v1(amp* __restrict__ psi, const __grid_constant__ DenseArgs<K> A) {
  constexpr int D = 1 << K;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < A.n_items; w += stride) {
    const uint64_t base = scatter_index(w, A.fix) | A.or_ctrl;
    amp v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = psi[base | A.off[j]];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      amp acc = cmul(A.m[r * D], v[0]);
#pragma unroll
      for (int c = 1; c < D; ++c) acc = cmad(A.m[r * D + c], v[c], acc);
      psi[base | A.off[r]] = acc;
    }
  }
}

// all fixed bits >= 1: two adjacent groups per thread with 256-bit accesses
template <int K>
__global__ void __launch_bounds__(kThreads) k_dense_v2(amp* __restrict__ psi, const __grid_constant__ DenseArgs<K> A) {
  constexpr int D = 1 << K;
  const uint64_t n2 = A.n_items >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n2; w += stride) {
    const uint64_t base = scatter_index(2 * w, A.fix) | A.or_ctrl;
    amp2 v[D];
#pragma unroll
    for (int j = 0; j < D; ++j) v[j] = ld_amp2(psi + (base | A.off[j]));
#pragma unroll
    for (int r = 0; r < D; ++r) {
      amp2 acc;
      acc.a = cmul(A.m[r * D], v[0].a);
      acc.b = cmul(A.m[r * D], v[0].b);
#pragma unroll
      for (int c = 1; c < D; ++c) {
        acc.a = cmad(A.m[r * D + c], v[c].a, acc.a);
        acc.b = cmad(A.m[r * D + c], v[c].b, acc.b);
      }
      st_amp2(psi + (base | A.off[r]), acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// DIAG
// ------------------------------------------------------------------------------------------------
struct DiagArgs {
  FixedBits ctrl;     // control positions (forced to 1), sorted
  uint64_t or_ctrl;
  int nsel;
  int selpos[3];      // selector qubit k contributes bit k of the table index
  amp table[8];
  uint64_t n_items;   // 2^(n_local - ctrl.n)
};

__device__ __forceinline__ int diag_sel(const DiagArgs& A, uint64_t idx) {
  int s = 0;
  if (A.nsel > 0) s |= (int)((idx >> A.selpos[0]) & 1ULL);
  if (A.nsel > 1) s |= (int)((idx >> A.selpos[1]) & 1ULL) << 1;
  if (A.nsel > 2) s |= (int)((idx >> A.selpos[2]) & 1ULL) << 2;
  return s;
}

__global__ void __launch_bounds__(kThreads) k_diag_v1(amp* __restrict__ psi, const __grid_constant__ DiagArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < A.n_items; w += stride) {
    const uint64_t idx = scatter_index(w, A.ctrl) | A.or_ctrl;
    psi[idx] = cmul(psi[idx], A.table[diag_sel(A, idx)]);
  }
}

// all control bits >= 1: 256-bit accesses, 2x unrolled
__global__ void __launch_bounds__(kThreads) k_diag_v2(amp* __restrict__ psi, const __grid_constant__ DiagArgs A) {
  const uint64_t n2 = A.n_items >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; w + stride < n2; w += 2 * stride) {
    const uint64_t i0 = scatter_index(2 * w, A.ctrl) | A.or_ctrl;
    const uint64_t i1 = scatter_index(2 * (w + stride), A.ctrl) | A.or_ctrl;
    amp2 x0 = ld_amp2(psi + i0), x1 = ld_amp2(psi + i1);
    x0.a = cmul(x0.a, A.table[diag_sel(A, i0)]);
    x0.b = cmul(x0.b, A.table[diag_sel(A, i0 | 1)]);
    x1.a = cmul(x1.a, A.table[diag_sel(A, i1)]);
    x1.b = cmul(x1.b, A.table[diag_sel(A, i1 | 1)]);
    st_amp2(psi + i0, x0);
    st_amp2(psi + i1, x1);
  }
  if (w < n2) {
    const uint64_t i0 = scatter_index(2 * w, A.ctrl) | A.or_ctrl;
    amp2 x0 = ld_amp2(psi + i0);
    x0.a = cmul(x0.a, A.table[diag_sel(A, i0)]);
    x0.b = cmul(x0.b, A.table[diag_sel(A, i0 | 1)]);
    st_amp2(psi + i0, x0);
  }
}

// ------------------------------------------------------------------------------------------------
// misc state kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_fill(amp* __restrict__ psi, uint64_t n, amp v) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) psi[i] = v;
}

__global__ void __launch_bounds__(kThreads) k_scale(amp* __restrict__ psi, uint64_t n, double f) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    amp a = psi[i];
    a.x *= f;
    a.y *= f;
    psi[i] = a;
  }
}

// collapse after Measure(first,last): keep (idx & mask) == want scaled by f, zero the rest
// (QubitRegisterCalculator.h:989-995, 1160-1166); `base` = global index of local element 0
__global__ void __launch_bounds__(kThreads)
k_collapse(amp* __restrict__ psi, uint64_t n, uint64_t base, uint64_t mask, uint64_t want, double f) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (((base + i) & mask) == want) {  // kept: read, scale, write
      amp a = psi[i];
      a.x *= f;
      a.y *= f;
      psi[i] = a;
    } else {
      psi[i] = make_amp(0.0, 0.0);  // dropped: written without being read (16 B instead of 32 B per amplitude)
    }
  }
}

}  // namespace qcsim
