// qft_kernels.cuh -- the quantum Fourier transform as radix-8 FFT passes, and bit-permutation passes.
//
// QuantumFourierTransform::QFT (QuantumFourierTransform.h:35-60) is, for every qubit `cur` from the
// top down:  H(cur), then a controlled phase pi/2^(cur-ctrl) from every lower qubit ctrl.  The
// controlled phases are diagonal, so for a GROUP of up to 3 adjacent qubits c2 > c1 > c0 all gates
// with their target in the group reduce to
//     [ H(c2) CP(c2,c1) CP(c2,c0) H(c1) CP(c1,c0) H(c0) ]   (a radix-8 butterfly on 8 amplitudes)
//   x  amp(x2 x1 x0) *= P^(x2 + 2 x1 + 4 x0),   P = exp(i pi R / 2^c2),
// where R is the value of ALL lower qubits of the amplitude's index (bits [sq, c0)): one sincospi
// and 13 complex multiplies per 8 amplitudes instead of up to 3 x 30 controlled-phase passes.
// This is exact algebra on the reference's gate sequence -- the same matrices (1/sqrt 2,
// std::polar(1, pi/2), std::polar(1, pi/4) come from the host), only the diagonal factors are
// multiplied together before they meet the amplitude; results agree with the gate-by-gate
// reference to ~1e-15 (tests: 1e-12).
//
// One pass = one kernel over tiles, as in tile_kernels.cuh: tile = the low L index bits (contiguous
// 2^L-amplitude runs in HBM) + the bits of up to 4 groups; each group is one round on the
// shared-memory tile.  A 30-qubit QFT is 3 passes (10 rounds) instead of 465 gate passes.
// IQFT (:62-87) is the adjoint: groups bottom-up, conjugate twiddle first, inverse butterfly after.
// Bound: HBM (32 B per amplitude per pass; ~25 DFMA-class instructions per amplitude per round).
#pragma once

#include "common.cuh"

namespace qcsim {

constexpr int kMaxQftGroups = 4;

struct QftGroup {
  int size;         // 1..3 qubits
  int top_qubit;    // logical index of the group's highest qubit (c2)
  uint32_t rb;      // tile-local bit of the group's qubits, lowest logical first (byte each)
  int pad;
  uint32_t tb[3];   // tile bit walked by item-index bit j (byte each), for the k - size other bits
  uint32_t pad2;
};

struct QftPassArgs {
  int k;                 // tile bits
  int low_identity;      // number of low tile bits that are the low index bits
  int n_groups;
  int inverse;
  uint64_t n_tiles;
  uint64_t rank_bits;    // physical index bits above the local slice (rank << n_local), 0 on one GPU
  int tpos[kMaxTileBits];
  int sq;                // lowest (logical) qubit of the transform: bits below it never enter R
  int n_phys;            // physical index bits (local + rank bits)
  signed char log_of[64];  // logical qubit held by physical index bit p: any layout (sharded registers keep qubit
                           // reversals and exchanges virtual, dist.cu) -- R is gathered bit by bit
  double s;              // 1/sqrt(2), HadamardGate (SimpleGates.h:588-596)
  double2 ph2, ph4;      // std::polar(1., +-pi/2), std::polar(1., +-pi/4) (QuantumGate.h:262-265), sign by direction
  QftGroup groups[kMaxQftGroups];
};

__device__ __forceinline__ void hadamard_pair(amp& a, amp& b, double s) {
  const amp x = a, y = b;
  a = make_amp(s * (x.x + y.x), s * (x.y + y.y));
  b = make_amp(s * (x.x - y.x), s * (x.y - y.y));
}

template <int G, class Args>
__device__ __forceinline__ void qft_group(amp (&v)[8], const Args& A, bool inverse, amp P) {
  constexpr int N = 1 << G;
  // amp x gets P^rev(x) = prod over set bits b of x of P^(2^(G-1-b)): the top register bit takes P,
  // the next one P^2, the lowest P^4 -- one power of P live at a time (register pressure)
  auto twiddle = [&]() {
    amp pw = P;
#pragma unroll
    for (int b = G - 1; b >= 0; --b) {
#pragma unroll
      for (int x = 0; x < N; ++x)
        if ((x >> b) & 1) v[x] = cmul(v[x], pw);
      if (b > 0) pw = cmul(pw, pw);
    }
  };
  auto hadamard = [&](int bit) {
#pragma unroll
    for (int x = 0; x < N; ++x)
      if (!((x >> bit) & 1)) hadamard_pair(v[x], v[x | (1 << bit)], A.s);
  };
  auto cphase = [&](int tbit, int cbit, amp ph) {
#pragma unroll
    for (int x = 0; x < N; ++x)
      if (((x >> tbit) & 1) && ((x >> cbit) & 1)) v[x] = cmul(v[x], ph);
  };
  if (!inverse) {
    if (G == 3) {
      hadamard(2);
      cphase(2, 1, A.ph2);
      cphase(2, 0, A.ph4);
      hadamard(1);
      cphase(1, 0, A.ph2);
      hadamard(0);
    } else if (G == 2) {
      hadamard(1);
      cphase(1, 0, A.ph2);
      hadamard(0);
    } else {
      hadamard(0);
    }
    twiddle();
  } else {
    twiddle();
    if (G == 3) {
      hadamard(0);
      cphase(1, 0, A.ph2);
      hadamard(1);
      cphase(2, 1, A.ph2);
      cphase(2, 0, A.ph4);
      hadamard(2);
    } else if (G == 2) {
      hadamard(0);
      cphase(1, 0, A.ph2);
      hadamard(1);
    } else {
      hadamard(0);
    }
  }
}

// Shared memory: tile (2^k amps) | per-tile CTA-part twiddles (kMaxQftGroups amps).  The thread part
// of the twiddles of 3-qubit groups in 12-bit tiles depends on the item only: it is tabulated once
// per pass in global memory (k_qft_item_table, 32 KiB, L1-resident) so that the tile kernel fits
// three CTAs per SM.
constexpr int kQftItemsMax = 2048;  // items of a 1-qubit group in a 12-bit tile (512 for a 3-qubit group)

template <class Args>
__device__ __forceinline__ uint64_t qft_logical(const Args& A, uint64_t phys) {
  uint64_t logical = 0;
#pragma unroll 1
  for (int p = 0; p < A.n_phys; ++p) logical |= ((phys >> p) & 1ULL) << A.log_of[p];
  return logical;
}

__device__ __forceinline__ uint32_t qft_lbase(const QftGroup& grp, uint32_t item) {
  uint32_t lbase = 0;
#pragma unroll
  for (int j = 0; j < 11; ++j) lbase |= ((item >> j) & 1u) << ((grp.tb[j >> 2] >> (8 * (j & 3))) & 31u);
  return lbase;
}

// twiddle base exp(+-i pi R / 2^top) for the lower-qubit value R (masked to [sq, c0))
template <class Args, class Group>
__device__ __forceinline__ amp qft_base(const Args& A, const Group& grp, uint64_t logical, bool inverse) {
  const int c0 = grp.top_qubit - grp.size + 1;
  const uint64_t r_mask = ((1ULL << c0) - 1ULL) & ~((1ULL << A.sq) - 1ULL);
  const uint64_t R = logical & r_mask;
  double sn, cs;
  sincospi((double)R * exp2(-(double)grp.top_qubit), &sn, &cs);  // exact argument: R < 2^53, scale a power of two
  return make_amp(cs, inverse ? -sn : sn);
}

__device__ __forceinline__ uint64_t qft_phys_of_local(const QftPassArgs& A, uint32_t lbase) {
  const uint32_t low_mask = (1u << A.low_identity) - 1u;
  uint64_t phys = lbase & low_mask;
#pragma unroll 1
  for (int j = A.low_identity; j < A.k; ++j) phys |= (uint64_t)((lbase >> j) & 1u) << A.tpos[j];
  return phys;
}

// table[g * kQftItemsMax + item] = thread part of the twiddle base of item `item` of group g (12-bit tiles);
// slots[same index] = swizzled shared-memory slot of the item's first amplitude (the 11-step bit deposit of
// qft_lbase, done once per pass instead of once per item per tile)
static __global__ void k_qft_item_table(amp* __restrict__ table, uint32_t* __restrict__ slots, const __grid_constant__ QftPassArgs A) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int gi = i / kQftItemsMax;
  if (gi >= A.n_groups) return;
  const QftGroup grp = A.groups[gi];
  const uint32_t item = i % kQftItemsMax;
  if (item >= (1u << (A.k - grp.size))) return;
  const uint32_t lbase = qft_lbase(grp, item);
  table[i] = qft_base(A, grp, qft_logical(A, qft_phys_of_local(A, lbase)), A.inverse != 0);
  slots[i] = swz(lbase);
}

static __global__ void __launch_bounds__(kTileThreads, 3) k_qft_pass(amp* __restrict__ psi, const amp* __restrict__ tw_item,
                                                                      const uint32_t* __restrict__ slot_item,
                                                                      const __grid_constant__ QftPassArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* tile = reinterpret_cast<amp*>(smem_raw);
  const int k = A.k;
  const uint32_t tile_amps = 1u << k;
  amp* tw_cta = tile + tile_amps;  // [group]
  const int L = A.low_identity;
  const uint32_t low_mask = (1u << L) - 1u;
  const uint32_t tid = threadIdx.x;
  const bool inverse = A.inverse != 0;
  const bool tabled = (k == 12);

  const uint32_t loc_fixed = (tid << 1) & (tile_amps - 1u);
  uint64_t g_fixed = loc_fixed & low_mask;
#pragma unroll 1
  for (int j = L; j < k; ++j) g_fixed |= (uint64_t)((loc_fixed >> j) & 1u) << A.tpos[j];
  const uint32_t s_fixed = swz(loc_fixed);
  const uint32_t n_it = (tile_amps >> 1) > kTileThreads ? (tile_amps >> 1) / kTileThreads : 1u;
  const bool mover = (tid << 1) < tile_amps;

  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    uint64_t gbase = t;
#pragma unroll 1
    for (int j = 0; j < k; ++j) gbase = insert_zero(gbase, A.tpos[j]);
    // CTA-part twiddles of this tile (bits outside the tile + rank bits)
    if (tid < (uint32_t)A.n_groups) tw_cta[tid] = qft_base(A, A.groups[tid], qft_logical(A, gbase | A.rank_bits), inverse);

    if (mover) {
      if (n_it == 8) {
        // 12-bit tile: all eight 32-byte loads of the thread in flight at once, straight-line (a partially
        // predicated register array ended up in local memory: the load phase was a third of the kernel)
        const uint64_t g9 = 1ULL << A.tpos[9], g10 = 1ULL << A.tpos[10], g11 = 1ULL << A.tpos[11];
        const amp* src = psi + (gbase | g_fixed);
        const amp2 x0 = ld_amp2(src), x1 = ld_amp2(src + g9), x2 = ld_amp2(src + g10), x3 = ld_amp2(src + (g9 | g10));
        const amp2 x4 = ld_amp2(src + g11), x5 = ld_amp2(src + (g11 | g9)), x6 = ld_amp2(src + (g11 | g10)), x7 = ld_amp2(src + (g11 | g10 | g9));
#define QCSIM_QFT_PUT(U, X)                           \
  {                                                   \
    const uint32_t s_ = s_fixed ^ swz((uint32_t)(U) << 9); \
    tile[s_] = (X).a;                                 \
    tile[s_ ^ 1u] = (X).b;                            \
  }
        QCSIM_QFT_PUT(0, x0) QCSIM_QFT_PUT(1, x1) QCSIM_QFT_PUT(2, x2) QCSIM_QFT_PUT(3, x3)
        QCSIM_QFT_PUT(4, x4) QCSIM_QFT_PUT(5, x5) QCSIM_QFT_PUT(6, x6) QCSIM_QFT_PUT(7, x7)
#undef QCSIM_QFT_PUT
      } else {
#pragma unroll 1
        for (uint32_t it = 0; it < n_it; ++it) {
          const uint32_t lv = it << 9;
          uint64_t gv = 0;
#pragma unroll 1
          for (int j = 9; j < k; ++j) gv |= (uint64_t)((lv >> j) & 1u) << A.tpos[j];
          const amp2 x = ld_amp2(psi + (gbase | g_fixed | gv));
          const uint32_t s = s_fixed ^ swz(lv);
          tile[s] = x.a;
          tile[s ^ 1u] = x.b;
        }
      }
    }
    __syncthreads();

#pragma unroll 1
    for (int gi = 0; gi < A.n_groups; ++gi) {
      const QftGroup grp = A.groups[gi];
      const int G = grp.size;
      const uint32_t items = tile_amps >> G;
      const uint32_t so0 = swz(1u << (grp.rb & 31u)), so1 = swz(1u << ((grp.rb >> 8) & 31u)), so2 = swz(1u << ((grp.rb >> 16) & 31u));
      const amp p_cta = tw_cta[gi];
#pragma unroll 1
      for (uint32_t item = tid; item < items; item += kTileThreads) {
        amp P;
        uint32_t sl;
        if (tabled) {
          P = cmul(p_cta, __ldg(tw_item + gi * kQftItemsMax + item));
          sl = __ldg(slot_item + gi * kQftItemsMax + item);
        } else {
          const uint32_t lbase = qft_lbase(grp, item);
          P = qft_base(A, grp, qft_logical(A, gbase | A.rank_bits | qft_phys_of_local(A, lbase)), inverse);
          sl = swz(lbase);
        }
        amp v[8];
        if (G == 3) {
#pragma unroll
          for (int x = 0; x < 8; ++x) v[x] = tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u)];
          qft_group<3>(v, A, inverse, P);
#pragma unroll
          for (int x = 0; x < 8; ++x) tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u)] = v[x];
        } else if (G == 2) {
#pragma unroll
          for (int x = 0; x < 4; ++x) v[x] = tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u)];
          qft_group<2>(v, A, inverse, P);
#pragma unroll
          for (int x = 0; x < 4; ++x) tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u)] = v[x];
        } else {
          v[0] = tile[sl];
          v[1] = tile[sl ^ so0];
          qft_group<1>(v, A, inverse, P);
          tile[sl] = v[0];
          tile[sl ^ so0] = v[1];
        }
      }
      __syncthreads();
    }

    if (mover) {
      if (n_it == 8) {
        const uint64_t g9 = 1ULL << A.tpos[9], g10 = 1ULL << A.tpos[10], g11 = 1ULL << A.tpos[11];
        amp* dst = psi + (gbase | g_fixed);
#pragma unroll
        for (uint32_t it = 0; it < 8; ++it) {
          const uint32_t s = s_fixed ^ swz(it << 9);
          amp2 x;
          x.a = tile[s];
          x.b = tile[s ^ 1u];
          st_amp2(dst + (((it & 1u) ? g9 : 0ULL) | ((it & 2u) ? g10 : 0ULL) | ((it & 4u) ? g11 : 0ULL)), x);
        }
      } else {
#pragma unroll 1
        for (uint32_t it = 0; it < n_it; ++it) {
          const uint32_t lv = it << 9;
          uint64_t gv = 0;
#pragma unroll 1
          for (int j = 9; j < k; ++j) gv |= (uint64_t)((lv >> j) & 1u) << A.tpos[j];
          const uint32_t s = s_fixed ^ swz(lv);
          amp2 x;
          x.a = tile[s];
          x.b = tile[s ^ 1u];
          st_amp2(psi + (gbase | g_fixed | gv), x);
        }
      }
    }
    __syncthreads();
  }
}

// ---- in-tile bit permutation: one HBM pass for any permutation of <= 12 index bits ------------------
// Used for the QFT's qubit reversal (QubitsSwapper.h:23-34: SWAP(s,e), SWAP(s+1,e-1) ...) and for
// putting a sharded register back into canonical qubit order.  The tile is loaded as in the gate
// passes; on the way out, output slot j receives the amplitude from slot src(j), where src permutes
// the bits of j.
struct PermPassArgs {
  int k;
  int low_identity;
  uint64_t n_tiles;
  int tpos[kMaxTileBits];
  int src_bit[kMaxTileBits];  // output local bit j is taken from input local bit src_bit[j]
};

static __global__ void __launch_bounds__(kTileThreads, 3) k_tile_permute(amp* __restrict__ psi, const __grid_constant__ PermPassArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* tile = reinterpret_cast<amp*>(smem_raw);
  const int k = A.k;
  const uint32_t tile_amps = 1u << k;
  const int L = A.low_identity;
  const uint32_t low_mask = (1u << L) - 1u;
  const uint32_t tid = threadIdx.x;
  const uint32_t loc_fixed = (tid << 1) & (tile_amps - 1u);
  uint64_t g_fixed = loc_fixed & low_mask;
#pragma unroll 1
  for (int j = L; j < k; ++j) g_fixed |= (uint64_t)((loc_fixed >> j) & 1u) << A.tpos[j];
  const uint32_t s_fixed = swz(loc_fixed);
  const uint32_t n_it = (tile_amps >> 1) > kTileThreads ? (tile_amps >> 1) / kTileThreads : 1u;
  const bool mover = (tid << 1) < tile_amps;
  // input slots of this thread's two output amplitudes per iteration: src() is linear over XOR
  auto src = [&](uint32_t j) {
    uint32_t o = 0;
#pragma unroll 1
    for (int b = 0; b < k; ++b) o |= ((j >> b) & 1u) << A.src_bit[b];
    return o;
  };
  const uint32_t src_fixed = swz(src(loc_fixed));
  const uint32_t src_one = swz(src(1u));

  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    uint64_t gbase = t;
#pragma unroll 1
    for (int j = 0; j < k; ++j) gbase = insert_zero(gbase, A.tpos[j]);
    if (mover) {
#pragma unroll 1
      for (uint32_t it0 = 0; it0 < n_it; it0 += 8) {
        amp2 x[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) {
          const uint32_t lv = (it0 + u) << 9;
          uint64_t gv = 0;
#pragma unroll 1
          for (int j = 9; j < k; ++j) gv |= (uint64_t)((lv >> j) & 1u) << A.tpos[j];
          if (it0 + u < n_it) x[u] = ld_amp2(psi + (gbase | g_fixed | gv));
        }
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u) {
          if (it0 + u >= n_it) continue;
          const uint32_t s = s_fixed ^ swz((it0 + u) << 9);
          tile[s] = x[u].a;
          tile[s ^ 1u] = x[u].b;
        }
      }
    }
    __syncthreads();
    if (mover) {
#pragma unroll 1
      for (uint32_t it = 0; it < n_it; ++it) {
        const uint32_t lv = it << 9;
        uint64_t gv = 0;
#pragma unroll 1
        for (int j = 9; j < k; ++j) gv |= (uint64_t)((lv >> j) & 1u) << A.tpos[j];
        const uint32_t s = src_fixed ^ swz(src(lv));
        amp2 x;
        x.a = tile[s];
        x.b = tile[s ^ src_one];
        st_amp2(psi + (gbase | g_fixed | gv), x);
      }
    }
    __syncthreads();
  }
}

// ---- full qubit reversal in ONE pass (COBRA-style) -----------------------------------------------------
// QubitsSwapper::Swap on qubits [0, m) (QubitsSwapper.h:23-34) maps index bits (a : u : c) -- top T,
// middle m-2T, low T -- to (rev c : rev u : rev a).  A CTA takes the two 2^T x 2^T tiles with middle
// values u and rev(u) (2^T runs of 2^T amplitudes each), keeps both in shared memory, and writes each
// to the other's place with the (a, c) roles exchanged and bit-reversed: every amplitude is read
// once and written once, all global accesses are 2^T x 16 B runs (T = 5: 512 B).  Shared-memory slot
// of element (a, c): row revT(c), column a ^ (a >> 3) ^ (row >> (T-4) & 7), which keeps both the
// transposing stores and the row-wise loads free of bank conflicts.  Small tiles (32 KiB per CTA)
// so that several CTAs per SM overlap their load and store phases.
constexpr int kRevT = 5;

struct BitRevArgs {
  int m;            // qubits [0, m) are reversed, m >= 2 kRevT
  int n_local;
  uint64_t n_work;  // 2^(n_local - 2 kRevT): (outer bits above m) x (middle value)
};

template <int T>
__device__ __forceinline__ uint32_t rev_t(uint32_t x) { return __brev(x) >> (32 - T); }
template <int T>
__device__ __forceinline__ uint32_t cobra_col(uint32_t col, uint32_t row) {
  return (col ^ (col >> 3) ^ ((row >> (T - 4)) & 7u)) & ((1u << T) - 1u);
}

static __global__ void __launch_bounds__(kTileThreads, 4) k_bit_reverse(amp* __restrict__ psi, const __grid_constant__ BitRevArgs A) {
  constexpr int T = kRevT;
  constexpr uint32_t W = 1u << T;                      // tile is W x W
  constexpr uint32_t ROWS_PER_IT = 2 * kTileThreads / W;  // each thread moves 2 adjacent amplitudes
  constexpr uint32_t ITS = W / ROWS_PER_IT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* bufA = reinterpret_cast<amp*>(smem_raw);
  amp* bufB = bufA + W * W;
  const int mid_bits = A.m - 2 * T;
  const uint32_t tid = threadIdx.x;
  const uint32_t c2 = (tid % (W / 2)) * 2u, a_lo = tid / (W / 2);
  const int hi = A.m - T;  // position of the `a` field
  for (uint64_t w = blockIdx.x; w < A.n_work; w += gridDim.x) {
    const uint64_t u = mid_bits ? (w & ((1ULL << mid_bits) - 1ULL)) : 0ULL;
    const uint64_t o = w >> mid_bits;
    const uint64_t ur = mid_bits ? (__brevll(u) >> (64 - mid_bits)) : 0ULL;
    if (u > ur) continue;  // the pair is handled by the CTA that drew the smaller middle value
    const uint64_t base_u = (o << A.m) | (u << T), base_r = (o << A.m) | (ur << T);
    amp2 x[ITS], y[ITS];
#pragma unroll
    for (uint32_t it = 0; it < ITS; ++it) x[it] = ld_amp2(psi + (base_u | ((uint64_t)(it * ROWS_PER_IT + a_lo) << hi) | c2));
    if (u != ur) {
#pragma unroll
      for (uint32_t it = 0; it < ITS; ++it) y[it] = ld_amp2(psi + (base_r | ((uint64_t)(it * ROWS_PER_IT + a_lo) << hi) | c2));
    }
    // element (a, c) -> row revT(c), swizzled column a
    const uint32_t r0 = rev_t<T>(c2), r1 = r0 | (W >> 1);
#pragma unroll
    for (uint32_t it = 0; it < ITS; ++it) {
      const uint32_t a = it * ROWS_PER_IT + a_lo;
      const uint32_t s0 = r0 * W + cobra_col<T>(a, r0), s1 = r1 * W + cobra_col<T>(a, r1);
      bufA[s0] = x[it].a;
      bufA[s1] = x[it].b;
      if (u != ur) {
        bufB[s0] = y[it].a;
        bufB[s1] = y[it].b;
      }
    }
    __syncthreads();
    // output element (a', c') of the tile at rev(u) is input element (revT c', revT a') of the tile at u:
    // row = a', column from a = revT(c')
#pragma unroll
    for (uint32_t it = 0; it < ITS; ++it) {
      const uint32_t ap = it * ROWS_PER_IT + a_lo;
      const uint32_t s0 = ap * W + cobra_col<T>(r0, ap), s1 = ap * W + cobra_col<T>(r1, ap);
      amp2 v;
      v.a = bufA[s0];
      v.b = bufA[s1];
      st_amp2(psi + (base_r | ((uint64_t)ap << hi) | c2), v);
      if (u != ur) {
        v.a = bufB[s0];
        v.b = bufB[s1];
        st_amp2(psi + (base_u | ((uint64_t)ap << hi) | c2), v);
      }
    }
    __syncthreads();
  }
}

}  // namespace qcsim
