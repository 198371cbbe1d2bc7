// tile_kernels.cuh -- fused gate blocks: several gates per HBM pass.
//
// One pass = one kernel.  The pass owns a set T of K "tile qubits" (the low L qubits, so that
// every global access is a run of 2^L contiguous amplitudes, plus K-L arbitrary higher qubits).
// A CTA stages one tile -- the 2^K amplitudes that differ only in the tile qubits -- in shared
// memory (XOR-swizzled, 16 B slots), applies every gate of the pass whose non-diagonal targets
// lie in T, and writes the tile back: 32 B of HBM traffic per amplitude for the whole block of
// gates instead of per gate.  Controls and diagonal selectors may sit on ANY qubit: outside the
// tile they are CTA-uniform (tested against the tile's base index).
//
// Inside a tile the gates are grouped into rounds.  A round names RB (= 3) tile bits; every
// thread pulls the 2^RB amplitudes spanning those bits into registers, applies all gates of the
// round there (branch-free, so the 16 independent DFMA chains of an item interleave), and stores
// them back: shared-memory traffic is one read + one write per ROUND, not per gate.
//
// The whole pass descriptor (rounds, ops, matrices) travels in the kernel parameter block, i.e.
// the constant bank: op decode and matrix entries are uniform constant loads, not LSU traffic.
//
// Bound: HBM (32 B x 2^n per pass) for short blocks; the fp64 pipe (64 DFMA/clk/SM) for long ones.
#pragma once

#include "common.cuh"

namespace qcsim {

constexpr int kMaxTileBits = 12;  // 2^12 x 16 B = 64 KiB of shared memory per tile
constexpr int kRoundBits = 3;     // amplitudes per thread per round = 2^3
constexpr int kTileThreads = 256;
constexpr int kMaxTileRounds = 64;
constexpr int kMaxTileOps = 64;   // also bounds the round count
constexpr int kMaxTilePool = 512;  // amps

enum TileOpKind : int {
  TK_PAIR1 = 0,       // general complex 2x2 on one round bit
  TK_PAIR1_REAL = 1,  // all four entries real (H, Ry)
  TK_PAIR1_RIM = 2,   // real diagonal, imaginary off-diagonal (Rx, SX-like)
  TK_PAIR1_X = 3,     // [[0,1],[1,0]]: X / CNOT / Toffoli core, no arithmetic
  TK_PAIR2 = 4,       // general 2x2 on the (|01>, |10>) pair of two round bits (iSWAP ...)
  TK_PAIR2_SWAP = 5,  // SWAP / Fredkin core
  TK_DENSE2 = 6,
  TK_DENSE3 = 7,
  TK_DIAG = 8,        // table lookup by up to 3 selector qubits
  TK_PHASE = 9,       // DIAG without selectors: one phase on the controlled subspace
};

// host-side view of an op; packed into two int4 words for the parameter block
struct TileOp {
  int kind;
  int r0, r1, r2;      // round-register bits of the targets (ascending; host permutes the matrix to match)
  uint32_t rctrl;      // controls that are round bits (mask over the register index)
  uint32_t lctrl;      // controls on other tile bits (mask over the local tile index)
  uint64_t gctrl;      // controls outside the tile (mask over the global index)
  int nsel;            // DIAG: selector count (table index bit k <- selector k)
  int sel_src[3];      // 0: round bit, 1: local tile bit, 2: global bit
  int sel_pos[3];
  int moff;            // offset (in amps) of the matrix / table in the pool
};

struct TileRound {
  int rb[4];           // local tile bits of this round, ascending (kRoundBits used)
  int op_begin, op_end;
};

struct TilePassArgs {
  int k;                        // tile bits
  int n_rounds;
  int low_identity;             // L: number of low tile bits that are the low global bits
  int pad0;
  uint64_t n_tiles;
  int tpos[kMaxTileBits];       // global bit position of tile bit j (ascending, tpos[j] == j for j < L)
  uint2 rounds[kMaxTileRounds]; // .x = rb0 | rb1<<8 | rb2<<16, .y = op_begin | op_end<<16
  int4 ops[2 * kMaxTileOps];    // see pack_tile_op
  double2 pool[kMaxTilePool];
};

// register-pair variant of a two-target op: (r0, r1) = (0,1) -> 0, (0,2) -> 1, (1,2) -> 2
inline int pair_variant(int r0, int r1) { return r0 == 0 ? (r1 == 1 ? 0 : 1) : 2; }

// word0: x = case id (kind * 4 + variant) | nsel<<8 ; y = ok mask of the round-bit controls | moff<<16 ;
//        z = lctrl ; w = sel (src0|pos0<<2|...)
// word1: x,y = gctrl ; z,w = 0
inline void pack_tile_op(const TileOp& t, int4* w) {
  int variant = 0;
  switch (t.kind) {
    case TK_PAIR1: case TK_PAIR1_REAL: case TK_PAIR1_RIM: case TK_PAIR1_X: variant = t.r0; break;
    case TK_PAIR2: case TK_PAIR2_SWAP: case TK_DENSE2: variant = pair_variant(t.r0, t.r1); break;
    default: break;
  }
  int ok = 0;  // bit x set <=> register x satisfies the round-bit controls
  for (int x = 0; x < 8; ++x)
    if (((uint32_t)x & t.rctrl) == t.rctrl) ok |= 1 << x;
  w[0].x = (t.kind * 4 + variant) | (t.nsel << 8);
  w[0].y = ok | (t.moff << 16);
  w[0].z = (int)t.lctrl;
  int sel = 0;
  for (int i = 0; i < 3; ++i) sel |= ((t.sel_src[i] & 3) | ((t.sel_pos[i] & 63) << 2)) << (8 * i);
  w[0].w = sel;
  w[1].x = (int)(uint32_t)(t.gctrl & 0xffffffffULL);
  w[1].y = (int)(uint32_t)(t.gctrl >> 32);
  w[1].z = 0;
  w[1].w = 0;
}

// 16-byte slot swizzle: spreads stride-8 (and most other power-of-two stride) accesses over banks
__device__ __forceinline__ uint32_t swz(uint32_t j) { return j ^ ((j >> 3) & 7u); }

__device__ __forceinline__ amp sel_amp(bool p, amp a, amp b) { return make_amp(p ? a.x : b.x, p ? a.y : b.y); }

// ---- register-level gate application -------------------------------------------------------------
// A thread holds NI items of 8 amplitudes (the 2^3 combinations of the round bits).  `ok[i]` bit x
// set <=> register x of item i takes part (round-bit controls, other-tile-bit controls and
// outside-tile controls all folded in).  MASKED = false is the straight-line path (every register
// takes part): no selects, and the 16 x NI independent DFMA chains interleave freely.

template <int R, int MODE, int NI, bool MASKED>
__device__ __forceinline__ void reg_pair1(amp (&v)[NI][8], const double2* __restrict__ m, const uint32_t (&ok)[NI]) {
  const amp m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (x & (1 << R)) continue;
      const int y = x | (1 << R);
      const amp a = v[i][x], b = v[i][y];
      amp oa, ob;
      if (MODE == TK_PAIR1_REAL) {
        oa = make_amp(m00.x * a.x + m01.x * b.x, m00.x * a.y + m01.x * b.y);
        ob = make_amp(m10.x * a.x + m11.x * b.x, m10.x * a.y + m11.x * b.y);
      } else if (MODE == TK_PAIR1_RIM) {  // m00, m11 real; m01, m10 imaginary
        oa = make_amp(m00.x * a.x - m01.y * b.y, m00.x * a.y + m01.y * b.x);
        ob = make_amp(m11.x * b.x - m10.y * a.y, m11.x * b.y + m10.y * a.x);
      } else if (MODE == TK_PAIR1_X) {
        oa = b;
        ob = a;
      } else {
        oa = cadd(cmul(m00, a), cmul(m01, b));
        ob = cadd(cmul(m10, a), cmul(m11, b));
      }
      if (MASKED) {
        const bool p = (ok[i] >> x) & 1u;
        v[i][x] = sel_amp(p, oa, a);
        v[i][y] = sel_amp(p, ob, b);
      } else {
        v[i][x] = oa;
        v[i][y] = ob;
      }
    }
  }
}

template <int R0, int R1, bool SWAP, int NI, bool MASKED>
__device__ __forceinline__ void reg_pair2(amp (&v)[NI][8], const double2* __restrict__ m, const uint32_t (&ok)[NI]) {
  const amp m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (x & ((1 << R0) | (1 << R1))) continue;
      const int ia = x | (1 << R0), ib = x | (1 << R1);
      const amp a = v[i][ia], b = v[i][ib];
      amp oa, ob;
      if (SWAP) {
        oa = b;
        ob = a;
      } else {
        oa = cadd(cmul(m00, a), cmul(m01, b));
        ob = cadd(cmul(m10, a), cmul(m11, b));
      }
      if (MASKED) {
        const bool p = (ok[i] >> x) & 1u;
        v[i][ia] = sel_amp(p, oa, a);
        v[i][ib] = sel_amp(p, ob, b);
      } else {
        v[i][ia] = oa;
        v[i][ib] = ob;
      }
    }
  }
}

template <int R0, int R1, int NI>
__device__ __forceinline__ void reg_dense2(amp (&v)[NI][8], const double2* __restrict__ m, const uint32_t (&ok)[NI]) {
#pragma unroll
  for (int i = 0; i < NI; ++i) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      if (x & ((1 << R0) | (1 << R1))) continue;
      const int i1 = x | (1 << R0), i2 = x | (1 << R1), i3 = i1 | i2;
      const amp a0 = v[i][x], a1 = v[i][i1], a2 = v[i][i2], a3 = v[i][i3];
      amp o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        amp acc = cmul(m[r * 4], a0);
        acc = cmad(m[r * 4 + 1], a1, acc);
        acc = cmad(m[r * 4 + 2], a2, acc);
        acc = cmad(m[r * 4 + 3], a3, acc);
        o[r] = acc;
      }
      const bool p = (ok[i] >> x) & 1u;
      v[i][x] = sel_amp(p, o[0], a0);
      v[i][i1] = sel_amp(p, o[1], a1);
      v[i][i2] = sel_amp(p, o[2], a2);
      v[i][i3] = sel_amp(p, o[3], a3);
    }
  }
}

template <int NI>
__device__ __forceinline__ void reg_dense3(amp (&v)[NI][8], const double2* __restrict__ m, const uint32_t (&ok)[NI]) {
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    if (ok[i] == 0) continue;  // dense 8x8 has no round-bit controls: all or nothing
    amp a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = v[i][c];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      amp acc = cmul(m[r * 8], a[0]);
#pragma unroll
      for (int c = 1; c < 8; ++c) acc = cmad(m[r * 8 + c], a[c], acc);
      v[i][r] = acc;
    }
  }
}

#define QCSIM_PAIR1_CASES(MODE)                                                            \
  case MODE * 4 + 0: reg_pair1<0, MODE, NI, MASKED>(v, m, ok); break;                      \
  case MODE * 4 + 1: reg_pair1<1, MODE, NI, MASKED>(v, m, ok); break;                      \
  case MODE * 4 + 2: reg_pair1<2, MODE, NI, MASKED>(v, m, ok); break;

template <int NI, bool MASKED>
__device__ __forceinline__ void apply_pair_op(int case_id, amp (&v)[NI][8], const double2* __restrict__ m,
                                              const uint32_t (&ok)[NI]) {
  switch (case_id) {
    QCSIM_PAIR1_CASES(TK_PAIR1)
    QCSIM_PAIR1_CASES(TK_PAIR1_REAL)
    QCSIM_PAIR1_CASES(TK_PAIR1_RIM)
    QCSIM_PAIR1_CASES(TK_PAIR1_X)
    case TK_PAIR2 * 4 + 0: reg_pair2<0, 1, false, NI, MASKED>(v, m, ok); break;
    case TK_PAIR2 * 4 + 1: reg_pair2<0, 2, false, NI, MASKED>(v, m, ok); break;
    case TK_PAIR2 * 4 + 2: reg_pair2<1, 2, false, NI, MASKED>(v, m, ok); break;
    case TK_PAIR2_SWAP * 4 + 0: reg_pair2<0, 1, true, NI, MASKED>(v, m, ok); break;
    case TK_PAIR2_SWAP * 4 + 1: reg_pair2<0, 2, true, NI, MASKED>(v, m, ok); break;
    case TK_PAIR2_SWAP * 4 + 2: reg_pair2<1, 2, true, NI, MASKED>(v, m, ok); break;
    default: break;
  }
}

// Dynamic shared memory: 2^k amps (tile, swizzled).  NI = items per thread per round
// (k = 12: 2, k <= 11: 1).
template <int NI, int MINB>
__global__ void __launch_bounds__(kTileThreads, MINB) k_tile_pass(amp* __restrict__ psi, const __grid_constant__ TilePassArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* tile = reinterpret_cast<amp*>(smem_raw);
  const uint32_t tile_amps = 1u << A.k;
  const int L = A.low_identity;
  const uint32_t low_mask = (1u << L) - 1u;
  const uint32_t items = tile_amps >> 3;

  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    // global base of this tile: scatter t into the non-tile bit positions
    uint64_t gbase = t;
#pragma unroll 1
    for (int j = 0; j < A.k; ++j) gbase = insert_zero(gbase, A.tpos[j]);

    // ---- HBM -> shared: each thread moves 2 adjacent amplitudes per 256-bit load
    for (uint32_t j2 = threadIdx.x; j2 < (tile_amps >> 1); j2 += kTileThreads) {
      const uint32_t loc = j2 << 1;
      uint64_t g = gbase | (loc & low_mask);
#pragma unroll 1
      for (int j = L; j < A.k; ++j) g |= (uint64_t)((loc >> j) & 1u) << A.tpos[j];
      const amp2 x = ld_amp2(psi + g);
      tile[swz(loc)] = x.a;
      tile[swz(loc + 1)] = x.b;
    }
    __syncthreads();

    // ---- rounds
#pragma unroll 1
    for (int r = 0; r < A.n_rounds; ++r) {
      const uint2 rd = A.rounds[r];
      const int rb0 = rd.x & 0xff, rb1 = (rd.x >> 8) & 0xff, rb2 = (rd.x >> 16) & 0xff;
      const int op_begin = rd.y & 0xffff, op_end = rd.y >> 16;
#pragma unroll 1
      for (uint32_t it0 = 0; it0 < items; it0 += NI * kTileThreads) {
      uint32_t lbase[NI];
      bool live[NI];
      amp v[NI][8];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const uint32_t it = it0 + threadIdx.x + i * kTileThreads;
        live[i] = it < items;
        uint32_t b = live[i] ? it : 0u;
        b = (uint32_t)insert_zero(b, rb0);
        b = (uint32_t)insert_zero(b, rb1);
        b = (uint32_t)insert_zero(b, rb2);
        lbase[i] = b;
#pragma unroll
        for (int x = 0; x < 8; ++x)
          v[i][x] = tile[swz(b | ((x & 1) ? (1u << rb0) : 0u) | ((x & 2) ? (1u << rb1) : 0u) | ((x & 4) ? (1u << rb2) : 0u))];
      }
#pragma unroll 1
      for (int o = op_begin; o < op_end; ++o) {
        const int4 w0 = A.ops[2 * o];
        const int4 w1 = A.ops[2 * o + 1];
        const uint64_t gctrl = (uint64_t)(uint32_t)w1.x | ((uint64_t)(uint32_t)w1.y << 32);
        if ((gbase & gctrl) != gctrl) continue;  // outside-tile controls: CTA-uniform
        const uint32_t lctrl = (uint32_t)w0.z;
        const uint32_t okr = (uint32_t)w0.y & 0xffu;
        uint32_t ok[NI];
        bool all_full = true;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          ok[i] = ((lbase[i] & lctrl) == lctrl) ? okr : 0u;
          all_full = all_full && (ok[i] == 0xffu);
        }
        const int case_id = w0.x & 0xff;
        const double2* m = A.pool + ((uint32_t)w0.y >> 16);
        if (case_id < TK_DENSE2 * 4) {
          if (all_full) apply_pair_op<NI, false>(case_id, v, m, ok);
          else apply_pair_op<NI, true>(case_id, v, m, ok);
        } else if (case_id < TK_DENSE3 * 4) {
          const int var = case_id & 3;
          if (var == 0) reg_dense2<0, 1, NI>(v, m, ok);
          else if (var == 1) reg_dense2<0, 2, NI>(v, m, ok);
          else reg_dense2<1, 2, NI>(v, m, ok);
        } else if (case_id < TK_DIAG * 4) {
          reg_dense3<NI>(v, m, ok);
        } else if (case_id >= TK_PHASE * 4) {
          const amp ph = m[0];
#pragma unroll
          for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int x = 0; x < 8; ++x) v[i][x] = sel_amp((ok[i] >> x) & 1u, cmul(v[i][x], ph), v[i][x]);
        } else {  // TK_DIAG
          const int nsel = (w0.x >> 8) & 0xff;
          if (nsel == 1) {  // Rz / CRz: both table entries are uniform constant loads, selected per amplitude
            const amp t0 = m[0], t1 = m[1];
            const int s0 = w0.w & 0xff;
            const int src = s0 & 3, pos = s0 >> 2;
            const int rm = (src == 0) ? (1 << pos) : 0;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const bool fx = (src == 1) ? ((lbase[i] >> pos) & 1u) : (src == 2) ? ((gbase >> pos) & 1ULL) : false;
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                const bool hi = fx || ((x & rm) != 0);
                const amp nv = cmul(v[i][x], sel_amp(hi, t1, t0));
                v[i][x] = all_full ? nv : sel_amp((ok[i] >> x) & 1u, nv, v[i][x]);
              }
            }
          } else {
            int rmask[3] = {0, 0, 0};
            int src3[3], pos3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const int sk = (w0.w >> (8 * k)) & 0xff;
              src3[k] = (k < nsel) ? (sk & 3) : 3;
              pos3[k] = sk >> 2;
              if (src3[k] == 0) rmask[k] = 1 << pos3[k];
            }
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              int fixed = 0;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                if (src3[k] == 1) fixed |= (int)((lbase[i] >> pos3[k]) & 1u) << k;
                else if (src3[k] == 2) fixed |= (int)((gbase >> pos3[k]) & 1ULL) << k;
              }
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                const int sidx = fixed | ((x & rmask[0]) ? 1 : 0) | ((x & rmask[1]) ? 2 : 0) | ((x & rmask[2]) ? 4 : 0);
                v[i][x] = sel_amp((ok[i] >> x) & 1u, cmul(v[i][x], m[sidx]), v[i][x]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (!live[i]) continue;
        const uint32_t b = lbase[i];
#pragma unroll
        for (int x = 0; x < 8; ++x)
          tile[swz(b | ((x & 1) ? (1u << rb0) : 0u) | ((x & 2) ? (1u << rb1) : 0u) | ((x & 4) ? (1u << rb2) : 0u))] = v[i][x];
      }
      }  // item groups
      __syncthreads();
    }

    // ---- shared -> HBM
    for (uint32_t j2 = threadIdx.x; j2 < (tile_amps >> 1); j2 += kTileThreads) {
      const uint32_t loc = j2 << 1;
      uint64_t g = gbase | (loc & low_mask);
#pragma unroll 1
      for (int j = L; j < A.k; ++j) g |= (uint64_t)((loc >> j) & 1u) << A.tpos[j];
      amp2 x;
      x.a = tile[swz(loc)];
      x.b = tile[swz(loc + 1)];
      st_amp2(psi + g, x);
    }
    __syncthreads();
  }
}

}  // namespace qcsim
