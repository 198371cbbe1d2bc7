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
// round there, and stores them back, so shared-memory traffic is one read + one write per ROUND,
// not per gate.
//
// Bound: HBM (32 B x 2^n per pass) up to roughly a dozen dense gates per pass, after which the
// fp64 pipe (64 DFMA/clk/SM) takes over.
#pragma once

#include "common.cuh"

namespace qcsim {

constexpr int kMaxTileBits = 12;  // 2^12 x 16 B = 64 KiB of shared memory per tile
constexpr int kRoundBits = 3;     // amplitudes per thread per round = 2^3
constexpr int kTileThreads = 256;
constexpr int kMaxPassDescBytes = 12 * 1024;

enum TileOpKind : int { TK_PAIR1 = 0, TK_PAIR2 = 1, TK_DENSE2 = 2, TK_DENSE3 = 3, TK_DIAG = 4 };

struct TileOp {        // 64 bytes
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
  int pad[2];
};

// layout of a pass descriptor in device memory: TileRound[n_rounds] | TileOp[n_ops] | amp pool[n_pool]
struct TilePassArgs {
  int k;                        // tile bits
  int n_rounds, n_ops, n_pool;
  int tpos[kMaxTileBits];       // global bit position of tile bit j (ascending, tpos[j] == j for j < L)
  int low_identity;             // L: number of low tile bits that are the low global bits
  uint64_t n_tiles;
  const unsigned char* desc;    // device pointer to the descriptor
  int desc_bytes;
};

// 16-byte slot swizzle: spreads stride-8 (and most other power-of-two stride) accesses over banks
__device__ __forceinline__ uint32_t swz(uint32_t j) { return j ^ ((j >> 3) & 7u); }

template <int RB, int R>
__device__ __forceinline__ void reg_pair1(amp (&v)[1 << RB], const amp* __restrict__ m, uint32_t rctrl) {
  const amp m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int x = 0; x < (1 << RB); ++x) {
    if (x & (1 << R)) continue;
    if ((x & rctrl) == rctrl) {
      const amp a = v[x], b = v[x | (1 << R)];
      v[x] = cadd(cmul(m00, a), cmul(m01, b));
      v[x | (1 << R)] = cadd(cmul(m10, a), cmul(m11, b));
    }
  }
}

template <int RB, int R0, int R1>
__device__ __forceinline__ void reg_pair2(amp (&v)[1 << RB], const amp* __restrict__ m, uint32_t rctrl) {
  const amp m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
  for (int x = 0; x < (1 << RB); ++x) {
    if (x & ((1 << R0) | (1 << R1))) continue;
    if ((x & rctrl) == rctrl) {
      const amp a = v[x | (1 << R0)], b = v[x | (1 << R1)];
      v[x | (1 << R0)] = cadd(cmul(m00, a), cmul(m01, b));
      v[x | (1 << R1)] = cadd(cmul(m10, a), cmul(m11, b));
    }
  }
}

template <int RB, int R0, int R1>
__device__ __forceinline__ void reg_dense2(amp (&v)[1 << RB], const amp* __restrict__ m, uint32_t rctrl) {
#pragma unroll
  for (int x = 0; x < (1 << RB); ++x) {
    if (x & ((1 << R0) | (1 << R1))) continue;
    if ((x & rctrl) == rctrl) {
      const amp a0 = v[x], a1 = v[x | (1 << R0)], a2 = v[x | (1 << R1)], a3 = v[x | (1 << R0) | (1 << R1)];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        amp acc = cmul(m[r * 4], a0);
        acc = cmad(m[r * 4 + 1], a1, acc);
        acc = cmad(m[r * 4 + 2], a2, acc);
        acc = cmad(m[r * 4 + 3], a3, acc);
        v[x | ((r & 1) ? (1 << R0) : 0) | ((r & 2) ? (1 << R1) : 0)] = acc;
      }
    }
  }
}

template <int RB>
__device__ __forceinline__ void reg_dense3(amp (&v)[1 << RB], const amp* __restrict__ m) {
  static_assert(RB == 3, "dense 8x8 uses all three round bits");
  amp a[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = v[c];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    amp acc = cmul(m[r * 8], a[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) acc = cmad(m[r * 8 + c], a[c], acc);
    v[r] = acc;
  }
}

template <int RB>
__device__ __forceinline__ void reg_diag(amp (&v)[1 << RB], const TileOp& op, const amp* __restrict__ table,
                                         uint32_t lbase, uint64_t gbase) {
  int fixed = 0;      // table-index bits that do not depend on the register
  int rpos[3] = {-1, -1, -1};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < op.nsel) {
      if (op.sel_src[k] == 0) rpos[k] = op.sel_pos[k];
      else if (op.sel_src[k] == 1) fixed |= (int)((lbase >> op.sel_pos[k]) & 1u) << k;
      else fixed |= (int)((gbase >> op.sel_pos[k]) & 1ULL) << k;
    }
  }
#pragma unroll
  for (int x = 0; x < (1 << RB); ++x) {
    if ((x & op.rctrl) == op.rctrl) {
      int s = fixed;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (rpos[k] >= 0) s |= ((x >> rpos[k]) & 1) << k;
      v[x] = cmul(v[x], table[s]);
    }
  }
}

template <int RB>
__device__ __forceinline__ void apply_tile_op(amp (&v)[1 << RB], const TileOp& op, const amp* __restrict__ pool,
                                              uint32_t lbase, uint64_t gbase) {
  const amp* m = pool + op.moff;
  switch (op.kind) {
    case TK_PAIR1:
      switch (op.r0) {
        case 0: reg_pair1<RB, 0>(v, m, op.rctrl); break;
        case 1: reg_pair1<RB, 1>(v, m, op.rctrl); break;
        default: reg_pair1<RB, 2>(v, m, op.rctrl); break;
      }
      break;
    case TK_PAIR2:
      if (op.r0 == 0 && op.r1 == 1) reg_pair2<RB, 0, 1>(v, m, op.rctrl);
      else if (op.r0 == 0 && op.r1 == 2) reg_pair2<RB, 0, 2>(v, m, op.rctrl);
      else reg_pair2<RB, 1, 2>(v, m, op.rctrl);
      break;
    case TK_DENSE2:
      if (op.r0 == 0 && op.r1 == 1) reg_dense2<RB, 0, 1>(v, m, op.rctrl);
      else if (op.r0 == 0 && op.r1 == 2) reg_dense2<RB, 0, 2>(v, m, op.rctrl);
      else reg_dense2<RB, 1, 2>(v, m, op.rctrl);
      break;
    case TK_DENSE3: reg_dense3<RB>(v, m); break;
    default: reg_diag<RB>(v, op, m, lbase, gbase); break;
  }
}

// Persistent-style kernel: CTAs stride over tiles.  Dynamic shared memory:
//   [ 2^k amps (tile, swizzled) | pass descriptor copy ]
__global__ void __launch_bounds__(kTileThreads, 2) k_tile_pass(amp* __restrict__ psi, const __grid_constant__ TilePassArgs A) {
  constexpr int RB = kRoundBits;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* tile = reinterpret_cast<amp*>(smem_raw);
  const uint32_t tile_amps = 1u << A.k;
  unsigned char* dsm = smem_raw + (size_t)tile_amps * sizeof(amp);
  {  // descriptor -> shared memory (16-byte words)
    const int4* src = reinterpret_cast<const int4*>(A.desc);
    int4* dst = reinterpret_cast<int4*>(dsm);
    for (int i = threadIdx.x; i < A.desc_bytes / 16; i += blockDim.x) dst[i] = src[i];
  }
  const TileRound* rounds = reinterpret_cast<const TileRound*>(dsm);
  const TileOp* ops = reinterpret_cast<const TileOp*>(dsm + sizeof(TileRound) * A.n_rounds);
  const amp* pool = reinterpret_cast<const amp*>(dsm + sizeof(TileRound) * A.n_rounds + sizeof(TileOp) * A.n_ops);
  const int L = A.low_identity;
  const uint32_t low_mask = (1u << L) - 1u;

  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    // global base of this tile: scatter t into the non-tile bit positions
    uint64_t gbase = t;
#pragma unroll 1
    for (int j = 0; j < A.k; ++j) gbase = insert_zero(gbase, A.tpos[j]);

    // ---- HBM -> shared: each thread moves 2 adjacent amplitudes per 256-bit load
    for (uint32_t j2 = threadIdx.x; j2 < (tile_amps >> 1); j2 += blockDim.x) {
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
    for (int r = 0; r < A.n_rounds; ++r) {
      const TileRound rd = rounds[r];
      const uint32_t items = tile_amps >> RB;
      for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
        uint32_t lbase = it;
        lbase = (uint32_t)insert_zero(lbase, rd.rb[0]);
        lbase = (uint32_t)insert_zero(lbase, rd.rb[1]);
        lbase = (uint32_t)insert_zero(lbase, rd.rb[2]);
        uint32_t off[1 << RB];
#pragma unroll
        for (int x = 0; x < (1 << RB); ++x)
          off[x] = swz(lbase | ((x & 1) ? (1u << rd.rb[0]) : 0u) | ((x & 2) ? (1u << rd.rb[1]) : 0u) |
                       ((x & 4) ? (1u << rd.rb[2]) : 0u));
        amp v[1 << RB];
#pragma unroll
        for (int x = 0; x < (1 << RB); ++x) v[x] = tile[off[x]];
        for (int o = rd.op_begin; o < rd.op_end; ++o) {
          const TileOp& op = ops[o];
          if ((gbase & op.gctrl) != op.gctrl) continue;          // CTA-uniform
          if ((lbase & op.lctrl) != op.lctrl) continue;          // per item
          apply_tile_op<RB>(v, op, pool, lbase, gbase);
        }
#pragma unroll
        for (int x = 0; x < (1 << RB); ++x) tile[off[x]] = v[x];
      }
      __syncthreads();
    }

    // ---- shared -> HBM
    for (uint32_t j2 = threadIdx.x; j2 < (tile_amps >> 1); j2 += blockDim.x) {
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
