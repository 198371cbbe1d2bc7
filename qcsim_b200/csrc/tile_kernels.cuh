// tile_kernels.cuh -- fused gate blocks: several gates per HBM pass.
//
// One pass = one kernel.  The pass owns a set T of K "tile qubits" (the low L qubits, so that
// every global access is a run of 2^L contiguous amplitudes, plus K-L arbitrary higher qubits).
// A CTA stages one tile -- the 2^K amplitudes that differ only in the tile qubits -- in shared
// memory (XOR-swizzled 16 B slots), applies every gate of the pass whose non-diagonal targets
// lie in T, and writes the tile back: 32 B of HBM traffic per amplitude for the whole block of
// gates instead of per gate.
//
// Inside a tile the gates are grouped into ROUNDS (planner.h: schedule_rounds).  A round names 3
// tile bits; the host multiplies every gate of the round into ONE 8x8 complex matrix over those
// bits, and each thread applies it to the 8 amplitudes that span them: one shared-memory read and
// one write per round, 256 DFMA-class instructions per 8 amplitudes, no per-gate decode and no
// data-dependent control flow -- the kernel is the same straight-line mat-vec whatever the gates
// were (X / CNOT / SWAP / diagonal gates cost nothing extra once folded into the matrix).
// Qubits a gate only looks at (controls, diagonal selectors) need not be round bits: up to 3 such
// "variant" bits per round select one of 2^v precomputed matrices.  A variant bit is either outside
// the tile (CTA-uniform) or a tile bit that the host maps onto the warp-index part of the item
// index, so the choice is warp-uniform: the matrices travel in the kernel parameter block (constant
// bank) and their entries reach the DFMAs through uniform registers -- no LSU traffic at all.
//
// Bound: HBM (32 B x 2^n per pass) up to ~3 rounds per pass; beyond that the fp64 pipe
// (64 DFMA/clk/SM: 32 DFMA per amplitude per round).
#pragma once

#include "common.cuh"

namespace qcsim {

constexpr int kRoundBits = 3;      // amplitudes per thread per round = 2^3
constexpr int kMaxVariantBits = 3; // matrices per round <= 2^3
constexpr int kMaxTileRounds = 7;  // per launch: 7 rounds x 4 matrices x 1 KiB fits the 32 KiB parameter block
constexpr int kMaxTileMats = 28;
constexpr int kRoundMatAmps = 64;  // one 8x8 complex matrix

struct TileRoundDesc {
  uint32_t rb;       // rb0 | rb1 << 8 | rb2 << 16 : tile bit held by register bit j
  uint32_t tb[3];    // tile bit walked by item-index bit j, 4 per word (byte each); 9 = K - 3 used at most
  uint32_t var;      // nvar | (src | pos << 1) << (8 + 8 j): src 0 = tile bit (of the local index), 1 = global bit
  uint32_t mat_off;  // first matrix of this round, in units of kRoundMatAmps
  uint32_t pad[2];
};

struct TilePassArgs {
  int k;                        // tile bits
  int n_rounds;
  int low_identity;             // L: number of low tile bits that are the low global bits
  int n_mats;                   // matrices used by this launch
  int pipelined;                // 1: double-buffered cp.async tile pipeline (one CTA per SM), 12-bit tiles only
  int pad1;
  uint64_t n_tiles;
  int tpos[kMaxTileBits];       // global bit position of tile bit j (ascending, tpos[j] == j for j < L)
  TileRoundDesc rounds[kMaxTileRounds];
  double2 mats[kMaxTileMats * kRoundMatAmps];  // round matrices, row-major 8x8, variant-major per round
};
static_assert(sizeof(TilePassArgs) <= 32764, "kernel parameter block limit");

// acc += m * a as four fused multiply-adds (the mat-vec of a round is fp64-pipe bound: 32 DFMA
// per amplitude; the unfused kernels keep the reference's separately rounded products instead)
__device__ __forceinline__ amp cfma(amp m, amp a, amp acc) {
  acc.x = fma(m.x, a.x, acc.x);
  acc.x = fma(-m.y, a.y, acc.x);
  acc.y = fma(m.x, a.y, acc.y);
  acc.y = fma(m.y, a.x, acc.y);
  return acc;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Dynamic shared memory: tile buffer(s) of 2^k amps (swizzled) | round matrices.  Default: two CTAs
// per SM, each with one tile buffer (one loads/stores while the other computes).  Optional
// (QCSIM_TILE_PIPE=1, 12-bit tiles): one CTA per SM with TWO tile buffers, the next tile streaming in
// with cp.async while the rounds of the current one run -- measured slightly slower on B200 in
// round 1 (39.5 vs 36.8 ms per 30-qubit layer: 8 warps per SM do not cover the LDS latency of the
// matrix loads), kept for the next round's tuning.
static __global__ void __launch_bounds__(kTileThreads, 2) k_tile_pass(amp* __restrict__ psi, const __grid_constant__ TilePassArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  amp* tile = reinterpret_cast<amp*>(smem_raw);
  const int k = A.k;
  const uint32_t tile_amps = 1u << k;
  const int L = A.low_identity;
  const uint32_t low_mask = (1u << L) - 1u;
  const uint32_t items = tile_amps >> kRoundBits;
  const uint32_t tid = threadIdx.x;

  // the launch's round matrices: parameter block -> shared memory, once per CTA.  A matrix entry is
  // then one uniform-address LDS.128 (a broadcast, one wavefront) instead of two indexed constant
  // loads: the indexed-constant cache saturates at half the fp64 rate of this kernel (ncu: IDC 50 %
  // busy at fp64 47 %).
  const bool pipe = (k == kMaxTileBits) && A.pipelined;
  amp* const buf0 = tile;
  amp* const buf1 = pipe ? tile + tile_amps : tile;
  amp* const smats = tile + (pipe ? 2u : 1u) * tile_amps;
  for (uint32_t i = tid; i < (uint32_t)A.n_mats * kRoundMatAmps; i += kTileThreads) smats[i] = A.mats[i];

  // global/shared offsets of this thread's amplitude pairs in the load/store phases: iteration
  // `it` handles local index loc = 2 * (tid + 256 * it); the tid part is fixed for the kernel
  const uint32_t loc_fixed = (tid << 1) & (tile_amps - 1u);
  uint64_t g_fixed = loc_fixed & low_mask;
#pragma unroll 1
  for (int j = L; j < k; ++j) g_fixed |= (uint64_t)((loc_fixed >> j) & 1u) << A.tpos[j];
  const uint32_t s_fixed = swz(loc_fixed);
  const uint32_t n_it = (tile_amps >> 1) > kTileThreads ? (tile_amps >> 1) / kTileThreads : 1u;
  const bool mover = (tid << 1) < tile_amps;

  auto gbase_of = [&](uint64_t t) {  // scatter t into the non-tile bit positions
    uint64_t g = t;
#pragma unroll 1
    for (int j = 0; j < k; ++j) g = insert_zero(g, A.tpos[j]);
    return g;
  };
  auto prefetch = [&](uint64_t t, amp* dst) {  // k == 12: 8 x 32 B per thread, asynchronous
    const uint64_t gb = gbase_of(t);
#pragma unroll
    for (uint32_t it = 0; it < 8; ++it) {
      const uint64_t gv = ((uint64_t)(it & 1u) << A.tpos[9]) | ((uint64_t)((it >> 1) & 1u) << A.tpos[10]) | ((uint64_t)((it >> 2) & 1u) << A.tpos[11]);
      const amp* src = psi + (gb | g_fixed | gv);
      const uint32_t s = s_fixed ^ swz(it << 9);
      cp_async16(dst + s, src);
      cp_async16(dst + (s ^ 1u), src + 1);
    }
  };
  if (pipe) {
    if (blockIdx.x < A.n_tiles) prefetch(blockIdx.x, buf0);
    cp_async_commit();
  }
  uint32_t cur = 0;

  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
    tile = cur ? buf1 : buf0;
    const uint64_t gbase = gbase_of(t);

    if (pipe) {
      // ---- the next tile starts streaming in; wait for this one only
      const uint64_t nt = t + gridDim.x;
      if (nt < A.n_tiles) prefetch(nt, cur ? buf0 : buf1);
      cp_async_commit();
      cp_async_wait<1>();
    } else if (mover) {
      // ---- HBM -> shared: each thread moves 2 adjacent amplitudes per 256-bit load; at K = 12 all
      // eight loads of a thread are in flight together (64 KiB per CTA)
      if (n_it == 8) {
        amp2 x[8];
#pragma unroll
        for (uint32_t it = 0; it < 8; ++it) {
          const uint64_t gv = ((uint64_t)(it & 1u) << A.tpos[9]) | ((uint64_t)((it >> 1) & 1u) << A.tpos[10]) | ((uint64_t)((it >> 2) & 1u) << A.tpos[11]);
          x[it] = ld_amp2(psi + (gbase | g_fixed | gv));
        }
#pragma unroll
        for (uint32_t it = 0; it < 8; ++it) {
          const uint32_t s = s_fixed ^ swz(it << 9);
          tile[s] = x[it].a;
          tile[s ^ 1u] = x[it].b;
        }
      } else {
#pragma unroll 1
        for (uint32_t it = 0; it < n_it; ++it) {
          const uint32_t lv = it << 9;  // bits 9.. of the local index (uniform)
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

    // ---- rounds: one 8x8 mat-vec per item; a thread does items (i, i + 256) with ONE set of matrix loads
#pragma unroll 1
    for (int r = 0; r < A.n_rounds; ++r) {
      const TileRoundDesc rd = A.rounds[r];
      const uint32_t so0 = swz(1u << (rd.rb & 31u)), so1 = swz(1u << ((rd.rb >> 8) & 31u)), so2 = swz(1u << ((rd.rb >> 16) & 31u));
      const int nvar = rd.var & 0x7f;
      const uint32_t n_var = 1u << nvar;
      const uint32_t pair_off = swz(1u << ((rd.tb[2]) & 31u));  // item bit 8 -> tile bit (only used when items > 256)
#pragma unroll 1
      for (uint32_t base = 0; base < items; base += 2 * kTileThreads) {
        const uint32_t item = base + tid;
        const bool valid = item < items;
        const bool two = item + kTileThreads < items;  // uniform: items is a power of two
        // item-index bits deposited on the non-round tile bits
        uint32_t lbase = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) lbase |= ((item >> j) & 1u) << ((rd.tb[j >> 2] >> (8 * (j & 3))) & 31u);
        uint32_t vidx = 0;
#pragma unroll
        for (int j = 0; j < kMaxVariantBits; ++j) {
          const uint32_t e = (rd.var >> (8 + 8 * j)) & 0xffu;
          const uint32_t bit = (e & 1u) ? (uint32_t)((gbase >> (e >> 1)) & 1ULL) : ((lbase >> (e >> 1)) & 1u);
          if (j < nvar) vidx |= bit << j;
        }
        const uint32_t sl = swz(lbase);
        uint32_t sa[8];
        amp v[2][8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          sa[x] = sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u);
          v[0][x] = valid ? tile[sa[x]] : make_amp(0, 0);
          v[1][x] = two ? tile[sa[x] ^ pair_off] : make_amp(0, 0);
        }
        // one trip per matrix variant, matching threads only (the host maps variant bits onto the
        // warp-index bits of the item index, so a warp computes in exactly one trip)
#pragma unroll 1
        for (uint32_t a = 0; a < n_var; ++a) {
          if (!valid || vidx != a) continue;
          const amp* __restrict__ M = smats + (size_t)(rd.mat_off + a) * kRoundMatAmps;
#pragma unroll
          for (int row = 0; row < 8; row += 2) {  // two rows x two items = 8 independent DFMA chains
            const amp ma = M[row * 8], mb = M[row * 8 + 8];
            amp a0 = cmul(ma, v[0][0]), a1 = cmul(ma, v[1][0]);
            amp b0 = cmul(mb, v[0][0]), b1 = cmul(mb, v[1][0]);
#pragma unroll
            for (int c = 1; c < 8; ++c) {
              const amp xa = M[row * 8 + c], xb = M[row * 8 + 8 + c];
              a0 = cfma(xa, v[0][c], a0);
              a1 = cfma(xa, v[1][c], a1);
              b0 = cfma(xb, v[0][c], b0);
              b1 = cfma(xb, v[1][c], b1);
            }
            tile[sa[row]] = a0;  // only this thread touches these slots in this round
            tile[sa[row + 1]] = b0;
            if (two) {
              tile[sa[row] ^ pair_off] = a1;
              tile[sa[row + 1] ^ pair_off] = b1;
            }
          }
        }
      }
      __syncthreads();
    }

    // ---- shared -> HBM
    if (mover) {
      if (n_it == 8) {
#pragma unroll
        for (uint32_t it = 0; it < 8; ++it) {
          const uint64_t gv = ((uint64_t)(it & 1u) << A.tpos[9]) | ((uint64_t)((it >> 1) & 1u) << A.tpos[10]) | ((uint64_t)((it >> 2) & 1u) << A.tpos[11]);
          const uint32_t s = s_fixed ^ swz(it << 9);
          amp2 x;
          x.a = tile[s];
          x.b = tile[s ^ 1u];
          st_amp2(psi + (gbase | g_fixed | gv), x);
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
    if (pipe) cur ^= 1u;
  }
  if (pipe) cp_async_wait<0>();
}

}  // namespace qcsim
