// tile_pipe.cuh -- fused gate blocks, Blackwell data path: TMA-staged tiles, mbarrier ring, warp
// specialisation.  Same mathematics as tile_kernels.cuh (dense 8x8 rounds on a 2^12-amplitude
// shared-memory tile), different machine mapping:
//
//   * one persistent CTA per SM: 2 consumer groups of 8 warps + 1 producer warp, SIX 32 KiB tile buffers
//     (2^11 amplitudes: qubits 0..2 + 8 others); the groups take alternate tiles;
//   * the producer warp moves tiles with cp.async.bulk.tensor (TMA, SASS UTMALDG / UTMASTG): the
//     host describes the state vector as a rank-5 tensor whose dimensions are the contiguous groups
//     of tile qubits (planner.h: tma_tile_geometry), so one TMA op brings a box of 2^(3+w) amplitudes
//     (8 amplitudes = 128 B innermost) and 2^(9-w) ops fill a tile; completion is an mbarrier
//     transaction count (SYNCS), stores leave through bulk groups.  CU_TENSOR_MAP_SWIZZLE_128B makes
//     the hardware XOR the 16 B chunk index with the 128 B row index: slot ^ ((slot >> 3) & 7), the
//     bank swizzle the rounds need, for free;
//   * while the groups run the rounds of tiles i and i+1, tiles i+2, i+3 are landing, tiles i-1, i-2 are
//     draining and the load of tile i+4 is issued as soon as the drain of its buffer has been read out:
//     HBM traffic and the fp64 pipe overlap instead of alternating (the round-1 kernel spent >50 % of
//     its time in one or the other);
//   * a group synchronises its rounds with its own named barrier (bar.sync 1|2, 256); the producer never
//     joins it, and the two groups drift apart, so one group's fragment loads / barrier waits are covered
//     by the other group's MMAs;
//   * a round is a 16x16 REAL matrix ([Re -Im; Im Re] of the 8x8 complex round matrix) times a
//     16 x (items) panel: the consumers run it on the FP64 tensor path (mma.sync m8n8k4 f64, SASS
//     DMMA.8x8x4), 8 instructions per 8 items.  tcgen05 has no f64 kind; DMMA is the Blackwell FP64
//     matrix instruction and has the same 64 FMA/clk/SM peak as DFMA (tools/fp64_mix_bench.cu:
//     63 vs 61 measured) -- the point is not more flops but operand delivery: the round matrix sits
//     in 16 registers per lane for the whole round, where the DFMA mat-vec needed one 16-byte
//     operand fetch per 4 DFMA and stalled the LSU / constant path at 46 % of the fp64 peak however
//     the matrix was delivered (shared memory, indexed constants, uniform constant loads: measured).
//
// Reference loops replaced: QubitRegisterCalculator.h:39-939 (one OpenMP pass per gate).
// Bound: HBM (32 B per amplitude per pass) up to ~3 dense rounds, fp64 pipe beyond.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "tile_kernels.cuh"

namespace qcsim {

constexpr int kPipeTileBits = 11;                        // 2^11 amplitudes = 32 KiB per tile
constexpr int kPipeStages = 6;                           // tile buffers in the ring (192 KiB)
#ifndef QCSIM_PIPE_GROUPS
#define QCSIM_PIPE_GROUPS 2
#endif
#ifndef QCSIM_PIPE_LOOKAHEAD
#define QCSIM_PIPE_LOOKAHEAD 4
#endif
constexpr int kPipeGroups = QCSIM_PIPE_GROUPS;           // consumer groups; each works on its own tile
constexpr int kPipeGroupWarps = 8;                       // a warp owns 32 round items = 4 MMA panels of 8 items
constexpr int kPipeGroupThreads = kPipeGroupWarps * 32;
constexpr int kPipeConsumerWarps = kPipeGroups * kPipeGroupWarps;
constexpr int kPipeConsumers = kPipeConsumerWarps * 32;
constexpr int kPipeThreads = kPipeConsumers + 32;        // + producer warp
constexpr int kPipeLookahead = QCSIM_PIPE_LOOKAHEAD;     // tiles requested ahead of the one being drained
constexpr uint32_t kPipeTileBytes = (uint32_t)sizeof(amp) << kPipeTileBits;

// what the producer warp needs: the tensor map and how tile numbers / TMA ops turn into coordinates
struct PipeGeom {
  CUtensorMap tmap;              // rank 5: dim 0 = qubits 0..2 (16 doubles), dims 1..4 = tile-qubit groups
  uint64_t n_tiles;
  int n_enum;                    // tile qubits not covered by the TMA box: 2^n_enum ops per tile
  int box_bytes;                 // bytes one TMA op moves
  int dim_lo[5];                 // lowest index bit of tensor dimension i
  int dim_mask_bits[5];          // log2 of its extent (coordinate = (index >> lo) & mask); 0 for padding dims
  int enum_pos[9];               // index bit of enumerated tile qubit j
  int sorted_pos[kPipeTileBits]; // tile qubits ascending (to scatter the tile number around them)
  int slot_pos[kPipeTileBits];   // index bit held by shared-memory slot bit j
  int pad;
};

struct PipePassArgs {
  PipeGeom geom;
  int n_rounds;
  int n_mats;
  TileRoundDesc rounds[kMaxTileRounds];
  double2 mats[kMaxTileMats * kRoundMatAmps];
};
static_assert(sizeof(PipePassArgs) <= 32764, "kernel parameter block limit");

// the 128 B TMA swizzle on 16 B slots (tile buffers are 1024 B aligned)
__host__ __device__ __forceinline__ uint32_t tswz(uint32_t s) { return s ^ ((s >> 3) & 7u); }

namespace pipe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// generic-proxy writes (st.shared by the consumers) -> visible to the async proxy (TMA store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier of one consumer group (ids 1, 2); the producer and the other group never join it
__device__ __forceinline__ void group_bar(uint32_t group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "n"(kPipeGroupThreads) : "memory"); }

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* src, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace pipe

namespace pipe {
// D(8x8) += A(8x4) * B(4x8), fp64.  Fragments: A: lane holds A[lane >> 2][lane & 3]; B: lane holds
// B[lane & 3][lane >> 2]; C/D: lane holds rows lane >> 2, columns 2 (lane & 3) and 2 (lane & 3) + 1.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// D = A * B + C with D and C in different registers (C stays live)
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b, const double (&c)[2]) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};" : "=d"(d[0]), "=d"(d[1]) : "d"(a), "d"(b), "d"(c[0]), "d"(c[1]));
}
// x with its sign flipped when `flip` is 1 (one integer XOR on the high word)
__device__ __forceinline__ double flip_sign(double x, uint32_t flip) {
  return __hiloint2double(__double2hiint(x) ^ (int)(flip << 31), __double2loint(x));
}
}  // namespace pipe

#ifdef QCSIM_PIPE_PROFILE
// diagnostics build only: cycles per phase, summed over warps (lane 0), read back by launch_pass_pipe
// [0] consumer total  [1] wait for the tile  [2] rounds incl. barriers  [3] barriers  [4] producer total  [5] producer wait done
// [6] producer wait for store reads  [7] consumer warps counted  [8] first-round loads issued -> round end (round 0 only)
__device__ unsigned long long g_pipe_prof[16];
#define PROF_T() clock64()
#define PROF_ADD(i, v) do { if ((threadIdx.x & 31u) == 0) atomicAdd(&g_pipe_prof[i], (unsigned long long)(v)); } while (0)
#else
#define PROF_T() 0LL
#define PROF_ADD(i, v) do { } while (0)
#endif

namespace pipe {

// scatter the tile number into the non-tile index bits
__device__ __forceinline__ uint64_t gbase_of(const PipeGeom& G, uint64_t t) {
  uint64_t g = t;
#pragma unroll
  for (int j = 0; j < kPipeTileBits; ++j) g = insert_zero(g, G.sorted_pos[j]);
  return g;
}

// carve the dynamic shared memory: 6 tile buffers (1024 B aligned: the 128 B swizzle pattern is a function of the
// shared-memory ADDRESS) | 28 KiB of per-pass tables (round matrices / twiddle tables) | mbarriers
struct Smem {
  amp* tiles;
  amp* tables;
  uint64_t* full;  // [stage]: the tile has landed (TMA transaction bytes)
  uint64_t* done;  // [stage]: the consumers have finished the rounds
};
__device__ __forceinline__ Smem carve(unsigned char* raw) {
  Smem m;
  unsigned char* const al = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  m.tiles = reinterpret_cast<amp*>(al);
  m.tables = m.tiles + (size_t)kPipeStages * (1u << kPipeTileBits);
  m.full = reinterpret_cast<uint64_t*>(m.tables + (size_t)kMaxTileMats * kRoundMatAmps);
  m.done = m.full + kPipeStages;
  return m;
}
// k_tile_pipe's carve: per-pass tables, mbarriers and round tables first, then the tile buffers on a 32 KiB boundary of
// the shared window -- a tile's base address then has its low 15 bits clear and every fragment address of a round is
// an XOR of precomputed parts (no adds).  `extra` = bytes needed behind the mbarriers (round tables).
constexpr uint32_t kPipeHeadBytes = 32768;  // tables + mbarriers + round tables + the pad up to the boundary (the window starts at 1 KiB)
__device__ __forceinline__ Smem carve_tiles_aligned(unsigned char* raw, uint32_t extra) {
  Smem m;
  const uint32_t sa0 = smem_u32(raw);
  unsigned char* const al = raw + ((16u - (sa0 & 15u)) & 15u);
  m.tables = reinterpret_cast<amp*>(al);
  m.full = reinterpret_cast<uint64_t*>(m.tables + (size_t)kMaxTileMats * kRoundMatAmps);
  m.done = m.full + kPipeStages;
  const uint32_t head_end = smem_u32(m.done + kPipeStages) + extra;
  const uint32_t tiles_sa = (head_end + 32767u) & ~32767u;
  uint32_t dyn;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
  if (tiles_sa + kPipeStages * kPipeTileBytes > sa0 + dyn) __trap();  // the launch did not provide the head room
  m.tiles = reinterpret_cast<amp*>(raw + (tiles_sa - sa0));
  return m;
}
__device__ __forceinline__ double lds_f64(uint32_t sa) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa));  // volatile: stays ordered with the barrier / mbarrier asm
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t sa, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(sa), "d"(v)); }
__device__ __forceinline__ void init_barriers(const Smem& m) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) {
      mbar_init(&m.full[s], 1);
      mbar_init(&m.done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
}

// the producer warp: all tile movement of the CTA
__device__ __forceinline__ void producer(const PipeGeom& G, const Smem& m) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n_ops = 1u << G.n_enum;
  const uint32_t box_amps = (uint32_t)G.box_bytes / (uint32_t)sizeof(amp);
  auto coords_of = [&](uint64_t gbase, uint32_t e, int* c) {
    uint64_t idx = gbase;
#pragma unroll 1
    for (int j = 0; j < G.n_enum; ++j) idx |= (uint64_t)((e >> j) & 1u) << G.enum_pos[j];
    c[0] = 0;
#pragma unroll
    for (int d = 1; d < 5; ++d) c[d] = (int)((idx >> G.dim_lo[d]) & ((1ULL << G.dim_mask_bits[d]) - 1ULL));
  };
  auto issue_load = [&](uint64_t t, int s) {
    if (lane == 0) mbar_arrive_expect_tx(&m.full[s], kPipeTileBytes);
    __syncwarp();
    const uint64_t gbase = gbase_of(G, t);
    amp* dst = m.tiles + (size_t)s * (1u << kPipeTileBits);
    for (uint32_t e = lane; e < n_ops; e += 32) {
      int c[5];
      coords_of(gbase, e, c);
      tma_load_5d(dst + (size_t)e * box_amps, &G.tmap, &m.full[s], c[0], c[1], c[2], c[3], c[4]);
    }
  };
  auto issue_store = [&](uint64_t t, int s) {
    const uint64_t gbase = gbase_of(G, t);
    const amp* src = m.tiles + (size_t)s * (1u << kPipeTileBits);
    for (uint32_t e = lane; e < n_ops; e += 32) {
      int c[5];
      coords_of(gbase, e, c);
      tma_store_5d(src + (size_t)e * box_amps, &G.tmap, c[0], c[1], c[2], c[3], c[4]);
    }
    bulk_commit();  // every lane closes its own (possibly empty) group: group counts stay aligned across lanes
  };
  const uint64_t step = gridDim.x;
  // prologue: kPipeLookahead tiles in flight before the consumers start
#pragma unroll 1
  for (int j = 0; j < kPipeLookahead; ++j)
    if (blockIdx.x + j * step < G.n_tiles) issue_load(blockIdx.x + j * step, j);
  uint32_t i = 0;
  const long long p_t0 = PROF_T();
  long long p_done = 0, p_read = 0;
  for (uint64_t t = blockIdx.x; t < G.n_tiles; t += step, ++i) {
    const int s = (int)(i % kPipeStages);
    const long long q0 = PROF_T();
    mbar_wait(&m.done[s], (i / kPipeStages) & 1u);  // rounds of tile i finished, fenced for the async proxy
    p_done += PROF_T() - q0;
    issue_store(t, s);
    const uint64_t t2 = t + kPipeLookahead * step;
    if (t2 < G.n_tiles) {
      // the buffer of tile i + lookahead is the one tile i + lookahead - stages drained from: wait until that
      // drain has been read out of shared memory (all but the most recent stages - lookahead groups)
      const long long q1 = PROF_T();
      bulk_wait_read<kPipeStages - kPipeLookahead>();
      __syncwarp();
      p_read += PROF_T() - q1;
      issue_load(t2, (int)((i + kPipeLookahead) % kPipeStages));
    }
  }
  PROF_ADD(4, PROF_T() - p_t0);
  PROF_ADD(5, p_done);
  PROF_ADD(6, p_read);
  bulk_wait<0>();  // all stores complete before the CTA exits
}

}  // namespace pipe

// per-round lookup tables of k_tile_pipe (shared memory, filled once per CTA)
struct RoundTable {
  // byte offsets inside a tile (swizzled slot * 16), combined by XOR
  uint32_t load_g[8], store_g[8], load_t[4], store_t[4], warp_hi[8], warp_ms[8];
  uint32_t x_hi, x_i0, x_p0, x_p1;
  uint32_t chain_next;  // the next round keeps every warp on its own amplitudes: __syncwarp() instead of the group barrier
  uint32_t pad[3];
};

// Dynamic shared memory (1024 B aligned): 6 tile buffers | round matrices | mbarriers | round tables.
static __global__ void __launch_bounds__(kPipeThreads, 1) k_tile_pipe(const __grid_constant__ PipePassArgs A) {
  using namespace pipe;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  const Smem sm = carve_tiles_aligned(pipe_smem, kMaxTileRounds * (uint32_t)sizeof(RoundTable));
  amp* const tiles = sm.tiles;
  amp* const smats = sm.tables;
  uint64_t* const full = sm.full;
  uint64_t* const done = sm.done;
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31u;
  init_barriers(sm);
  // round matrices: parameter block -> shared memory once per CTA (a lane reads two entries per round; per-lane
  // addresses would serialise in the constant cache)
  // Entry [o][a] of a matrix goes to slot (a >> 2) * 32 + o * 4 + (a & 3): lane (g, t) of a warp reads entries [g][t] and
  // [g][4 + t], i.e. slots `lane` and 32 + lane -- two conflict-free LDS.128 (row-major slots g * 8 + t were 2-way).
  for (uint32_t i = tid; i < (uint32_t)A.n_mats * kRoundMatAmps; i += kPipeThreads) {
    const uint32_t e = i & 63u, o = e >> 3, a = e & 7u;
    smats[(i & ~63u) + ((a >> 2) << 5) + (o << 2) + (a & 3u)] = A.mats[i];
  }
  __syncthreads();

  if (warp == kPipeConsumerWarps) {
    producer(A.geom, sm);
    return;
  }

  // ------------------------------------ consumer warps: the rounds ------------------------------------
  // Round r multiplies every group of 8 amplitudes that differ in the round's 3 register bits by an 8x8 complex
  // matrix M (one per value of the variant bits), as THREE real 8x8 products on the FP64 tensor path (the 3-multiplication
  // form of the complex product; the plain real form [Re M, -Im M; Im M, Re M] needs four).
  // MMA mapping (per warp, 4 panels of 8 items): rows = output amplitude o, columns = items, contraction = input amplitude
  // a in 2 K-blocks (a >> 2), k = a & 3:
  //   lane (g = lane >> 2, t = lane & 3) holds   A: entries [g][t], [g][4 + t] of the three real matrices (from M, once per round)
  //                                              B: amplitudes t and 4 + t of item g of the panel (4 LDS.64: both halves of each)
  //                                              D: output amplitude g of items 2t and 2t + 1 (4 STS.64: both halves of each)
  // Fragments move as 64-bit halves: odd k-lanes (t & 1) fetch the imaginary half first, odd rows (g & 1) store it
  // first.  A half-warp then covers 16 distinct 8-byte columns of a 128 B row even when register bit 0 sits where the
  // TMA swizzle does not fold (planner.h, swizzle_kind 2); the parity is absorbed by the operands (see the round body).
  // Item index (8 bits): bits 0..2 = column in the panel, bits 3..4 = panel, bits 5..7 = warp of the group.
  // Two consumer groups take alternate tiles, each with its own named barrier: while one group waits for its
  // fragment loads or its round barrier, the other keeps the fp64 pipe busy.
  const uint32_t g = lane >> 2, tq = lane & 3u;
  const uint32_t group = warp / kPipeGroupWarps, gwarp = warp % kPipeGroupWarps;

  // Everything about a round that does not depend on the tile is tabulated ONCE per CTA in shared memory: the
  // swizzled slot of a lane's fragment loads / result stores is the XOR of a part chosen by its column / row index g,
  // a part chosen by its k-lane tq and a part chosen by its warp (the swizzle is linear over XOR); plus the slot strides
  // of the panel / register / column bits and, per warp, the matrix it uses.  (Per-thread arrays indexed by the round
  // went to local memory, and with all of L1 carved out as shared memory every such load was an L2 round trip at the
  // head of every round.)
  RoundTable* const rt = reinterpret_cast<RoundTable*>(sm.done + kPipeStages);
  for (uint32_t e = tid; e < (uint32_t)A.n_rounds * 48u; e += kPipeThreads) {
    const uint32_t r = e / 48u, j = e % 48u;
    const TileRoundDesc rd = A.rounds[r];
    const uint32_t rb0 = 1u << (rd.rb & 31u), rb1 = 1u << ((rd.rb >> 8) & 31u), rb2 = 1u << ((rd.rb >> 16) & 31u);
    auto ib = [&](int q) { return 1u << ((rd.tb[q >> 2] >> (8 * (q & 3))) & 31u); };  // slot bit walked by item bit q
    if (j < 8) {  // loads: item g of the panel
      rt[r].load_g[j] = tswz(((j & 1u) ? ib(0) : 0u) | ((j & 2u) ? ib(1) : 0u) | ((j & 4u) ? ib(2) : 0u)) << 4;
    } else if (j < 16) {  // stores: amplitude g
      const uint32_t x = j - 8;
      rt[r].store_g[x] = tswz(((x & 1u) ? rb0 : 0u) | ((x & 2u) ? rb1 : 0u) | ((x & 4u) ? rb2 : 0u)) << 4;
    } else if (j < 20) {  // loads: amplitudes tq (+4)
      const uint32_t x = j - 16;
      rt[r].load_t[x] = (tswz(((x & 1u) ? rb0 : 0u) | ((x & 2u) ? rb1 : 0u)) << 4) | ((x & 1u) << 3);  // + the half fetched first
    } else if (j < 24) {  // stores: items 2 tq (+1)
      const uint32_t x = j - 20;
      rt[r].store_t[x] = tswz(((x & 1u) ? ib(1) : 0u) | ((x & 2u) ? ib(2) : 0u)) << 4;
    } else if (j < 32) {  // warp part of the item index, and the matrix the warp uses
      const uint32_t w = j - 24;
      uint32_t hi = 0;
      for (int q = 0; q < 3; ++q) hi |= ((w >> q) & 1u) ? ib(5 + q) : 0u;
      const int nvar = rd.var & 0x7f;
      uint32_t vidx = 0, gs = 0;
      for (int q = 0; q < kMaxVariantBits; ++q) {
        const uint32_t en = (rd.var >> (8 + 8 * q)) & 0xffu;
        if (q < nvar) {
          if (en & 1u) gs |= (0x80u | (en >> 1)) << (8 * q);  // index bit outside the tile: resolved per tile
          else vidx |= ((hi >> (en >> 1)) & 1u) << q;
        }
      }
      rt[r].warp_hi[w] = tswz(hi) << 4;
      rt[r].warp_ms[w] = gs | ((rd.mat_off + vidx) << 24);
    } else if (j == 32) {
      rt[r].x_hi = tswz(rb2) << 4;
      rt[r].x_i0 = tswz(ib(0)) << 4;
      rt[r].x_p0 = tswz(ib(3)) << 4;
      rt[r].x_p1 = tswz(ib(4)) << 4;
      rt[r].chain_next = (rd.var >> 7) & 1u;
    }
  }
  asm volatile("bar.sync 8, %0;" ::"n"(kPipeConsumers) : "memory");  // all consumers (the producer is already moving tiles)

  const long long c_t0 = PROF_T();
  long long c_full = 0, c_rounds = 0, c_bar = 0;
  for (uint64_t t = blockIdx.x + (uint64_t)group * gridDim.x, i = group; t < A.geom.n_tiles; t += (uint64_t)kPipeGroups * gridDim.x, i += kPipeGroups) {
    const int s = (int)(i % kPipeStages);
    const uint32_t tile_sa = smem_u32(tiles) + (uint32_t)s * kPipeTileBytes;  // low 15 bits clear
    const uint64_t gbase = gbase_of(A.geom, t);
    const long long w0 = PROF_T();
    mbar_wait(&full[s], (uint32_t)(i / kPipeStages) & 1u);
    const long long w1 = PROF_T();
    c_full += w1 - w0;

#pragma unroll 1
    for (int r = 0; r < A.n_rounds; ++r) {
      {
        const RoundTable& T = rt[r];
        const uint32_t wh = T.warp_hi[gwarp], wm = T.warp_ms[gwarp];
        const uint32_t lq = tq & 1u, sq = g & 1u;  // half taken first by this lane's loads / stores (0 = real)
        // shared-window byte addresses of this lane's fragment halves: tile base ^ warp part ^ lane parts (^ panel part)
        const uint32_t ld0 = tile_sa ^ wh ^ T.load_g[g] ^ T.load_t[tq];                     // first half of amplitude tq, panel 0
        const uint32_t st0 = tile_sa ^ wh ^ T.store_g[g] ^ T.store_t[tq] ^ (sq << 3);       // first half of output g, item 2 tq, panel 0
        uint32_t midx = wm >> 24;
#pragma unroll
        for (int j = 0; j < kMaxVariantBits; ++j) {
          const uint32_t e = (wm >> (8 * j)) & 0xffu;
          if (e & 0x80u) midx += (uint32_t)((gbase >> (e & 0x7fu)) & 1ULL) << j;
        }
        // A fragments of this warp's matrix
        const amp* __restrict__ M = smats + (size_t)midx * kRoundMatAmps;
        const double2 m_lo = M[lane], m_hi = M[32 + lane];
        const uint32_t x_hi = T.x_hi, x_i0 = T.x_i0, x_p0 = T.x_p0, x_p1 = T.x_p1;
        uint32_t chain_next;  // read here, with the other table entries: at the end of the round its latency sat in front of the barrier
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(chain_next) : "r"(smem_u32(&T.chain_next)));
        double b[4][4];                            // [panel]: first half of amplitudes tq, 4 + tq, then their second halves
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const uint32_t a0 = ld0 ^ ((p & 1) ? x_p0 : 0u) ^ ((p & 2) ? x_p1 : 0u);
          b[p][0] = lds_f64(a0);
          b[p][1] = lds_f64(a0 ^ x_hi);
          b[p][2] = lds_f64(a0 ^ 8u);
          b[p][3] = lds_f64(a0 ^ x_hi ^ 8u);
        }
        // The parity of the halves is folded into the operands instead of being undone with selects: with
        // v'_a = (-i)^(a & 1) v_a and o'_o = (-i)^(o & 1) o_o the product is o' = M' v', M'[o][a] = (-i)^(o & 1) i^(a & 1) M[o][a], and
        //   Re v'_a = the half this lane fetched first,  Im v'_a = +-(the half fetched second)   (- for odd a, i.e. lq),
        //   Re o'_o = the half this lane stores first,   Im o'_o = +-(the half stored second)    (- for odd o, i.e. sq).
        // Three real 8x8 products instead of four (M' = P + iQ, v' = x + iy):
        //   S = P (x + y),   Re o' = S - (P + Q) y,   Im o' = S + (Q - P) x
        // = 6 DMMA per panel of 8 items: 2 for S, then 2 + 2 that start from S.
        // M' = (-i)^sq i^lq M: unchanged when lq == sq, times i (lq = 1) or -i (lq = 0) otherwise; signs by integer XOR (a DADD
        // negation would queue behind the other warps' DMMA in the FP64 pipe)
        const bool turn = lq != sq;
        const double2 ml = turn ? make_double2(flip_sign(m_lo.y, lq), flip_sign(m_lo.x, lq ^ 1u)) : m_lo;
        const double2 mh = turn ? make_double2(flip_sign(m_hi.y, lq), flip_sign(m_hi.x, lq ^ 1u)) : m_hi;
        const double n1[2] = {ml.x, mh.x};
        const double n2[2] = {-(ml.x + ml.y), -(mh.x + mh.y)};
        const double n3[2] = {ml.y - ml.x, mh.y - mh.x};
        double xs[4][2], S[4][2], c_first[4][2], c_second[4][2];
#ifndef QCSIM_PIPE_HALF_ROUNDS
#define QCSIM_PIPE_HALF_ROUNDS 0
#endif
        // QCSIM_PIPE_HALF_ROUNDS: the panels go through DMMA and out to shared memory two at a time, so that the first
        // half's stores run under the second half's DMMA (experiment switch; 0 = all four panels at once)
        constexpr int kPanelsAtOnce = QCSIM_PIPE_HALF_ROUNDS ? 2 : 4;
#pragma unroll
        for (int p0 = 0; p0 < 4; p0 += kPanelsAtOnce) {
#pragma unroll
          for (int p = p0; p < p0 + kPanelsAtOnce; ++p) {
            b[p][2] = flip_sign(b[p][2], lq);
            b[p][3] = flip_sign(b[p][3], lq);
            xs[p][0] = b[p][0] + b[p][2];
            xs[p][1] = b[p][1] + b[p][3];
            S[p][0] = S[p][1] = 0.0;
          }
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int p = p0; p < p0 + kPanelsAtOnce; ++p) dmma884(S[p], n1[kb], xs[p][kb]);
          }
#pragma unroll
          for (int p = p0; p < p0 + kPanelsAtOnce; ++p) {
            dmma884(c_first[p], n2[0], b[p][2], S[p]);
            dmma884(c_second[p], n3[0], b[p][0], S[p]);
          }
#pragma unroll
          for (int p = p0; p < p0 + kPanelsAtOnce; ++p) {
            dmma884(c_first[p], n2[1], b[p][3]);
            dmma884(c_second[p], n3[1], b[p][1]);
          }
          // A lane stores into slots other lanes of the warp loaded from: the loads are complete (their values fed the
          // warp-wide mma.sync above); the explicit warp barrier states that ordering for the memory model / racecheck
          __syncwarp();
#pragma unroll
          for (int p = p0; p < p0 + kPanelsAtOnce; ++p) {
            const uint32_t a0 = st0 ^ ((p & 1) ? x_p0 : 0u) ^ ((p & 2) ? x_p1 : 0u);
            sts_f64(a0, c_first[p][0]);
            sts_f64(a0 ^ x_i0, c_first[p][1]);
            sts_f64(a0 ^ 8u, flip_sign(c_second[p][0], sq));
            sts_f64(a0 ^ x_i0 ^ 8u, flip_sign(c_second[p][1], sq));
          }
        }
        if (r + 1 == A.n_rounds) fence_proxy_async();
        const long long b0 = PROF_T();
        if (chain_next && r + 1 < A.n_rounds) __syncwarp();
        else group_bar(group);
        c_bar += PROF_T() - b0;
      }
    }
    c_rounds += PROF_T() - w1;
    if (A.n_rounds == 0) {
      fence_proxy_async();
      group_bar(group);
    }
    if (gwarp == 0 && lane == 0) mbar_arrive(&done[s]);
  }
  PROF_ADD(0, PROF_T() - c_t0);
  PROF_ADD(1, c_full);
  PROF_ADD(2, c_rounds);
  PROF_ADD(3, c_bar);
  PROF_ADD(7, 1);
}

}  // namespace qcsim
