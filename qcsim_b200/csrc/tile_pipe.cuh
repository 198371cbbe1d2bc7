// tile_pipe.cuh -- fused gate blocks, Blackwell data path: TMA-staged tiles, mbarrier ring, warp
// specialisation.  Same mathematics as tile_kernels.cuh (dense 8x8 rounds on a 2^12-amplitude
// shared-memory tile), different machine mapping:
//
//   * one persistent CTA per SM: 8 consumer warps + 1 producer warp, THREE 64 KiB tile buffers;
//   * the producer warp moves tiles with cp.async.bulk.tensor (TMA, SASS UTMALDG / UTMASTG): the
//     host describes the state vector as a rank-5 tensor whose dimensions are the contiguous groups
//     of tile qubits (planner.h: tma_tile_geometry), so one TMA op brings a box of 2^(3+w) amplitudes
//     (8 amplitudes = 128 B innermost) and 2^(9-w) ops fill a tile; completion is an mbarrier
//     transaction count (SYNCS), stores leave through bulk groups.  CU_TENSOR_MAP_SWIZZLE_128B makes
//     the hardware XOR the 16 B chunk index with the 128 B row index: slot ^ ((slot >> 3) & 7), the
//     bank swizzle the rounds need, for free;
//   * while the consumers run the rounds of tile i, tile i+1 is landing, tile i-1 is draining and the
//     load of tile i+2 is issued as soon as that drain has been read out: HBM traffic and the fp64
//     pipe overlap instead of alternating (the round-1 kernel spent >50 % of its time in one or the
//     other);
//   * consumers synchronise among themselves with a named barrier (bar.sync 1, 256); the producer
//     never joins it.
//
// Reference loops replaced: QubitRegisterCalculator.h:39-939 (one OpenMP pass per gate).
// Bound: HBM (32 B per amplitude per pass) up to ~3 dense rounds, fp64 pipe beyond.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "tile_kernels.cuh"

namespace qcsim {

constexpr int kPipeStages = 3;
constexpr int kPipeConsumerWarps = 8;
constexpr int kPipeConsumers = kPipeConsumerWarps * 32;  // two round items (2 x 8 amplitudes) per thread: one matrix fetch feeds both
constexpr int kPipeThreads = kPipeConsumers + 32;        // + producer warp
constexpr int kPipeTileBits = 12;
constexpr uint32_t kPipeTileBytes = (uint32_t)sizeof(amp) << kPipeTileBits;

struct PipePassArgs {
  CUtensorMap tmap;              // rank 5: dim 0 = qubits 0..2 (16 doubles), dims 1..4 = tile-qubit groups
  uint64_t n_tiles;
  int n_rounds;
  int n_mats;
  int n_enum;                    // tile qubits not covered by the TMA box: 2^n_enum ops per tile
  int box_bytes;                 // bytes one TMA op moves
  int dim_lo[5];                 // lowest index bit of tensor dimension i
  int dim_mask_bits[5];          // log2 of its extent (coordinate = (index >> lo) & mask); 0 for padding dims
  int enum_pos[9];               // index bit of enumerated tile qubit j
  int sorted_pos[kPipeTileBits]; // tile qubits ascending (to scatter the tile number around them)
  int slot_pos[kPipeTileBits];   // index bit held by shared-memory slot bit j
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
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kPipeConsumers) : "memory"); }

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
// tile[sa*] = M_I * v* for two items, M_I = matrix slot I of the parameter block: the entries sit at compile-time
// offsets of the parameter block, so they are fetched by uniform constant loads (LDCU.128 c[0x0][imm]) and reach
// the DFMAs as uniform-register operands -- no shared-memory traffic; one fetch feeds 8 DFMAs (two items).
template <int I>
__device__ __forceinline__ void matvec8x2(const PipePassArgs& A, const amp (&v0)[8], const amp (&v1)[8], amp* __restrict__ tile,
                                          const uint32_t (&sa)[8], uint32_t pair_off) {
#pragma unroll
  for (int row = 0; row < 8; row += 2) {  // two rows x two items = 8 independent DFMA chains
    const amp ma = A.mats[I * kRoundMatAmps + row * 8], mb = A.mats[I * kRoundMatAmps + row * 8 + 8];
    amp a0 = cmul(ma, v0[0]), a1 = cmul(ma, v1[0]);
    amp b0 = cmul(mb, v0[0]), b1 = cmul(mb, v1[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) {
      const amp xa = A.mats[I * kRoundMatAmps + row * 8 + c], xb = A.mats[I * kRoundMatAmps + row * 8 + 8 + c];
      a0 = cfma(xa, v0[c], a0);
      a1 = cfma(xa, v1[c], a1);
      b0 = cfma(xb, v0[c], b0);
      b1 = cfma(xb, v1[c], b1);
    }
    tile[sa[row]] = a0;  // only this thread touches these slots in this round, and it has read them all
    tile[sa[row + 1]] = b0;
    tile[sa[row] ^ pair_off] = a1;
    tile[sa[row + 1] ^ pair_off] = b1;
  }
}
}  // namespace pipe

// Dynamic shared memory (1024 B aligned): 3 tile buffers | round matrices | mbarriers.
static __global__ void __launch_bounds__(kPipeThreads, 1) k_tile_pipe(const __grid_constant__ PipePassArgs A) {
  using namespace pipe;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  // the 128 B swizzle pattern is a function of the shared-memory ADDRESS: tile buffers start on a 1024 B boundary
  unsigned char* const smem_al = pipe_smem + ((1024u - (smem_u32(pipe_smem) & 1023u)) & 1023u);
  amp* const tiles = reinterpret_cast<amp*>(smem_al);
  amp* const smats = tiles + (size_t)kPipeStages * (1u << kPipeTileBits);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smats + (size_t)kMaxTileMats * kRoundMatAmps);
  uint64_t* const full = bars;                 // [stage]: the tile has landed (TMA transaction bytes)
  uint64_t* const done = bars + kPipeStages;   // [stage]: the consumers have finished the rounds
  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31u;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kPipeStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto gbase_of = [&](uint64_t t) {  // scatter the tile number into the non-tile index bits
    uint64_t g = t;
#pragma unroll
    for (int j = 0; j < kPipeTileBits; ++j) g = insert_zero(g, A.sorted_pos[j]);
    return g;
  };

  if (warp == kPipeConsumerWarps) {
    // ------------------------------- producer warp: all tile movement -------------------------------
    const uint32_t n_ops = 1u << A.n_enum;
    const uint32_t box_amps = (uint32_t)A.box_bytes / (uint32_t)sizeof(amp);
    auto coords_of = [&](uint64_t gbase, uint32_t e, int* c) {
      uint64_t idx = gbase;
#pragma unroll 1
      for (int j = 0; j < A.n_enum; ++j) idx |= (uint64_t)((e >> j) & 1u) << A.enum_pos[j];
      c[0] = 0;
#pragma unroll
      for (int d = 1; d < 5; ++d) c[d] = (int)((idx >> A.dim_lo[d]) & ((1ULL << A.dim_mask_bits[d]) - 1ULL));
    };
    auto issue_load = [&](uint64_t t, int s) {
      if (lane == 0) mbar_arrive_expect_tx(&full[s], kPipeTileBytes);
      __syncwarp();
      const uint64_t gbase = gbase_of(t);
      amp* dst = tiles + (size_t)s * (1u << kPipeTileBits);
      for (uint32_t e = lane; e < n_ops; e += 32) {
        int c[5];
        coords_of(gbase, e, c);
        tma_load_5d(dst + (size_t)e * box_amps, &A.tmap, &full[s], c[0], c[1], c[2], c[3], c[4]);
      }
    };
    auto issue_store = [&](uint64_t t, int s) {
      const uint64_t gbase = gbase_of(t);
      const amp* src = tiles + (size_t)s * (1u << kPipeTileBits);
      for (uint32_t e = lane; e < n_ops; e += 32) {
        int c[5];
        coords_of(gbase, e, c);
        tma_store_5d(src + (size_t)e * box_amps, &A.tmap, c[0], c[1], c[2], c[3], c[4]);
      }
      bulk_commit();  // every lane closes its own (possibly empty) group: group counts stay aligned across lanes
    };
    const uint64_t step = gridDim.x;
    // prologue: two tiles in flight before the consumers start
    if (blockIdx.x < A.n_tiles) issue_load(blockIdx.x, 0);
    if (blockIdx.x + step < A.n_tiles) issue_load(blockIdx.x + step, 1);
    uint32_t i = 0;
    for (uint64_t t = blockIdx.x; t < A.n_tiles; t += step, ++i) {
      const int s = (int)(i % kPipeStages);
      mbar_wait(&done[s], (i / kPipeStages) & 1u);  // rounds of tile i finished, fenced for the async proxy
      issue_store(t, s);
      const uint64_t t2 = t + 2 * step;
      if (t2 < A.n_tiles) {
        // the buffer of tile i+2 is the one tile i-1 is draining from: wait until that drain has read it
        bulk_wait_read<1>();
        __syncwarp();
        issue_load(t2, (int)((i + 2) % kPipeStages));
      }
    }
    bulk_wait<0>();  // all stores complete before the CTA exits
    return;
  }

  // ------------------------------------ consumer warps: the rounds ------------------------------------
  uint32_t i = 0;
  for (uint64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x, ++i) {
    const int s = (int)(i % kPipeStages);
    amp* const tile = tiles + (size_t)s * (1u << kPipeTileBits);
    const uint64_t gbase = gbase_of(t);
    mbar_wait(&full[s], (i / kPipeStages) & 1u);

#pragma unroll 1
    for (int r = 0; r < A.n_rounds; ++r) {
      const TileRoundDesc rd = A.rounds[r];
      const uint32_t so0 = tswz(1u << (rd.rb & 31u)), so1 = tswz(1u << ((rd.rb >> 8) & 31u)), so2 = tswz(1u << ((rd.rb >> 16) & 31u));
      const int nvar = rd.var & 0xff;
      // item index deposited on the non-round slot bits: a thread does items (tid, tid + 256); item bit 8 walks
      // rd.tb byte 8 and never selects the matrix (the host keeps variant qubits on item bits 5..7)
      uint32_t lbase = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) lbase |= ((tid >> j) & 1u) << ((rd.tb[j >> 2] >> (8 * (j & 3))) & 31u);
      const uint32_t pair_off = tswz(1u << (rd.tb[2] & 31u));
      uint32_t vidx = 0;
#pragma unroll
      for (int j = 0; j < kMaxVariantBits; ++j) {
        const uint32_t e = (rd.var >> (8 + 8 * j)) & 0xffu;
        const uint32_t bit = (e & 1u) ? (uint32_t)((gbase >> (e >> 1)) & 1ULL) : ((lbase >> (e >> 1)) & 1u);
        if (j < nvar) vidx |= bit << j;
      }
      const uint32_t sl = tswz(lbase);
      uint32_t sa[8];
      amp v0[8], v1[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        sa[x] = sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u);
        v0[x] = tile[sa[x]];
        v1[x] = tile[sa[x] ^ pair_off];
      }
      // The variant is warp-uniform by construction.  One straight-line body per matrix slot (see matvec8x2).
      switch (rd.mat_off + vidx) {
#define QCSIM_PIPE_CASE(I) \
  case I:                  \
    pipe::matvec8x2<I>(A, v0, v1, tile, sa, pair_off); \
    break;
        QCSIM_PIPE_CASE(0) QCSIM_PIPE_CASE(1) QCSIM_PIPE_CASE(2) QCSIM_PIPE_CASE(3) QCSIM_PIPE_CASE(4) QCSIM_PIPE_CASE(5) QCSIM_PIPE_CASE(6)
        QCSIM_PIPE_CASE(7) QCSIM_PIPE_CASE(8) QCSIM_PIPE_CASE(9) QCSIM_PIPE_CASE(10) QCSIM_PIPE_CASE(11) QCSIM_PIPE_CASE(12) QCSIM_PIPE_CASE(13)
        QCSIM_PIPE_CASE(14) QCSIM_PIPE_CASE(15) QCSIM_PIPE_CASE(16) QCSIM_PIPE_CASE(17) QCSIM_PIPE_CASE(18) QCSIM_PIPE_CASE(19) QCSIM_PIPE_CASE(20)
        QCSIM_PIPE_CASE(21) QCSIM_PIPE_CASE(22) QCSIM_PIPE_CASE(23) QCSIM_PIPE_CASE(24) QCSIM_PIPE_CASE(25) QCSIM_PIPE_CASE(26) QCSIM_PIPE_CASE(27)
#undef QCSIM_PIPE_CASE
        default: break;
      }
      static_assert(kMaxTileMats == 28, "one switch case per matrix slot");
      if (r + 1 == A.n_rounds) fence_proxy_async();
      consumer_bar();
    }
    if (A.n_rounds == 0) {
      fence_proxy_async();
      consumer_bar();
    }
    if (tid == 0) mbar_arrive(&done[s]);
  }
}

}  // namespace qcsim
