// qft_pipe.cu -- QFT / IQFT passes on the TMA-staged, warp-specialised tile pipeline (tile_pipe.cuh).
//
// Same mathematics as qft_kernels.cuh (radix-8/4/2 butterflies of the reference's own H / controlled-phase
// matrices + one twiddle per group of amplitudes, QuantumFourierTransform.h:35-87), different data movement: the
// tile loads and stores are TMA operations issued by a producer warp into a six-buffer mbarrier ring, two consumer
// groups take alternate tiles, so the load / butterfly / store phases of different tiles overlap.  (ncu of the
// round-1 kernel: 40 % of its samples were long-scoreboard stalls of the tile load phase, fp64 pipe 32 % busy,
// 12 ms per pass against a 5.4 ms HBM floor.)
//
// A pass holds qubits 0..2 (the innermost TMA box) + up to 8 more index bits; its transform qubits are cut into
// butterfly groups of 3 (or 2/1 for the remainder), one shared-memory round each.  Per group and item (an item =
// the 2^G amplitudes that differ in the group's bits) the host-side tables give the swizzled slot of the first
// amplitude and the item's part of the twiddle; the tile's part is one sincospi per tile and group.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.h"
#include "fusion.h"
#include "planner.h"
#include "qft_kernels.cuh"
#include "tile_pipe.cuh"

namespace qcsim {

namespace {

constexpr int kQftPipeMaxGroups = 4;
constexpr int kQftPipeMaxItems = 1280;  // (3,3,3,2): 3 x 256 + 512; 20 bytes per item fit the 28 KiB table area
static_assert(kQftPipeMaxItems * (sizeof(amp) + sizeof(uint32_t)) <= (size_t)kMaxTileMats * kRoundMatAmps * sizeof(amp), "table area");

struct QftPipeGroup {
  int size;            // 1..3 qubits
  int top_qubit;       // logical index of the group's highest qubit
  uint32_t rb;         // slot bit of the group's qubits, lowest logical first (byte each)
  uint32_t table_off;  // first entry of the group in the item tables
  uint32_t tb[3];      // slot bit walked by item-index bit j (byte each)
  uint32_t pad;
};

struct QftPipeArgs {
  PipeGeom geom;
  int n_groups;
  int inverse;
  int sq;      // lowest logical qubit of the whole transform
  int n_phys;  // physical index bits (local + rank bits)
  uint64_t rank_bits;
  signed char log_of[64];
  double s;
  double2 ph2, ph4;
  QftPipeGroup groups[kQftPipeMaxGroups];
  amp* tw_tab;          // [items] item part of the twiddle base (device global, filled by k_qft_pipe_tables)
  uint32_t* slot_tab;   // [items] swizzled slot of the item's first amplitude
  int n_items;
  int pad;
};

__device__ __forceinline__ uint32_t item_slot(const QftPipeGroup& grp, uint32_t item, int n_item_bits) {
  uint32_t lbase = 0;
  for (int j = 0; j < n_item_bits; ++j) lbase |= ((item >> j) & 1u) << ((grp.tb[j >> 2] >> (8 * (j & 3))) & 31u);
  return lbase;
}

__global__ void k_qft_pipe_tables(const __grid_constant__ QftPipeArgs A) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)A.n_items) return;
  int gi = 0;
  while (gi + 1 < A.n_groups && i >= A.groups[gi + 1].table_off) ++gi;
  const QftPipeGroup grp = A.groups[gi];
  const uint32_t item = i - grp.table_off;
  const uint32_t lbase = item_slot(grp, item, kPipeTileBits - grp.size);
  uint64_t phys = 0;
  for (int j = 0; j < kPipeTileBits; ++j) phys |= (uint64_t)((lbase >> j) & 1u) << A.geom.slot_pos[j];
  A.tw_tab[i] = qft_base(A, grp, qft_logical(A, phys), A.inverse != 0);
  A.slot_tab[i] = tswz(lbase);
}

template <int G>
__device__ __forceinline__ void qft_pipe_round(amp* __restrict__ tile, const QftPipeArgs& A, const QftPipeGroup& grp, const amp* __restrict__ tw,
                                               const uint32_t* __restrict__ slots, amp p_cta, uint32_t gtid, bool inverse) {
  constexpr uint32_t items = 1u << (kPipeTileBits - G);
  const uint32_t so0 = tswz(1u << (grp.rb & 31u)), so1 = tswz(1u << ((grp.rb >> 8) & 31u)), so2 = tswz(1u << ((grp.rb >> 16) & 31u));
#pragma unroll 1
  for (uint32_t item = gtid; item < items; item += kPipeGroupThreads) {
    const amp P = cmul(p_cta, tw[grp.table_off + item]);
    const uint32_t sl = slots[grp.table_off + item];
    amp v[8];
#pragma unroll
    for (int x = 0; x < (1 << G); ++x) v[x] = tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u)];
    qft_group<G>(v, A, inverse, P);
#pragma unroll
    for (int x = 0; x < (1 << G); ++x) tile[sl ^ ((x & 1) ? so0 : 0u) ^ ((x & 2) ? so1 : 0u) ^ ((x & 4) ? so2 : 0u)] = v[x];
  }
}

__global__ void __launch_bounds__(kPipeThreads, 1) k_qft_pipe(const __grid_constant__ QftPipeArgs A) {
  using namespace pipe;
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  const Smem sm = carve(pipe_smem);
  amp* const s_tw = sm.tables;                                                   // [n_items]
  uint32_t* const s_slot = reinterpret_cast<uint32_t*>(sm.tables + kQftPipeMaxItems);  // [n_items]
  amp* const s_cta = reinterpret_cast<amp*>(sm.done + kPipeStages);              // [group of consumers][transform group]
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  init_barriers(sm);
  for (uint32_t i = tid; i < (uint32_t)A.n_items; i += kPipeThreads) {
    s_tw[i] = A.tw_tab[i];
    s_slot[i] = A.slot_tab[i];
  }
  __syncthreads();
  if (warp == kPipeConsumerWarps) {
    producer(A.geom, sm);
    return;
  }
  const uint32_t group = warp / kPipeGroupWarps, gtid = tid - group * kPipeGroupThreads;
  const bool inverse = A.inverse != 0;
  amp* const my_cta = s_cta + group * kQftPipeMaxGroups;
  for (uint64_t t = blockIdx.x + (uint64_t)group * gridDim.x, i = group; t < A.geom.n_tiles; t += (uint64_t)kPipeGroups * gridDim.x, i += kPipeGroups) {
    const int s = (int)(i % kPipeStages);
    amp* const tile = sm.tiles + (size_t)s * (1u << kPipeTileBits);
    // the tile's part of every group's twiddle base: index bits outside the tile + the rank bits
    if (gtid < (uint32_t)A.n_groups) my_cta[gtid] = qft_base(A, A.groups[gtid], qft_logical(A, gbase_of(A.geom, t) | A.rank_bits), inverse);
    mbar_wait(&sm.full[s], (uint32_t)(i / kPipeStages) & 1u);
    group_bar(group);
#pragma unroll 1
    for (int gi = 0; gi < A.n_groups; ++gi) {
      const QftPipeGroup grp = A.groups[gi];
      const amp p_cta = my_cta[gi];
      if (grp.size == 3) qft_pipe_round<3>(tile, A, grp, s_tw, s_slot, p_cta, gtid, inverse);
      else if (grp.size == 2) qft_pipe_round<2>(tile, A, grp, s_tw, s_slot, p_cta, gtid, inverse);
      else qft_pipe_round<1>(tile, A, grp, s_tw, s_slot, p_cta, gtid, inverse);
      if (gi + 1 == A.n_groups) fence_proxy_async();
      group_bar(group);
    }
    if (gtid == 0) mbar_arrive(&sm.done[s]);
  }
}

constexpr size_t kQftPipeSmemBytes = 1024 + (size_t)kPipeStages * kPipeTileBytes + (size_t)kMaxTileMats * kRoundMatAmps * sizeof(amp) +
                                     2 * kPipeStages * sizeof(uint64_t) + kPipeGroups * kQftPipeMaxGroups * sizeof(amp);
static_assert(kQftPipeSmemBytes <= 227 * 1024, "shared memory per CTA");

}  // namespace

int qft_pipe_init_device_kernels() {
  CUDA_TRY(cudaFuncSetAttribute(k_qft_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQftPipeSmemBytes));
  return QCSIM_OK;
}

// how many of the next transform qubits (logical, in processing order `order`) fit one pipe pass: their physical
// bits + qubits 0..2 must fit the 11-bit tile
int qft_pipe_pass_capacity(const int* phys, int count) {
  uint64_t bits = 7;
  int n = 0;
  while (n < count && __builtin_popcountll(bits | (1ULL << phys[n])) <= kPipeTileBits) bits |= 1ULL << phys[n++];
  return n;
}

// One pass: the QFT / IQFT gates whose targets are the logical qubits [lo, hi] (all on local positions, fitting one
// tile with qubits 0..2).  Returns QCSIM_ERR_UNSUPPORTED when the pass has no TMA geometry (tiny registers).
int qft_pipe_pass(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of) {
  const int nl = h->n_local, k = kPipeTileBits;
  if (nl < k) return QCSIM_ERR_UNSUPPORTED;
  uint64_t bits = 7;
  for (int q = lo; q <= hi; ++q) bits |= 1ULL << phys_of[q];
  if (__builtin_popcountll(bits) > k) return fail(QCSIM_ERR_BAD_ARG, "internal: QFT pass does not fit a tile");
  for (int q = 0; q < nl && __builtin_popcountll(bits) < k; ++q) bits |= 1ULL << q;  // pad with the lowest unused bits
  std::vector<int> tile;
  for (int q = 0; q < nl; ++q)
    if ((bits >> q) & 1ULL) tile.push_back(q);
  TmaTileGeom g;
  if (!tma_tile_geometry(tile, nl, &g)) return QCSIM_ERR_UNSUPPORTED;
  static thread_local QftPipeArgs A;
  std::memset(&A, 0, sizeof A);
  const int rc = fusion_fill_pipe_geom(h, tile, g, &A.geom);
  if (rc != QCSIM_OK) return rc;
  int slot_of[64];
  for (int q = 0; q < 64; ++q) slot_of[q] = -1;
  for (int j = 0; j < k; ++j) slot_of[g.slot_qubit[j]] = j;
  // groups in processing order: top-down for the QFT, bottom-up for the IQFT; sizes 3,3,...,(2,2)|3|2|1
  std::vector<std::pair<int, int>> groups;  // (top logical qubit, size), listed top-down
  for (int top = hi; top >= lo;) {
    const int left = top - lo + 1;
    const int size = left > 4 ? 3 : left == 4 ? 2 : left;
    groups.push_back({top, size});
    top -= size;
  }
  if ((int)groups.size() > kQftPipeMaxGroups) return fail(QCSIM_ERR_BAD_ARG, "internal: too many QFT groups in one pass");
  if (inverse) std::reverse(groups.begin(), groups.end());
  A.n_groups = (int)groups.size();
  A.inverse = inverse ? 1 : 0;
  A.sq = r_floor;
  A.n_phys = h->n;
  A.rank_bits = (uint64_t)h->rank << nl;
  for (int q = 0; q < h->n; ++q) A.log_of[phys_of[q]] = (signed char)q;
  const double pi = 3.14159265358979323846, sign = inverse ? -1.0 : 1.0;
  A.s = 1. / std::sqrt(2.);
  A.ph2 = make_amp(std::cos(sign * pi / 2), std::sin(sign * pi / 2));  // std::polar(1., theta), QuantumGate.h:262-265
  A.ph4 = make_amp(std::cos(sign * pi / 4), std::sin(sign * pi / 4));
  uint32_t n_items = 0;
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    QftPipeGroup& G = A.groups[gi];
    G.size = groups[gi].second;
    G.top_qubit = groups[gi].first;
    G.table_off = n_items;
    n_items += 1u << (k - G.size);
    uint32_t reg_mask = 0;
    for (int j = 0; j < G.size; ++j) {
      const int sb = slot_of[phys_of[G.top_qubit - G.size + 1 + j]];
      G.rb |= (uint32_t)sb << (8 * j);
      reg_mask |= 1u << sb;
    }
    // item bits: one thread = one item; the low three item bits (the lanes of a quarter-warp) take slot bits with
    // distinct swizzle classes (slot bits 0..5, class = bit mod 3) where the group leaves them free
    int order[16], n = 0;
    bool used[16] = {false};
    for (int cls = 0; cls < 3; ++cls)
      for (int b = cls; b < 6; b += 3)
        if (!((reg_mask >> b) & 1u) && !used[b]) {
          order[n++] = b;
          used[b] = true;
          break;
        }
    for (int b = 0; b < k; ++b)
      if (!((reg_mask >> b) & 1u) && !used[b]) order[n++] = b;
    for (int j = 0; j < n; ++j) G.tb[j >> 2] |= (uint32_t)order[j] << (8 * (j & 3));
  }
  if (n_items > (uint32_t)kQftPipeMaxItems) return QCSIM_ERR_UNSUPPORTED;
  A.n_items = (int)n_items;
  // item tables live in the handle's QFT table buffer: twiddles (16 B) then slots (4 B)
  if (!h->d_qft_table) CUDA_TRY(cudaMalloc(&h->d_qft_table, (sizeof(amp) + sizeof(uint32_t)) * 4 * 2048));
  A.tw_tab = h->d_qft_table;
  A.slot_tab = reinterpret_cast<uint32_t*>(h->d_qft_table + 4 * 2048);
  k_qft_pipe_tables<<<(n_items + 255) / 256, 256, 0, h->stream>>>(A);
  const uint64_t grid = std::min<uint64_t>(A.geom.n_tiles, (uint64_t)kNumSMs);
  k_qft_pipe<<<(unsigned)grid, kPipeThreads, kQftPipeSmemBytes, h->stream>>>(A);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 2;
  h->stats.state_passes += 1;
  h->stats.bytes_moved += 32ULL * h->dim_local;
  h->stats.fused_rounds += groups.size();
  return QCSIM_OK;
}

}  // namespace qcsim
