// fusion.cu -- gate-stream execution: QFT recognition, pass planning, round matrices, launches
// (see fusion.h, planner.h, tile_kernels.cuh).
//
//   * fusion_execute      : cuts QFT / IQFT gate streams out of the list (planner.h: match_qft) and
//                           runs them as radix-8 passes; the rest goes to the sharded planner or to
//   * fusion_execute_local: plan_passes (greedy, order-preserving: an op joins the current pass when
//                           its non-diagonal targets fit the tile and it commutes with every op
//                           skipped before it), then per pass schedule_rounds + one 8x8 matrix per
//                           round and variant (apply_small), then k_tile_pass;
//   * a pass whose rounds would cost more than its gates run one by one is executed unfused.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cudaTypedefs.h>

#include "fusion.h"
#include "planner.h"
#include "tile_kernels.cuh"
#include "tile_pipe.cuh"

namespace qcsim {

namespace {

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : dflt;
}

}  // namespace

// head (round matrices, mbarriers, round tables, padded to the first 32 KiB boundary of the shared window) + tile ring
constexpr size_t kPipeHeadUsed = 16 + (size_t)kMaxTileMats * kRoundMatAmps * sizeof(amp) + 2 * kPipeStages * sizeof(uint64_t) +
                                 kMaxTileRounds * sizeof(RoundTable);
static_assert(kPipeHeadUsed + 1024 /* the window's reserved first KiB */ <= pipe::kPipeHeadBytes, "head does not fit below the first tile");
constexpr size_t kPipeSmemBytes = pipe::kPipeHeadBytes + (size_t)kPipeStages * kPipeTileBytes;
static_assert(kPipeSmemBytes <= 227 * 1024, "shared memory per CTA");

int fusion_init_device_kernels() {  // per device, from engine_create
  CUDA_TRY(cudaFuncSetAttribute(k_tile_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + 28) * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_tile_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmemBytes));
  QCSIM_TRY(qft_pipe_init_device_kernels());
  return QCSIM_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }();
  return fn;
}

bool fusion_pipe_available() {
  static const int legacy = env_int("QCSIM_TILE_LEGACY", 0);
  return !legacy && tensor_map_encoder() != nullptr;
}

// Tensor map + coordinate recipe of a tile set (tile_pipe.cuh: PipeGeom) for the handle's state buffer.
int fusion_fill_pipe_geom(qcsim_sv* h, const std::vector<int>& tile_sorted, const TmaTileGeom& g, PipeGeom* G) {
  PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
  if (!encode) return QCSIM_ERR_UNSUPPORTED;
  const int k = kPipeTileBits;
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estride[5] = {1, 1, 1, 1, 1};
  gdim[0] = 16;  // doubles: 8 amplitudes
  box[0] = 16;
  for (int d = 1; d < 5; ++d) {
    gdim[d] = 1ULL << g.dim_bits[d];
    box[d] = 1u << g.box_bits[d];
    gstride[d - 1] = (cuuint64_t)sizeof(amp) << g.dim_lo[d];  // bytes between consecutive coordinates of dim d
  }
  const CUresult cr = encode(&G->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, h->psi, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(QCSIM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
  G->n_tiles = 1ULL << (h->n_local - k);
  G->n_enum = g.n_enum;
  G->box_bytes = (int)(sizeof(amp) << g.box_log2);
  for (int d = 0; d < 5; ++d) {
    G->dim_lo[d] = g.dim_lo[d];
    G->dim_mask_bits[d] = g.dim_bits[d];
  }
  for (int j = 0; j < 9; ++j) G->enum_pos[j] = j < g.n_enum ? g.enum_pos[j] : 0;
  for (int j = 0; j < k; ++j) {
    G->sorted_pos[j] = tile_sorted[j];
    G->slot_pos[j] = g.slot_qubit[j];
  }
  G->pad = 0;
  return QCSIM_OK;
}

// The TMA-staged, warp-specialised pass (tile_pipe.cuh).  Returns QCSIM_ERR_UNSUPPORTED when the tile
// set has no TMA geometry (small registers), in which case the caller uses k_tile_pass.
static int launch_pass_pipe(qcsim_sv* h, const std::vector<Op>& all, const PassPlan& plan_in) {
  TmaTileGeom g;
  if ((int)plan_in.tile.size() != kPipeTileBits || !tma_tile_geometry(plan_in.tile, h->n_local, &g)) return QCSIM_ERR_UNSUPPORTED;
  if (!tensor_map_encoder()) return QCSIM_ERR_UNSUPPORTED;
  const int k = kPipeTileBits;
  // Tensor-dimension order = shared-memory slot order of the boxed qubits, chosen for the fewest bank conflicts of
  // this pass's DMMA rounds (planner.h); everything downstream (round bits, item bits, variants) is in slot bits.
  static const int layout_search = env_int("QCSIM_PIPE_LAYOUT", 1), always_chain = env_int("QCSIM_PIPE_CHAIN", 0);
  static thread_local PipePassArgs A;  // ~30 KiB: keep it off the stack; the launch copies it
  static thread_local bool reorder_rejected = false;  // the driver refused a reordered tensor map once: stay on the ascending order
  PassPlan plan = plan_in;
  std::vector<RoundPlan> rplan;
  for (int attempt = 0;; ++attempt) {
    TmaTileGeom chosen = g;
    rplan = schedule_rounds_best_layout(all, plan_in, g, kMaxVariantBits, &chosen, &plan, nullptr, always_chain != 0,
                                        layout_search != 0 && !reorder_rejected);
    const int rc = fusion_fill_pipe_geom(h, plan_in.tile, chosen, &A.geom);
    if (rc == QCSIM_OK) break;
    if (attempt > 0 || reorder_rejected || !layout_search) return rc;
    reorder_rejected = true;
  }
  int local_of[64];
  for (int q = 0; q < 64; ++q) local_of[q] = -1;
  for (int j = 0; j < k; ++j) local_of[plan.tile[j]] = j;

  const uint64_t grid = std::min<uint64_t>(A.geom.n_tiles, (uint64_t)kNumSMs);
  static const int debug = env_int("QCSIM_DEBUG_PLAN", 0);

  size_t r = 0;
  do {  // a pass without rounds cannot happen (plan_passes only fuses >= 2 ops), but the loop tolerates it
    int n_rounds = 0;
    size_t mat_index = 0, n_ops = 0;
    while (r < rplan.size() && n_rounds < kMaxTileRounds && mat_index + ((size_t)1 << rplan[r].vq.size()) <= (size_t)kMaxTileMats) {
      const RoundPlan& rp = rplan[r];
      const int nv = (int)rp.vq.size();
      build_round_matrices(all, plan, rp, reinterpret_cast<cplx*>(A.mats + mat_index * kRoundMatAmps));
      const RoundDescHost hd = make_round_desc(rp, local_of, (uint32_t)mat_index);
      TileRoundDesc& rd = A.rounds[n_rounds];
      rd.rb = hd.rb;
      for (int w = 0; w < 3; ++w) rd.tb[w] = hd.tb[w];
      rd.var = hd.var;
      rd.mat_off = hd.mat_off;
      rd.pad[0] = rd.pad[1] = 0;
      mat_index += (size_t)1 << nv;
      n_ops += rp.ops.size();
      ++n_rounds;
      ++r;
    }
    A.n_rounds = n_rounds;
    A.n_mats = (int)mat_index;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (debug > 2) {  // QCSIM_DEBUG_PLAN=3: time every launch (serialises the stream; diagnostics only)
      CUDA_TRY(cudaEventCreate(&ev0));
      CUDA_TRY(cudaEventCreate(&ev1));
      CUDA_TRY(cudaEventRecord(ev0, h->stream));
    }
    k_tile_pipe<<<(unsigned)grid, kPipeThreads, kPipeSmemBytes, h->stream>>>(A);
    CUDA_TRY(cudaGetLastError());
    if (debug > 2) {
      CUDA_TRY(cudaEventRecord(ev1, h->stream));
      CUDA_TRY(cudaEventSynchronize(ev1));
      float ms = 0;
      CUDA_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
      int chained = 0;
      for (int q = 0; q < n_rounds; ++q) chained += (A.rounds[q].var >> 7) & 1;
      std::fprintf(stderr, "[qcsim pipe launch] rounds=%d chained=%d matrices=%zu enum=%d ms=%.3f\n", n_rounds, chained, mat_index, g.n_enum, ms);
#ifdef QCSIM_PIPE_PROFILE
      unsigned long long pr[16], zero[16] = {0};
      CUDA_TRY(cudaMemcpyFromSymbol(pr, g_pipe_prof, sizeof(pr)));
      CUDA_TRY(cudaMemcpyToSymbol(g_pipe_prof, zero, sizeof(zero)));
      const double w = (double)pr[7], cta = w / kPipeConsumerWarps;
      if (w > 0)
        std::fprintf(stderr, "[qcsim pipe profile] rounds=%d per consumer warp: total %.0f kcyc, wait tile %.1f%%, rounds %.1f%% (barriers %.1f%%), other %.1f%% | producer: total %.0f kcyc, wait done %.1f%%, wait store reads %.1f%%\n",
                     n_rounds, pr[0] / w / 1e3, 100.0 * pr[1] / pr[0], 100.0 * pr[2] / pr[0], 100.0 * pr[3] / pr[0], 100.0 * (pr[0] - pr[1] - pr[2]) / (double)pr[0],
                     pr[4] / cta / 1e3, 100.0 * pr[5] / pr[4], 100.0 * pr[6] / pr[4]);
#endif
      cudaEventDestroy(ev0);
      cudaEventDestroy(ev1);
    }
    h->stats.kernel_launches += 1;
    h->stats.state_passes += 1;
    h->stats.bytes_moved += 32ULL * h->dim_local;
    h->stats.fused_rounds += n_rounds;
    h->stats.fused_ops += n_ops;
    if (debug > 1)
      std::fprintf(stderr, "[qcsim pipe pass] box=2^%d amps x %d ops, ops=%zu rounds=%d matrices=%zu\n", g.box_log2, 1 << g.n_enum, n_ops,
                   n_rounds, mat_index);
  } while (r < rplan.size());
  return QCSIM_OK;
}

// Build the round matrices + parameter block of one pass and launch it.  A launch holds at most
// kMaxTileRounds rounds / kMaxTileMats matrices (parameter-block limit); a longer pass is cut into
// several launches over the same tile set (each is still fp64-bound, not HBM-bound, at that length).
static int launch_pass(qcsim_sv* h, const std::vector<Op>& all, const PassPlan& plan, int L, const std::vector<RoundPlan>& rplan) {
  const int k = (int)plan.tile.size();
  int local_of[64];
  for (int q = 0; q < 64; ++q) local_of[q] = -1;
  for (int j = 0; j < k; ++j) local_of[plan.tile[j]] = j;

  static thread_local TilePassArgs A;  // ~30 KiB: keep it off the stack; the launch copies it
  A.k = k;
  A.low_identity = L;
  A.n_tiles = 1ULL << (h->n_local - k);
  for (int j = 0; j < kMaxTileBits; ++j) A.tpos[j] = j < k ? plan.tile[j] : 0;
  static const int want_pipe = env_int("QCSIM_TILE_PIPE", 0);
  const bool pipe = want_pipe && (k == kMaxTileBits);  // double-buffered tiles, one CTA per SM
  A.pipelined = pipe ? 1 : 0;
  const size_t tile_smem = ((size_t)sizeof(amp) << k) * (pipe ? 2 : 1);
  const size_t smem_max = tile_smem + (size_t)kMaxTileMats * kRoundMatAmps * sizeof(amp);
  const int per_sm = pipe ? 1 : std::max(1, std::min(2, (int)((220 * 1024) / (smem_max + 1024))));
  const uint64_t grid = std::min<uint64_t>(A.n_tiles, (uint64_t)kNumSMs * per_sm);
  static const int debug = env_int("QCSIM_DEBUG_PLAN", 0);

  size_t r = 0;
  while (r < rplan.size()) {
    int n_rounds = 0;
    size_t mat_index = 0, n_ops = 0;
    while (r < rplan.size() && n_rounds < kMaxTileRounds && mat_index + ((size_t)1 << rplan[r].vq.size()) <= (size_t)kMaxTileMats) {
      const RoundPlan& rp = rplan[r];
      const int nv = (int)rp.vq.size();
      // one 8x8 matrix per value of the variant qubits (planner.h), straight into the parameter block
      build_round_matrices(all, plan, rp, reinterpret_cast<cplx*>(A.mats + mat_index * kRoundMatAmps));
      const RoundDescHost hd = make_round_desc(rp, local_of, (uint32_t)mat_index);
      TileRoundDesc& rd = A.rounds[n_rounds];
      rd.rb = hd.rb;
      for (int w = 0; w < 3; ++w) rd.tb[w] = hd.tb[w];
      rd.var = hd.var;
      rd.mat_off = hd.mat_off;
      rd.pad[0] = rd.pad[1] = 0;
      mat_index += (size_t)1 << nv;
      n_ops += rp.ops.size();
      ++n_rounds;
      ++r;
    }
    A.n_rounds = n_rounds;
    A.n_mats = (int)mat_index;
    const size_t smem = tile_smem + mat_index * kRoundMatAmps * sizeof(amp);
    k_tile_pass<<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    h->stats.state_passes += 1;
    h->stats.bytes_moved += 32ULL * h->dim_local;
    h->stats.fused_rounds += n_rounds;
    h->stats.fused_ops += n_ops;
    if (debug > 1)
      std::fprintf(stderr, "[qcsim pass] k=%d L=%d ops=%zu rounds=%d matrices=%zu\n", k, L, n_ops, n_rounds, mat_index);
  }
  return QCSIM_OK;
}

static int execute_plain(qcsim_sv* h, const std::vector<Op>& ops) {
  if (ops.empty()) return QCSIM_OK;
  if (h->world > 1) return dist_execute(h, ops);
  return fusion_execute_local(h, ops);
}

int fusion_execute(qcsim_sv* h, const std::vector<Op>& ops_in) { return fusion_execute_partial(h, ops_in, nullptr); }

// true when the list holds a QFT / IQFT gate stream (or the start of one): such a list must be flushed whole, a
// subset of it would no longer be recognised and would run gate by gate
bool fusion_holds_qft(const std::vector<Op>& ops) {
  if (ops.size() < 10) return false;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (!(ops[i].n_ctrl == 0 && ops[i].kind == OP_PAIR)) continue;
    bool ran_off = false;
    const QftMatch m = match_qft(ops, i, 4, &ran_off);
    if (m.length > 0 || ran_off) return true;
  }
  return false;
}

// `deferred` != nullptr: a trailing run of ops that is still a valid QFT prefix when the list ends is
// handed back instead of executed (the caller keeps it queued until the rest of the transform arrives).
int fusion_execute_partial(qcsim_sv* h, const std::vector<Op>& ops_in, std::vector<Op>* deferred) {
  // QFT / IQFT gate streams (QCSim's own QuantumFourierTransform emits them gate by gate) run as
  // radix-8 FFT passes; everything around them goes through the gate-block planner
  static const int no_qft = env_int("QCSIM_QFT_GENERIC", 0);
  if (no_qft || ops_in.size() < 10) return execute_plain(h, ops_in);
  std::vector<Op> plain;
  size_t i = 0;
  while (i < ops_in.size()) {
    const Op& op = ops_in[i];
    const bool candidate = op.n_ctrl == 0 && op.kind == OP_PAIR;  // a pattern starts with H or SWAP
    QftMatch m;
    bool ran_off = false;
    if (candidate) m = match_qft(ops_in, i, 4, &ran_off);
    const bool touches_end = i + m.length == ops_in.size();  // more of the transform (or its swaps) may still arrive
    if (m.length > 0 && !(deferred && touches_end)) {
      QCSIM_TRY(execute_plain(h, plain));
      plain.clear();
      QCSIM_TRY(engine_qft_direct(h, m.sq, m.eq, m.do_swap, m.inverse));
      i += m.length;
    } else if (deferred && (ran_off || m.length > 0)) {
      deferred->assign(ops_in.begin() + i, ops_in.end());
      break;
    } else {
      plain.push_back(op);
      ++i;
    }
  }
  return execute_plain(h, plain);
}

// All qubit indices in `ops` are physical bit positions of the local slice.
int fusion_execute_local(qcsim_sv* h, const std::vector<Op>& ops) {
  NvtxRange nvtx_range("qcsim.gate_blocks");
  const int nl = h->n_local;
  const int N = (int)ops.size();
  static const int K_env = env_int("QCSIM_TILE_BITS", kMaxTileBits);
  static const int L_env = env_int("QCSIM_TILE_LOW", 4);
  static const int no_fuse = env_int("QCSIM_NO_FUSION", 0);
  static const int legacy = env_int("QCSIM_TILE_LEGACY", 0);
  // TMA-staged pass (tile_pipe.cuh): 2^11-amplitude tiles whose innermost TMA box is qubits 0..2 (128 B);
  // small registers and QCSIM_TILE_LEGACY=1 use the plain tile pass (2^12 tiles, 256 B runs)
  const bool pipe = !legacy && nl >= kPipeTileBits && tensor_map_encoder() != nullptr;
  const int K = pipe ? kPipeTileBits : std::max(kRoundBits, std::min({K_env, kMaxTileBits, nl}));
  const int L = pipe ? 3 : std::max(1, std::min(L_env, K - kRoundBits));
  if (no_fuse || nl < 6 || N < 2) {
    for (const Op& op : ops) QCSIM_TRY(engine_launch_local(h, op));
    return QCSIM_OK;
  }

  const std::vector<PlanStep> steps = plan_passes(ops, nl, K, L, 1 << 20, 1 << 30);
  static const int debug = env_int("QCSIM_DEBUG_PLAN", 0);
  if (debug) {
    int nf = 0, absorbed = 0;
    for (const PlanStep& st : steps)
      if (st.fused) {
        ++nf;
        absorbed += (int)st.pass.ops.size();
      }
    std::fprintf(stderr, "[qcsim plan] %d ops -> %zu steps (%d fused passes holding %d ops), K=%d L=%d\n", N, steps.size(), nf,
                 absorbed, K, L);
  }
  for (const PlanStep& st : steps) {
    if (!st.fused) {
      QCSIM_TRY(engine_launch_local(h, ops[st.pass.ops[0]]));
      continue;
    }
    int Lrun = 0;  // the low run of identity-mapped tile bits may be longer than L
    while (Lrun < (int)st.pass.tile.size() && st.pass.tile[Lrun] == Lrun) ++Lrun;
    // A round costs fp64-pipe time worth ~24 B of HBM traffic per amplitude; a pass whose rounds cost
    // more than running its gates one by one (long runs of cheap diagonal gates) is not fused
    const std::vector<RoundPlan> rplan = schedule_rounds(ops, st.pass, kMaxVariantBits);
    size_t n_mats = 0;
    for (const RoundPlan& rp : rplan) n_mats += (size_t)1 << rp.vq.size();
    const double launches = std::max(std::ceil(rplan.size() / (double)kMaxTileRounds), std::ceil(n_mats / (double)kMaxTileMats));
    const double cost_fused = std::max(32.0 * launches, 24.0 * rplan.size());
    double cost_alone = 0;
    for (int idx : st.pass.ops) cost_alone += standalone_cost(ops[idx]);
    if (cost_fused >= cost_alone) {
      for (int idx : st.pass.ops) QCSIM_TRY(engine_launch_local(h, ops[idx]));
      continue;
    }
    int rc = QCSIM_ERR_UNSUPPORTED;
    if (pipe) rc = launch_pass_pipe(h, ops, st.pass);          // TMA-staged warp-specialised pass (tile_pipe.cuh)
    if (rc == QCSIM_ERR_UNSUPPORTED) rc = launch_pass(h, ops, st.pass, Lrun, rplan);  // small registers: plain tile pass
    QCSIM_TRY(rc);
  }
  return QCSIM_OK;
}


}  // namespace qcsim
