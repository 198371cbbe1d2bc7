// dist.cu -- sharded registers: one process per GPU, state split on the top log2(world) physical
// index bits.  Planning (which steps are shard-local, when to exchange global and local qubits)
// is in dist_plan.h; this file executes the plan:
//   * local steps  -> fusion_execute_local (the same fused / single-gate kernels as one GPU);
//   * exchange     -> ncclSend/ncclRecv over NVLink among the 2^k ranks that differ in the swapped
//                     global bits.  The swapped local bits are the top k local positions, so every
//                     peer's share is one contiguous block: it is sent straight out of the state and
//                     received into a small double-buffered staging area (the state itself leaves no
//                     room for a second copy at 33+ qubits per GPU), then copied into place on a
//                     second stream while the next chunk is on the wire;
//   * reductions   -> one double(-double) per rank, all-gathered with NCCL, resolved identically
//                     on every rank.
#include <nccl.h>

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "dist.h"
#include "dist_plan.h"
#include "fusion.h"
#include "qft_kernels.cuh"
#include "reduce_kernels.cuh"

namespace qcsim {

#define NCCL_TRY(expr)                                                                            \
  do {                                                                                            \
    const ncclResult_t nr__ = (expr);                                                             \
    if (nr__ != ncclSuccess) return fail(QCSIM_ERR_NCCL, "%s: %s", #expr, ncclGetErrorString(nr__)); \
  } while (0)

namespace {

static uint64_t stage_chunk_amps() {  // amps per peer per buffer (default 2^22 = 64 MiB)
  static const uint64_t v = [] {
    const char* s = std::getenv("QCSIM_EXCHANGE_CHUNK_LOG2");
    const int l = s ? std::atoi(s) : 22;
    return 1ULL << std::max(4, std::min(l, 32));
  }();
  return v;
}

struct DistState {
  ncclComm_t comm = nullptr;
  DistLayout layout;
  cudaStream_t copy_stream = nullptr;
  amp* stage = nullptr;
  uint64_t stage_chunk = 0;  // amps per (peer, buffer) slot
  int stage_peers = 0;
  cudaEvent_t ev_group[2] = {nullptr, nullptr};
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  double* d_small = nullptr;  // 256 doubles
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
  // peer-memory path: every rank's slice mapped into this process through CUDA IPC
  bool p2p = false;
  amp* peer_psi[64] = {};
  bool peer_is_ipc[64] = {};  // mapped through CUDA IPC (another process) -- same-process peers are plain pointers
  // SPMD discipline: every rank must make the same collective-bearing calls in the same order.
  // seq_hash folds (kind, arguments) of each such call; with QCSIM_SPMD_CHECK=1 the hashes are
  // compared across ranks before every collective and a mismatch is an error instead of a hang.
  uint64_t seq_hash = 0x9e3779b97f4a7c15ULL;
  uint64_t seq_count = 0;
  bool broken = false;  // a collective timed out / mismatched: the communicator was aborted
};

DistState* st(qcsim_sv* h) { return static_cast<DistState*>(h->dist); }

int log2i(int w) {
  int l = 0;
  while ((1 << l) < w) ++l;
  return l;
}

int env_flag(const char* name) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : 0;
}

double collective_timeout_s() {
  static const double v = [] {
    const char* s = std::getenv("QCSIM_COLLECTIVE_TIMEOUT_S");
    const double t = s ? std::atof(s) : 300.0;
    return t > 0 ? t : 300.0;
  }();
  return v;
}

}  // namespace

// Bounded wait: a stream that holds a collective whose peers never arrive (SPMD violation: the ranks
// made different calls) would otherwise spin forever inside NCCL.
int dist_wait(qcsim_sv* h) {
  DistState* d = st(h);
  if (d && d->broken) return fail(QCSIM_ERR_NCCL, "sharded register is unusable after a failed collective");
  const auto t0 = std::chrono::steady_clock::now();
  int spins = 0;
  for (;;) {
    const cudaError_t ce = cudaStreamQuery(h->stream);
    if (ce == cudaSuccess) return QCSIM_OK;
    if (ce != cudaErrorNotReady) return fail(QCSIM_ERR_CUDA, "stream: %s", cudaGetErrorString(ce));
    if (++spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(50));
    if ((spins & 1023) == 0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (el > collective_timeout_s()) {
        if (d && d->comm) {
          ncclCommAbort(d->comm);
          d->comm = nullptr;
          h->nccl_comm = nullptr;
          d->broken = true;
        }
        return fail(QCSIM_ERR_NCCL,
                    "collective did not complete within %.0f s on rank %d after %llu collective calls: the ranks of a sharded "
                    "register must make the same calls in the same order (SPMD); set QCSIM_SPMD_CHECK=1 to locate the mismatch",
                    collective_timeout_s(), h->rank, (unsigned long long)(d ? d->seq_count : 0));
      }
    }
  }
}

// Folds one collective-bearing call into the sequence hash; in check mode compares it across ranks first.
static int spmd_note(qcsim_sv* h, uint64_t kind, uint64_t a = 0, uint64_t b = 0) {
  DistState* d = st(h);
  if (d->broken || !d->comm) return fail(QCSIM_ERR_NCCL, "sharded register is unusable after a failed collective");
  auto mix = [](uint64_t x, uint64_t y) {
    x ^= y + 0x9e3779b97f4a7c15ULL + (x << 6) + (x >> 2);
    x *= 0xff51afd7ed558ccdULL;
    return x ^ (x >> 33);
  };
  d->seq_hash = mix(mix(mix(d->seq_hash, kind), a), b);
  d->seq_count++;
  static const int check = env_flag("QCSIM_SPMD_CHECK");
  if (!check) return QCSIM_OK;
  // all-gather (hash, count) as exact 26-bit pieces in doubles; region [208, 208 + 4 * world) of d_small
  double mine[4] = {(double)(d->seq_hash & 0x3ffffffULL), (double)((d->seq_hash >> 26) & 0x3ffffffULL), (double)(d->seq_hash >> 52),
                    (double)d->seq_count};
  double all[4 * kMaxWorld];
  double* slot = d->d_small + 208;
  CUDA_TRY(cudaMemcpyAsync(slot + 4 * h->rank, mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(ncclAllGather(slot + 4 * h->rank, slot, 4, ncclDouble, d->comm, h->stream));
  CUDA_TRY(cudaMemcpyAsync(all, slot, sizeof(double) * 4 * h->world, cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(dist_wait(h));
  for (int r = 0; r < h->world; ++r)
    for (int j = 0; j < 4; ++j)
      if (all[4 * r + j] != mine[j]) {
        d->broken = true;
        return fail(QCSIM_ERR_NCCL, "SPMD violation: collective call #%llu (kind %llu) on rank %d does not match rank %d",
                    (unsigned long long)d->seq_count, (unsigned long long)kind, h->rank, r);
      }
  return QCSIM_OK;
}

static int setup_peers(qcsim_sv* h);
static void close_peers(qcsim_sv* h);
static int stream_barrier(qcsim_sv* h);

int dist_unique_id(void* out) {
  static_assert(sizeof(ncclUniqueId) == 128, "the ABI passes the NCCL id as 128 bytes");
  if (!out) return fail(QCSIM_ERR_BAD_ARG, "null output");
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  std::memcpy(out, &id, sizeof id);
  return QCSIM_OK;
}

int engine_nccl_unique_id(void* out) { return dist_unique_id(out); }

int dist_init(qcsim_sv* h, const void* nccl_id) {
  if (!nccl_id) return fail(QCSIM_ERR_BAD_ARG, "sharded register needs the NCCL unique id");
  if (h->n_local < 4) return fail(QCSIM_ERR_BAD_ARG, "a sharded register needs at least 4 local qubits per rank");
  DistState* d = new DistState();
  h->dist = d;
  d->layout.reset(h->n, h->n_local);
  ncclUniqueId id;
  std::memcpy(&id, nccl_id, sizeof id);
  NCCL_TRY(ncclCommInitRank(&d->comm, h->world, id, h->rank));
  h->nccl_comm = d->comm;
  CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_group[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&d->ev_copy[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaMalloc(&d->d_small, 256 * sizeof(double)));  // [0,200): reductions, [200]: barrier token, [208,240): SPMD check
  CUDA_TRY(cudaMemset(d->d_small, 0, 256 * sizeof(double)));
  return setup_peers(h);
}

void dist_shutdown(qcsim_sv* h) {
  DistState* d = st(h);
  if (!d) return;
  for (auto& p : d->timed) {
    cudaEventDestroy(p.first);
    cudaEventDestroy(p.second);
  }
  if (d->copy_stream) {
    cudaStreamSynchronize(d->copy_stream);
    cudaStreamDestroy(d->copy_stream);
  }
  for (int i = 0; i < 2; ++i) {
    if (d->ev_group[i]) cudaEventDestroy(d->ev_group[i]);
    if (d->ev_copy[i]) cudaEventDestroy(d->ev_copy[i]);
  }
  if (d->comm && d->p2p) {  // nobody may unmap a slice a peer is still swapping with
    stream_barrier(h);
    cudaStreamSynchronize(h->stream);
  }
  close_peers(h);
  cudaFree(d->stage);
  cudaFree(d->d_small);
  if (d->comm) ncclCommDestroy(d->comm);
  delete d;
  h->dist = nullptr;
  h->nccl_comm = nullptr;
}

void dist_reset_layout(qcsim_sv* h) {
  if (st(h)) st(h)->layout.reset(h->n, h->n_local);
}

int dist_buffers_changed(qcsim_sv* h) { return setup_peers(h); }

void dist_map_mask(qcsim_sv* h, uint64_t mask, uint64_t want, uint64_t* pmask, uint64_t* pwant) {
  const DistLayout& L = st(h)->layout;
  uint64_t pm = 0, pw = 0;
  for (int q = 0; q < h->n; ++q) {
    if ((mask >> q) & 1ULL) pm |= 1ULL << L.phys_of[q];
    if ((want >> q) & 1ULL) pw |= 1ULL << L.phys_of[q];
  }
  *pmask = pm;
  *pwant = pw;
}

// sum over ranks of `count` host doubles (count <= 256)
int dist_allreduce_host(qcsim_sv* h, double* vals, int count) {
  DistState* d = st(h);
  if (count > 200) return fail(QCSIM_ERR_BAD_ARG, "internal: allreduce too large");
  QCSIM_TRY(spmd_note(h, 1, (uint64_t)count));
  CUDA_TRY(cudaMemcpyAsync(d->d_small, vals, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(ncclAllReduce(d->d_small, d->d_small, count, ncclDouble, ncclSum, d->comm, h->stream));
  CUDA_TRY(cudaMemcpyAsync(vals, d->d_small, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  return QCSIM_OK;
}

// every rank contributes `per` doubles; all[r * per + i] afterwards (exact: zeros elsewhere)
static int allgather_host(qcsim_sv* h, const double* mine, int per, double* all) {
  double buf[256];
  const int total = per * h->world;
  if (total > 200) return fail(QCSIM_ERR_BAD_ARG, "internal: allgather too large");
  for (int i = 0; i < total; ++i) buf[i] = 0.0;
  for (int i = 0; i < per; ++i) buf[h->rank * per + i] = mine[i];
  QCSIM_TRY(dist_allreduce_host(h, buf, total));
  for (int i = 0; i < total; ++i) all[i] = buf[i];
  return QCSIM_OK;
}

// ---- peer-memory exchange kernel --------------------------------------------------------------------
// In-place swap of physical index bits gpos[j] (rank bits) with local bits lpos[j], straight over NVLink: no
// staging buffer, no second copy.  For rank r with value a on the swapped global bits and every t != a, the
// amplitudes of r whose local bits lpos hold t trade places with the amplitudes of peer(t) whose local bits lpos
// hold a.  Of every pair of ranks the lower one swaps the first half of the (strided) block and the higher one
// the second half, so each direction of each link carries the same load (half as remote stores issued here,
// half as remote loads issued by the peer).  The partner bits may sit anywhere in the local index: the block is a
// set of runs of 2^(lowest partner position) amplitudes, addressed like a controlled gate addresses its subspace.
struct SwapArgs {
  int n_peers;
  FixedBits fix;           // the local partner positions, ascending (removed from the work-item index)
  amp* remote[7];          // peer's slice
  uint64_t mine_or[7];     // partner bits set to t (my side of the trade with that peer)
  uint64_t remote_or[7];   // partner bits set to a (the peer's side)
  uint64_t first[7];       // my share of the block: work items [first, first + count)
  uint64_t count[7];       // in units of amp2 (two adjacent amplitudes, 32 bytes) -- or single amplitudes (k_exchange_swap_v1)
};

__global__ void __launch_bounds__(256) k_exchange_swap(amp* __restrict__ mine, const __grid_constant__ SwapArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (int p = 0; p < A.n_peers; ++p) {
    amp* const m = mine + A.mine_or[p];
    amp* const r = A.remote[p] + A.remote_or[p];
    const uint64_t lo = A.first[p], hi = A.first[p] + A.count[p];
    uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < hi; i += 4 * stride) {  // 4 x 32 B remote loads in flight per thread
      amp2 x[4], y[4];
      uint64_t o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) o[u] = scatter_index(2 * (i + u * stride), A.fix);
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = ld_amp2(r + o[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = ld_amp2(m + o[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        st_amp2(m + o[u], y[u]);
        st_amp2(r + o[u], x[u]);
      }
    }
    for (; i < hi; i += stride) {
      const uint64_t o = scatter_index(2 * i, A.fix);
      const amp2 y = ld_amp2(r + o), x = ld_amp2(m + o);
      st_amp2(m + o, y);
      st_amp2(r + o, x);
    }
  }
}

// a partner on local bit 0: single amplitudes (tiny registers only; the planner avoids low partner positions)
__global__ void __launch_bounds__(256) k_exchange_swap_v1(amp* __restrict__ mine, const __grid_constant__ SwapArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (int p = 0; p < A.n_peers; ++p) {
    amp* const m = mine + A.mine_or[p];
    amp* const r = A.remote[p] + A.remote_or[p];
    const uint64_t hi = A.first[p] + A.count[p];
    for (uint64_t i = A.first[p] + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
      const uint64_t o = scatter_index(i, A.fix);
      const amp y = r[o], x = m[o];
      m[o] = y;
      r[o] = x;
    }
  }
}

static void close_peers(qcsim_sv* h) {
  DistState* d = st(h);
  for (int r = 0; r < h->world; ++r) {
    if (d->peer_psi[r] && r != h->rank && d->peer_is_ipc[r]) cudaIpcCloseMemHandle(d->peer_psi[r]);
    d->peer_psi[r] = nullptr;
    d->peer_is_ipc[r] = false;
  }
  d->p2p = false;
}

// Collective: (re)map every rank's slice.  Falls back to the NCCL send/recv path if CUDA IPC is
// not usable here (QCSIM_EXCHANGE=nccl forces that path).
static int setup_peers(qcsim_sv* h) {
  DistState* d = st(h);
  close_peers(h);
  QCSIM_TRY(spmd_note(h, 2));
  const char* mode = std::getenv("QCSIM_EXCHANGE");
  const bool want = !(mode && std::strcmp(mode, "nccl") == 0);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  // per rank: 64-byte IPC handle | pid | raw device pointer | device ordinal  (11 doubles)
  struct PeerRecord {
    cudaIpcMemHandle_t ipc;
    int64_t pid;
    uint64_t ptr;
    int64_t device;
  };
  static_assert(sizeof(PeerRecord) == 88, "peer record layout");
  constexpr int kRec = sizeof(PeerRecord) / sizeof(double);
  if (h->world * kRec > 200) return fail(QCSIM_ERR_BAD_ARG, "internal: too many ranks for the handle exchange");
  PeerRecord mine;
  std::memset(&mine, 0, sizeof mine);
  mine.pid = (int64_t)getpid();
  mine.ptr = (uint64_t)(uintptr_t)h->psi;
  mine.device = h->device;
  double ok = want ? 1.0 : 0.0;
  if (want && cudaIpcGetMemHandle(&mine.ipc, h->psi) != cudaSuccess) {
    cudaGetLastError();  // not fatal yet: same-process peers do not need the IPC handle
    std::memset(&mine.ipc, 0, sizeof mine.ipc);
  }
  double* slot_mine = d->d_small + kRec * h->rank;
  CUDA_TRY(cudaMemcpyAsync(slot_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
  NCCL_TRY(ncclAllGather(slot_mine, d->d_small, kRec, ncclDouble, d->comm, h->stream));
  PeerRecord all[kMaxWorld];
  CUDA_TRY(cudaMemcpyAsync(all, d->d_small, sizeof(PeerRecord) * h->world, cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  if (ok != 0.0) {
    for (int r = 0; r < h->world; ++r) {
      d->peer_is_ipc[r] = false;
      if (r == h->rank) {
        d->peer_psi[r] = h->psi;
        continue;
      }
      if (all[r].pid == mine.pid) {
        // same process (qcsim_sv_create_multi: one worker thread per device): plain pointer + peer access
        const cudaError_t ce = cudaDeviceEnablePeerAccess((int)all[r].device, 0);
        if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          ok = 0.0;
          break;
        }
        cudaGetLastError();
        d->peer_psi[r] = (amp*)(uintptr_t)all[r].ptr;
        continue;
      }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[r].ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0.0;
        break;
      }
      d->peer_psi[r] = (amp*)ptr;
      d->peer_is_ipc[r] = true;
    }
  }
  double agree = ok;  // every rank must take the same path
  double buf[64];
  for (int r = 0; r < h->world; ++r) buf[r] = 0.0;
  buf[h->rank] = agree;
  QCSIM_TRY(dist_allreduce_host(h, buf, h->world));
  bool all_ok = true;
  for (int r = 0; r < h->world; ++r) all_ok = all_ok && buf[r] != 0.0;
  if (!all_ok) close_peers(h);
  d->p2p = all_ok;
  return QCSIM_OK;
}

static int stream_barrier(qcsim_sv* h) {  // stream-ordered barrier over all ranks
  DistState* d = st(h);
  NCCL_TRY(ncclAllReduce(d->d_small + 200, d->d_small + 200, 1, ncclDouble, ncclSum, d->comm, h->stream));
  return QCSIM_OK;
}

// ---- exchange: swap physical global positions gpos[j] with the top-k local positions -------------

static int ensure_stage(qcsim_sv* h, uint64_t chunk, int peers) {
  DistState* d = st(h);
  if (d->stage && d->stage_chunk >= chunk && d->stage_peers >= peers) return QCSIM_OK;
  QCSIM_TRY(engine_wait(h));
  CUDA_TRY(cudaStreamSynchronize(d->copy_stream));
  cudaFree(d->stage);
  d->stage = nullptr;
  const int all_peers = h->world - 1;
  CUDA_TRY(cudaMalloc(&d->stage, sizeof(amp) * chunk * 2 * all_peers));
  d->stage_chunk = chunk;
  d->stage_peers = all_peers;
  return QCSIM_OK;
}

// Resolve the timings of exchanges that have finished (all of them when `wait`), so the event list stays short.
static void harvest_timings(qcsim_sv* h, bool wait) {
  DistState* d = st(h);
  size_t keep = 0;
  for (size_t i = 0; i < d->timed.size(); ++i) {
    auto& p = d->timed[i];
    const bool done = wait ? cudaEventSynchronize(p.second) == cudaSuccess : cudaEventQuery(p.second) == cudaSuccess;
    if (!done && !wait) {
      d->timed[keep++] = p;
      continue;
    }
    float ms = 0;
    if (done && cudaEventElapsedTime(&ms, p.first, p.second) == cudaSuccess) h->stats.exchange_ms += ms;
    cudaEventDestroy(p.first);
    cudaEventDestroy(p.second);
  }
  d->timed.resize(keep);
  cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
}

static int do_exchange_nccl_top(qcsim_sv* h, const DistStep& ex);

static int do_exchange(qcsim_sv* h, const DistStep& ex) {
  NvtxRange nvtx_range("qcsim.exchange");
  DistState* d = st(h);
  const int k = ex.k, nl = h->n_local;
  if (k < 1 || k > kMaxExchange) return fail(QCSIM_ERR_BAD_ARG, "internal: bad exchange width");
  for (int j = 0; j < k; ++j)
    if (ex.gpos[j] < nl || ex.gpos[j] >= h->n || ex.lpos[j] < 0 || ex.lpos[j] >= nl) return fail(QCSIM_ERR_BAD_ARG, "internal: bad exchange positions");
  QCSIM_TRY(spmd_note(h, 3, (uint64_t)k,
                      (uint64_t)ex.gpos[0] | ((uint64_t)ex.gpos[1] << 8) | ((uint64_t)ex.gpos[2] << 16) | ((uint64_t)ex.lpos[0] << 24) |
                          ((uint64_t)ex.lpos[1] << 32) | ((uint64_t)ex.lpos[2] << 40)));
  harvest_timings(h, false);
  if (!d->p2p) {
    // NCCL send/recv moves contiguous blocks: park the partners on the top k local positions (one local
    // permutation pass), exchange there, and put them back
    bool top = true;
    for (int j = 0; j < k; ++j) top = top && ex.lpos[j] == nl - k + j;
    if (top) return do_exchange_nccl_top(h, ex);
    int src_of[64];
    for (int p = 0; p < 64; ++p) src_of[p] = p;
    for (int j = 0; j < k; ++j) {  // bring what sat on lpos[j] to the top slot nl-k+j (a product of position swaps)
      int pa = -1;
      for (int p = 0; p < nl; ++p)
        if (src_of[p] == ex.lpos[j]) pa = p;
      std::swap(src_of[pa], src_of[nl - k + j]);
    }
    int inv[64];
    for (int p = 0; p < 64; ++p) inv[p] = p;
    for (int p = 0; p < nl; ++p) inv[src_of[p]] = p;
    DistStep t = ex;
    for (int j = 0; j < k; ++j) t.lpos[j] = nl - k + j;
    QCSIM_TRY(engine_permute_bits(h, src_of));
    QCSIM_TRY(do_exchange_nccl_top(h, t));
    return engine_permute_bits(h, inv);
  }
  const uint64_t blk = h->dim_local >> k;  // amplitudes per sub-block
  const int n_sub = 1 << k;
  int a = 0;  // my value of the swapped global bits
  for (int j = 0; j < k; ++j) a |= ((h->rank >> (ex.gpos[j] - nl)) & 1) << j;
  auto peer_of = [&](int t) {
    int p = h->rank;
    for (int j = 0; j < k; ++j) {
      const int bit = 1 << (ex.gpos[j] - nl);
      p = ((t >> j) & 1) ? (p | bit) : (p & ~bit);
    }
    return p;
  };
  auto deposit = [&](int v) {  // value v of the k swapped bits -> local partner bits
    uint64_t o = 0;
    for (int j = 0; j < k; ++j)
      if ((v >> j) & 1) o |= 1ULL << ex.lpos[j];
    return o;
  };
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  SwapArgs A;
  std::memset(&A, 0, sizeof A);
  int sorted[kMaxExchange];
  for (int j = 0; j < k; ++j) sorted[j] = ex.lpos[j];
  std::sort(sorted, sorted + k);
  A.fix.n = k;
  for (int j = 0; j < 3; ++j) A.fix.pos[j] = j < k ? sorted[j] : 0;
  const bool wide = sorted[0] >= 1 && blk >= 4;  // 32-byte units
  const uint64_t units = wide ? blk / 2 : blk;
  // peers in XOR order: in step d every rank of the group talks to rank ^ d -- a perfect matching, so no rank is
  // the target of everybody at once
  for (int dd = 1; dd < n_sub; ++dd) {
    const int t = a ^ dd;
    const int peer = peer_of(t);
    const int j = A.n_peers++;
    A.remote[j] = d->peer_psi[peer];
    A.mine_or[j] = deposit(t);    // my amplitudes with partner bits = t ...
    A.remote_or[j] = deposit(a);  // ... trade with the peer's amplitudes with partner bits = a
    if (units < 2) {
      A.first[j] = 0;
      A.count[j] = h->rank < peer ? units : 0;
    } else {
      A.first[j] = h->rank < peer ? 0 : units / 2;
      A.count[j] = units / 2;
    }
  }
  QCSIM_TRY(stream_barrier(h));  // every rank has finished writing its slice
  CUDA_TRY(cudaEventRecord(e0, h->stream));
  if (wide) k_exchange_swap<<<kNumSMs * 8, 256, 0, h->stream>>>(h->psi, A);
  else k_exchange_swap_v1<<<kNumSMs * 8, 256, 0, h->stream>>>(h->psi, A);
  CUDA_TRY(cudaGetLastError());
  QCSIM_TRY(stream_barrier(h));  // every rank has finished swapping
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  d->timed.push_back({e0, e1});
  h->stats.kernel_launches += 1;
  h->stats.exchange_calls += 1;
  h->stats.exchange_bytes += (uint64_t)(n_sub - 1) * blk * sizeof(amp);
  return QCSIM_OK;
}

// NCCL send/recv path (QCSIM_EXCHANGE=nccl or no peer access): the partners are the top k local positions, so every
// peer's share is one contiguous block; double-buffered staging
static int do_exchange_nccl_top(qcsim_sv* h, const DistStep& ex) {
  DistState* d = st(h);
  const int k = ex.k, nl = h->n_local;
  const uint64_t blk = h->dim_local >> k;
  const uint64_t chunk = std::min<uint64_t>(blk, stage_chunk_amps());
  const int n_sub = 1 << k;
  int a = 0;
  for (int j = 0; j < k; ++j) a |= ((h->rank >> (ex.gpos[j] - nl)) & 1) << j;
  auto peer_of = [&](int t) {
    int p = h->rank;
    for (int j = 0; j < k; ++j) {
      const int bit = 1 << (ex.gpos[j] - nl);
      p = ((t >> j) & 1) ? (p | bit) : (p & ~bit);
    }
    return p;
  };
  auto slot = [&](int buf, int t) { return d->stage + ((uint64_t)(buf * (n_sub - 1) + (t < a ? t : t - 1))) * d->stage_chunk; };
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  QCSIM_TRY(ensure_stage(h, chunk, n_sub - 1));
  CUDA_TRY(cudaEventRecord(e0, h->stream));
  const uint64_t n_chunks = (blk + chunk - 1) / chunk;
  bool copied[2] = {false, false};
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const int buf = (int)(c & 1);
    const uint64_t off = c * chunk;
    const uint64_t cnt = std::min<uint64_t>(chunk, blk - off);
    if (copied[buf]) CUDA_TRY(cudaStreamWaitEvent(h->stream, d->ev_copy[buf], 0));  // staging slot drained
    NCCL_TRY(ncclGroupStart());
    for (int t = 0; t < n_sub; ++t) {
      if (t == a) continue;
      const int peer = peer_of(t);
      NCCL_TRY(ncclSend(h->psi + (uint64_t)t * blk + off, cnt * 2, ncclDouble, peer, d->comm, h->stream));
      NCCL_TRY(ncclRecv(slot(buf, t), cnt * 2, ncclDouble, peer, d->comm, h->stream));
    }
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaEventRecord(d->ev_group[buf], h->stream));
    CUDA_TRY(cudaStreamWaitEvent(d->copy_stream, d->ev_group[buf], 0));
    for (int t = 0; t < n_sub; ++t) {
      if (t == a) continue;
      CUDA_TRY(cudaMemcpyAsync(h->psi + (uint64_t)t * blk + off, slot(buf, t), cnt * sizeof(amp), cudaMemcpyDeviceToDevice,
                               d->copy_stream));
    }
    CUDA_TRY(cudaEventRecord(d->ev_copy[buf], d->copy_stream));
    copied[buf] = true;
  }
  for (int b = 0; b < 2; ++b)
    if (copied[b]) CUDA_TRY(cudaStreamWaitEvent(h->stream, d->ev_copy[b], 0));
  CUDA_TRY(cudaEventRecord(e1, h->stream));
  d->timed.push_back({e0, e1});
  h->stats.exchange_calls += 1;
  h->stats.exchange_bytes += (uint64_t)(n_sub - 1) * blk * sizeof(amp);
  return QCSIM_OK;
}

static int run_steps(qcsim_sv* h, const std::vector<DistStep>& steps) {
  for (const DistStep& s : steps) {
    if (s.exchange) QCSIM_TRY(do_exchange(h, s));
    else QCSIM_TRY(fusion_execute_local(h, s.ops));
  }
  return QCSIM_OK;
}

int dist_execute(qcsim_sv* h, const std::vector<Op>& ops) {
  DistState* d = st(h);
  return run_steps(h, dist_plan(d->layout, ops, h->rank));
}

int dist_apply(qcsim_sv* h, const Op& op) {
  std::vector<Op> one(1, op);
  return dist_execute(h, one);
}

int dist_canonicalize(qcsim_sv* h) {
  NvtxRange nvtx_range("qcsim.canonicalize_layout");
  DistState* d = st(h);
  if (d->layout.is_identity()) return QCSIM_OK;
  return run_steps(h, dist_plan_canonicalize(d->layout));
}

// QFT / IQFT on a sharded register (see engine_qft), in WHATEVER layout the register is in: the qubit reversal
// (QubitsSwapper, QubitsSwapper.h:23-34) is a relabelling, and nothing is canonicalised between chained
// transforms.  The targets are processed in the reference's order (QuantumFourierTransform.h:35-87: top-down
// for the QFT, bottom-up for the IQFT) as runs of qubits that sit on local positions -- radix-8 passes whose
// twiddles gather the lower qubits bit by bit, rank bits included.  When the next target sits on a global
// position, ONE exchange brings every unprocessed global target down, trading against (in this order) local
// qubits outside the transform, targets that are already done, targets that come last; partners may be any local
// position (k_exchange_swap), so there is no parking pass.
int dist_qft(qcsim_sv* h, int sq, int eq, bool do_swap, bool inverse, int* handled) {
  DistState* d = st(h);
  DistLayout& Lo = d->layout;
  const int nl = h->n_local, n = h->n;
  *handled = 1;
  auto virtual_swaps = [&]() {
    for (int s = sq, e = eq; s < e; ++s, --e) Lo.swap_logical(s, e);
  };
  if (inverse && do_swap) virtual_swaps();  // QubitsSwapper first (QuantumFourierTransform.h:67)
  const int step = inverse ? 1 : -1;
  int cur = inverse ? sq : eq;
  auto pending = [&](int q) { return inverse ? (q >= cur && q <= eq) : (q <= cur && q >= sq); };
  while (inverse ? cur <= eq : cur >= sq) {
    if (!Lo.is_global(cur)) {
      int end = cur;  // maximal run of local targets in processing order
      while ((inverse ? end + 1 <= eq : end - 1 >= sq) && !Lo.is_global(end + step)) end += step;
      QCSIM_TRY(engine_qft_passes(h, std::min(cur, end), std::max(cur, end), inverse, sq, Lo.phys_of));
      cur = end + step;
      continue;
    }
    // bring every pending global target down with one exchange
    std::vector<int> in_q;
    for (int q = cur; pending(q) && (int)in_q.size() < kMaxExchange; q += step)
      if (Lo.is_global(q)) in_q.push_back(q);
    struct Cand { int cls, order, pos, q; };
    std::vector<Cand> cand;
    const int min_pos = std::min(kMinExchangePos, std::max(0, nl - (n - nl) - 1));
    for (int q = 0; q < n; ++q) {
      if (Lo.is_global(q)) continue;
      Cand c;
      c.q = q;
      c.pos = Lo.phys_of[q];
      if (q < sq || q > eq) c.cls = 0, c.order = 0;                     // not part of the transform
      else if (!pending(q)) c.cls = 1, c.order = 0;                     // already transformed
      else c.cls = 2, c.order = inverse ? (eq - q) : (q - sq);          // pending: the one processed last goes first
      if (c.pos < min_pos) c.cls += 3;                                  // short runs over NVLink: last resort
      cand.push_back(c);
    }
    std::sort(cand.begin(), cand.end(), [](const Cand& a, const Cand& b) {
      if (a.cls != b.cls) return a.cls < b.cls;
      if (a.order != b.order) return a.order < b.order;
      return a.pos > b.pos;
    });
    if (cand.size() < in_q.size()) return fail(QCSIM_ERR_BAD_ARG, "internal: no local partner for the exchange");
    DistStep ex;
    ex.exchange = true;
    ex.k = (int)in_q.size();
    for (int j = 0; j < ex.k; ++j) {
      ex.gpos[j] = Lo.phys_of[in_q[j]];
      ex.lpos[j] = cand[j].pos;
    }
    QCSIM_TRY(do_exchange(h, ex));
    for (int j = 0; j < ex.k; ++j) Lo.swap_physical(ex.gpos[j], ex.lpos[j]);
  }
  if (!inverse && do_swap) virtual_swaps();
  return QCSIM_OK;
}

void dist_collect_stats(qcsim_sv* h) {
  if (!st(h)) return;
  harvest_timings(h, true);
}

// ---- measurement scan over all ranks (layout is canonical here; engine.cu: engine_resolve_draws) -------------

// exact (double-double) probability mass held by the ranks below this one; k_chunk_sums has been queued
int dist_scan_offset(qcsim_sv* h, dd* offset) {
  k_chunk_prefix<<<1, 1024, 0, h->stream>>>(h->d_chunk_sums, h->n_chunks, dd_make(0, 0), h->d_prefix_hi, h->d_total);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  dd* tot = (dd*)h->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(tot, h->d_total, sizeof(dd), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  double mine[2] = {tot->hi, tot->lo};
  double all[2 * kMaxWorld];
  QCSIM_TRY(allgather_host(h, mine, 2, all));
  dd off = dd_make(0, 0);
  for (int r = 0; r < h->rank; ++r) off = dd_add(off, dd_make(all[2 * r], all[2 * r + 1]));
  *offset = off;
  return QCSIM_OK;
}

// the reference's running sum continues from slice to slice: rank r walks its chunks starting from the sum rank r-1 ended with
int dist_chained_walk(qcsim_sv* h) {
  double acc = 0.0;
  double* stage = (double*)h->h_pinned;
  for (int r = 0; r < h->world; ++r) {
    double out = 0.0;
    if (r == h->rank) {
      k_sequential_walk<<<1, kWalkThreads, 0, h->stream>>>(h->psi, h->dim_local, h->n_chunks, h->d_prefix_hi, h->d_chunk_K, h->d_chunk_flags, acc,
                                                       h->d_acc_start, nullptr);
      CUDA_TRY(cudaGetLastError());
      h->stats.kernel_launches += 1;
      CUDA_TRY(cudaMemcpyAsync(stage, h->d_acc_start + h->n_chunks, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      QCSIM_TRY(engine_wait(h));
      out = *stage;
    }
    if (r + 1 < h->world) {
      QCSIM_TRY(dist_allreduce_host(h, &out, 1));  // only rank r contributed
      acc = out;
    }
  }
  return QCSIM_OK;
}

static __global__ void k_outcomes_to_global(unsigned long long* o, uint64_t count, unsigned long long base) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) o[i] = (o[i] == ~0ULL) ? 0ULL : (base | o[i]) + 1ULL;  // 0 = not in this slice
}
static __global__ void k_outcomes_finish(unsigned long long* o, uint64_t count) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) o[i] = o[i] == 0ULL ? ~0ULL : o[i] - 1ULL;
}

// every draw is resolved by at most one rank (the running sums of the slices do not overlap): max over ranks
int dist_combine_outcomes(qcsim_sv* h, unsigned long long* d_outcomes, uint64_t count) {
  DistState* d = st(h);
  QCSIM_TRY(spmd_note(h, 4, count));
  const unsigned grid = (unsigned)((count + 255) / 256);
  k_outcomes_to_global<<<grid, 256, 0, h->stream>>>(d_outcomes, count, (unsigned long long)h->rank << h->n_local);
  NCCL_TRY(ncclAllReduce(d_outcomes, d_outcomes, count, ncclUint64, ncclMax, d->comm, h->stream));
  k_outcomes_finish<<<grid, 256, 0, h->stream>>>(d_outcomes, count);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 2;
  return QCSIM_OK;
}

}  // namespace qcsim
