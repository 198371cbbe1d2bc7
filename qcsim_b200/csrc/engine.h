// engine.h -- the register object behind the C ABI and the host-side engine entry points.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/qcsim_b200.h"
#include "classify.h"
#include "common.cuh"

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: named ranges for nsys / ncu timelines

namespace qcsim {

// RAII NVTX range around an engine phase: qcsim.fused_block, qcsim.qft, qcsim.measure, qcsim.exchange ...
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

extern thread_local std::string g_last_error;
int fail(int code, const char* fmt, ...);

#define QCSIM_TRY(expr)              \
  do {                               \
    const int rc__ = (expr);         \
    if (rc__ != QCSIM_OK) return rc__; \
  } while (0)

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    const cudaError_t ce__ = (expr);                                                                \
    if (ce__ != cudaSuccess)                                                                        \
      return ::qcsim::fail(ce__ == cudaErrorMemoryAllocation ? QCSIM_ERR_OOM : QCSIM_ERR_CUDA, "%s: %s", #expr, \
                           cudaGetErrorString(ce__));                                               \
  } while (0)


}  // namespace qcsim

// The opaque handle of include/qcsim_b200.h.  Replaces the reference's registerStorage /
// resultsStorage / savedStateStorage host vectors (QubitRegister.h:718-721) with ONE device
// buffer (all kernels are in place) plus an optional saved copy.
struct qcsim_sv {
  int n = 0;             // total qubits
  int n_local = 0;       // qubits addressed inside this rank's slice
  uint64_t dim = 0;      // 2^n
  uint64_t dim_local = 0;
  int device = 0;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;

  qcsim::amp* psi = nullptr;
  qcsim::amp* saved = nullptr;

  // scratch for reductions / scans
  double* d_partials = nullptr;  // 2 * kMaxPartials doubles
  double* d_scalars = nullptr;   // small result area
  qcsim::dd* d_chunk_sums = nullptr;
  uint64_t n_chunks = 0;
  // measurement scan (reduce_kernels.cuh), allocated at the first measurement
  double* d_prefix_hi = nullptr;          // [n_chunks] exact mass before the chunk (binade predictor)
  unsigned long long* d_chunk_K = nullptr;  // [n_chunks] integer increments of the chunk
  int* d_chunk_flags = nullptr;           // [n_chunks]
  double* d_acc_start = nullptr;          // [n_chunks + 1] the reference's running sum at every chunk start
  qcsim::dd* d_total = nullptr;           // exact mass of the slice (+ offset)
  double* d_draws = nullptr;              // staging for draws / outcomes of a batch
  unsigned long long* d_outcomes = nullptr;
  uint64_t draws_capacity = 0;
  void* h_pinned = nullptr;      // 4 KiB pinned staging for scalar results
  qcsim::amp* d_qft_table = nullptr;  // per-pass item twiddle table of the QFT kernel (32 KiB)

  bool fusion = false;
  bool strict_measure = false;
  std::vector<qcsim::Op> queue;  // deferred gates (fusion mode / apply_batch)

  void* nccl_comm = nullptr;     // ncclComm_t when world > 1
  void* dist = nullptr;          // sharding state (dist.cu)
  void* multi = nullptr;         // single-process multi-device front (multi.cu): this handle owns one shard per device

  qcsim_stats stats = {};
};

namespace qcsim {

constexpr int kMaxWorld = 8;  // shards per register: one NVSwitch domain; exchange widths (dist_plan.h) are sized for log2 = 3
int engine_init_device_kernels();
int fusion_init_device_kernels();
// Wait for the handle's stream.  On a sharded register the stream may hold a collective whose peers
// never arrive (the ranks made different calls): the wait is bounded (QCSIM_COLLECTIVE_TIMEOUT_S,
// default 300 s) and ends in QCSIM_ERR_NCCL instead of spinning.
int engine_wait(qcsim_sv* h);
// Single-process multi-GPU (qcsim_sv_create_multi): the front handle owns one sharded register per device, each
// driven by its own worker thread; every API call is forwarded to all shards in lock step (multi.cu).
int multi_create(qcsim_sv** out, int n_qubits, int n_devices, const int* device_ids);
int multi_destroy(qcsim_sv* front);
int multi_world(const qcsim_sv* front);
qcsim_sv* multi_shard(const qcsim_sv* front, int rank);
// runs fn(shard, rank) on every worker thread at once; returns the first failing rank's code (its message becomes
// this thread's qcsim_last_error)
int multi_forward(qcsim_sv* front, const std::function<int(qcsim_sv*, int)>& fn);
int engine_create(qcsim_sv** out, int n_qubits, int device, int rank, int world, const void* nccl_id);
int engine_nccl_unique_id(void* out128);
int engine_destroy(qcsim_sv* h);
int engine_clone(const qcsim_sv* src, qcsim_sv** out);

int engine_set_basis_state(qcsim_sv* h, uint64_t state);
int engine_fill(qcsim_sv* h, double re, double im);
int engine_set_amplitude(qcsim_sv* h, uint64_t state, double re, double im);
int engine_get_amplitude(qcsim_sv* h, uint64_t state, double* re_im);
int engine_transfer(qcsim_sv* h, double* host, uint64_t first, uint64_t count, bool to_device);
int engine_masked_norm2(qcsim_sv* h, uint64_t mask, uint64_t want, double* out);
int engine_scale(qcsim_sv* h, double f);
int engine_collapse(qcsim_sv* h, uint64_t mask, uint64_t want, double f);
int engine_save(qcsim_sv* h);
int engine_restore(qcsim_sv* h, bool destructive);
int engine_inner_product(qcsim_sv* a, qcsim_sv* b, double* re_im);

int engine_apply_operator(qcsim_sv* h, const double* m_row_major);
int engine_apply_now(qcsim_sv* h, const Op& op);
int engine_enqueue(qcsim_sv* h, const Op& op);
int engine_flush(qcsim_sv* h);
int engine_flush_for_diagonal_observable(qcsim_sv* h, uint64_t qmask);  // gates the observable cannot see stay queued
void engine_drop_queue(qcsim_sv* h);
int engine_canonicalize(qcsim_sv* h);
int engine_qft(qcsim_sv* h, uint64_t sq, uint64_t eq, bool do_swap, bool inverse);
int engine_qft_passes(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of);
int engine_permute_bits(qcsim_sv* h, const int* src_of);
int engine_reverse_bits(qcsim_sv* h, int sq, int eq);
int engine_qft_direct(qcsim_sv* h, int sq, int eq, bool do_swap, bool inverse);

// outcomes[i] = first basis state whose running sum reaches probs[i] (the reference's sequential fp64 sum), or
// ~0 when the draw is beyond the total; exact for every draw
int engine_resolve_draws(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes);
int engine_pick_state(qcsim_sv* h, double prob, uint64_t fallback, uint64_t* outcome);
int engine_sample(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes);

}  // namespace qcsim
