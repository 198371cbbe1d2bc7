// engine.cu -- host side of the engine: owns the device buffers, lowers classified ops to
// kernel launches on the handle's stream.  Single-GPU paths live here; sharding is in dist.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "gate_kernels.cuh"
#include "reduce_kernels.cuh"
#include "dist.h"
#include "fusion.h"
#include "planner.h"
#include "qft_kernels.cuh"

namespace qcsim {

constexpr int kMaxPartials = kNumSMs * 8;

static inline int grid_for(uint64_t threads_needed) {
  uint64_t b = (threads_needed + kThreads - 1) / kThreads;
  if (b < 1) b = 1;
  if (b > (uint64_t)kMaxPartials) b = kMaxPartials;
  return (int)b;
}

static inline void count_pass(qcsim_sv* h, uint64_t amps_touched, int launches = 1) {
  h->stats.kernel_launches += launches;
  h->stats.state_passes += 1;
  h->stats.bytes_moved += 32ULL * amps_touched;
}

static inline amp to_amp(const cplx& z) { return make_amp(z.real(), z.imag()); }

int engine_wait(qcsim_sv* h) {
  if (h->world == 1) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return QCSIM_OK;
  }
  return dist_wait(h);
}

// ---- lifecycle ---------------------------------------------------------------------------------

// cudaFuncSetAttribute is per device (context): called from engine_create after cudaSetDevice, so a
// process that holds registers on several devices has it set on each of them
int engine_init_device_kernels() {
  CUDA_TRY(cudaFuncSetAttribute(k_qft_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_tile_permute, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CUDA_TRY(cudaFuncSetAttribute(k_bit_reverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(amp) << (2 * kRevT))));
  return QCSIM_OK;
}

int engine_create(qcsim_sv** out, int n_qubits, int device, int rank, int world, const void* nccl_id) {
  if (!out) return fail(QCSIM_ERR_BAD_ARG, "null output handle");
  *out = nullptr;
  if (n_qubits < 1 || n_qubits > 48) return fail(QCSIM_ERR_BAD_ARG, "n_qubits must be in 1..48");
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world)
    return fail(QCSIM_ERR_BAD_ARG, "world must be a power of two and 0 <= rank < world");
  if (world > kMaxWorld) return fail(QCSIM_ERR_BAD_ARG, "at most %d shards (one NVSwitch domain) are supported", kMaxWorld);
  int log2w = 0;
  while ((1 << log2w) < world) ++log2w;
  if (n_qubits - log2w < 1) return fail(QCSIM_ERR_BAD_ARG, "too few qubits for this many shards");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(QCSIM_ERR_CUDA, "no CUDA device (%s); qcsim_b200 has no CPU fallback",
                ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(QCSIM_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  // opt-in shared-memory sizes are per device: set them for this device now (idempotent)
  QCSIM_TRY(engine_init_device_kernels());
  QCSIM_TRY(fusion_init_device_kernels());

  qcsim_sv* h = new qcsim_sv();
  h->n = n_qubits;
  h->n_local = n_qubits - log2w;
  h->dim = 1ULL << n_qubits;
  h->dim_local = 1ULL << h->n_local;
  h->device = device;
  h->rank = rank;
  h->world = world;
  auto bail = [&](int rc) {
    engine_destroy(h);
    return rc;
  };
#define CREATE_TRY(expr)                                                                             \
  do {                                                                                               \
    const cudaError_t ce__ = (expr);                                                                 \
    if (ce__ != cudaSuccess)                                                                         \
      return bail(fail(ce__ == cudaErrorMemoryAllocation ? QCSIM_ERR_OOM : QCSIM_ERR_CUDA, "%s: %s", #expr, \
                       cudaGetErrorString(ce__)));                                                   \
  } while (0)
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaMalloc(&h->psi, h->dim_local * sizeof(amp)));
  CREATE_TRY(cudaMalloc(&h->d_partials, 2 * kMaxPartials * sizeof(double)));
  CREATE_TRY(cudaMalloc(&h->d_scalars, 64 * sizeof(double)));
  h->n_chunks = (h->dim_local + kChunk - 1) / kChunk;
  CREATE_TRY(cudaMalloc(&h->d_chunk_sums, h->n_chunks * sizeof(dd)));
  CREATE_TRY(cudaMalloc(&h->d_total, sizeof(dd)));
  CREATE_TRY(cudaMallocHost(&h->h_pinned, 4096));
#undef CREATE_TRY
  if (world > 1) {
    const int rc = dist_init(h, nccl_id);
    if (rc != QCSIM_OK) return bail(rc);
  }
  const int rc = engine_set_basis_state(h, 0);  // QubitRegister.h:36
  if (rc != QCSIM_OK) return bail(rc);
  *out = h;
  return QCSIM_OK;
}

int engine_destroy(qcsim_sv* h) {
  if (!h) return QCSIM_OK;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->world > 1) dist_shutdown(h);
  cudaFree(h->psi);
  cudaFree(h->saved);
  cudaFree(h->d_partials);
  cudaFree(h->d_scalars);
  cudaFree(h->d_chunk_sums);
  cudaFree(h->d_total);
  cudaFree(h->d_prefix_hi);
  cudaFree(h->d_chunk_K);
  cudaFree(h->d_chunk_flags);
  cudaFree(h->d_acc_start);
  cudaFree(h->d_draws);
  cudaFree(h->d_outcomes);
  cudaFree(h->d_qft_table);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QCSIM_OK;
}

int engine_clone(const qcsim_sv* src, qcsim_sv** out) {
  if (src->world > 1) return fail(QCSIM_ERR_UNSUPPORTED, "clone of a sharded register is not supported");
  qcsim_sv* s = const_cast<qcsim_sv*>(src);
  CUDA_TRY(cudaSetDevice(s->device));
  QCSIM_TRY(engine_flush(s));
  qcsim_sv* h = nullptr;
  QCSIM_TRY(engine_create(&h, src->n, src->device, 0, 1, nullptr));
  h->fusion = src->fusion;
  h->strict_measure = src->strict_measure;
  cudaError_t ce = cudaStreamSynchronize(s->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->psi, s->psi, s->dim_local * sizeof(amp), cudaMemcpyDeviceToDevice, h->stream);
  if (ce == cudaSuccess && s->saved) {
    ce = cudaMalloc(&h->saved, s->dim_local * sizeof(amp));
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(h->saved, s->saved, s->dim_local * sizeof(amp), cudaMemcpyDeviceToDevice, h->stream);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  if (ce != cudaSuccess) {
    engine_destroy(h);
    return fail(ce == cudaErrorMemoryAllocation ? QCSIM_ERR_OOM : QCSIM_ERR_CUDA, "clone: %s", cudaGetErrorString(ce));
  }
  *out = h;
  return QCSIM_OK;
}

// ---- state setters / getters -------------------------------------------------------------------

__global__ void k_set_amp(amp* psi, uint64_t idx, amp v) { psi[idx] = v; }

int engine_fill(qcsim_sv* h, double re, double im) {
  if (h->world > 1) dist_reset_layout(h);
  if (re == 0.0 && im == 0.0) {
    CUDA_TRY(cudaMemsetAsync(h->psi, 0, h->dim_local * sizeof(amp), h->stream));
  } else {
    k_fill<<<grid_for(h->dim_local), kThreads, 0, h->stream>>>(h->psi, h->dim_local, make_amp(re, im));
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches++;
  }
  return QCSIM_OK;
}

int engine_set_basis_state(qcsim_sv* h, uint64_t state) {
  QCSIM_TRY(engine_fill(h, 0, 0));  // Clear(), QubitRegister.h:78
  const uint64_t owner = state >> h->n_local;
  if (owner == (uint64_t)h->rank) {
    k_set_amp<<<1, 1, 0, h->stream>>>(h->psi, state & (h->dim_local - 1), make_amp(1, 0));
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches++;
  }
  return QCSIM_OK;
}

int engine_set_amplitude(qcsim_sv* h, uint64_t state, double re, double im) {
  QCSIM_TRY(engine_canonicalize(h));
  if ((state >> h->n_local) != (uint64_t)h->rank) return QCSIM_OK;  // other rank's element
  k_set_amp<<<1, 1, 0, h->stream>>>(h->psi, state & (h->dim_local - 1), make_amp(re, im));
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches++;
  return QCSIM_OK;
}

int engine_get_amplitude(qcsim_sv* h, uint64_t state, double* re_im) {
  QCSIM_TRY(engine_canonicalize(h));
  double* stage = (double*)h->h_pinned;
  stage[0] = stage[1] = 0;
  if ((state >> h->n_local) == (uint64_t)h->rank)
    CUDA_TRY(cudaMemcpyAsync(stage, h->psi + (state & (h->dim_local - 1)), sizeof(amp), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  if (h->world > 1) QCSIM_TRY(dist_allreduce_host(h, stage, 2));
  re_im[0] = stage[0];
  re_im[1] = stage[1];
  return QCSIM_OK;
}

int engine_transfer(qcsim_sv* h, double* host, uint64_t first, uint64_t count, bool to_device) {
  QCSIM_TRY(engine_canonicalize(h));
  const uint64_t base = (uint64_t)h->rank << h->n_local;
  if (count == 0) return QCSIM_OK;
  if (first < base || first + count > base + h->dim_local || first + count < first)
    return fail(QCSIM_ERR_BAD_ARG, "range [%llu, +%llu) is outside this rank's slice", (unsigned long long)first,
                (unsigned long long)count);
  amp* d = h->psi + (first - base);
  if (to_device)
    CUDA_TRY(cudaMemcpyAsync(d, host, count * sizeof(amp), cudaMemcpyHostToDevice, h->stream));
  else
    CUDA_TRY(cudaMemcpyAsync(host, d, count * sizeof(amp), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  return QCSIM_OK;
}

int engine_masked_norm2(qcsim_sv* h, uint64_t mask, uint64_t want, double* out) {
  NvtxRange nvtx_range("qcsim.probability_reduction");
  // logical (mask, want) -> physical bit positions of the current layout
  uint64_t pmask = mask, pwant = want;
  if (h->world > 1) dist_map_mask(h, mask, want, &pmask, &pwant);
  int g = grid_for(h->dim_local / 2);
  const uint64_t base = (uint64_t)h->rank << h->n_local;
  // contiguous selector range inside the local index (GetQubitProbability, Measure on an unsharded register): only the
  // selected subspace is read
  const int first = pmask ? __builtin_ctzll(pmask) : 0;
  const int width = __builtin_popcountll(pmask);
  const bool contiguous = pmask != 0 && (pmask >> first) == ((1ULL << width) - 1ULL) && first + width <= h->n_local;
  if (contiguous) {
    const uint64_t n_sub = h->dim_local >> width;
    g = grid_for(n_sub / 2 + 1);
    k_subspace_norm2<<<g, kThreads, 0, h->stream>>>(h->psi, n_sub, first, width, pwant & pmask, h->d_partials);
    h->stats.bytes_moved += 16ULL * n_sub;
  } else {
    k_masked_norm2<<<g, kThreads, 0, h->stream>>>(h->psi, h->dim_local, base, pmask, pwant, h->d_partials);
    h->stats.bytes_moved += 16ULL * h->dim_local;
  }
  k_final_sum<<<1, kThreads, 0, h->stream>>>(h->d_partials, g, 1, h->d_scalars);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 2;
  h->stats.state_passes += 1;
  double* stage = (double*)h->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(stage, h->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  if (h->world > 1) QCSIM_TRY(dist_allreduce_host(h, stage, 1));
  *out = stage[0];
  return QCSIM_OK;
}

int engine_scale(qcsim_sv* h, double f) {
  k_scale<<<grid_for(h->dim_local), kThreads, 0, h->stream>>>(h->psi, h->dim_local, f);
  CUDA_TRY(cudaGetLastError());
  count_pass(h, h->dim_local);
  return QCSIM_OK;
}

int engine_collapse(qcsim_sv* h, uint64_t mask, uint64_t want, double f) {
  NvtxRange nvtx_range("qcsim.collapse");
  uint64_t pmask = mask, pwant = want;
  if (h->world > 1) dist_map_mask(h, mask, want, &pmask, &pwant);
  const uint64_t base = (uint64_t)h->rank << h->n_local;
  k_collapse<<<grid_for(h->dim_local), kThreads, 0, h->stream>>>(h->psi, h->dim_local, base, pmask, pwant, f);
  CUDA_TRY(cudaGetLastError());
  count_pass(h, h->dim_local);
  return QCSIM_OK;
}

int engine_save(qcsim_sv* h) {
  QCSIM_TRY(engine_canonicalize(h));
  if (!h->saved) CUDA_TRY(cudaMalloc(&h->saved, h->dim_local * sizeof(amp)));
  CUDA_TRY(cudaMemcpyAsync(h->saved, h->psi, h->dim_local * sizeof(amp), cudaMemcpyDeviceToDevice, h->stream));
  return QCSIM_OK;
}

int engine_restore(qcsim_sv* h, bool destructive) {
  if (!h->saved) return QCSIM_OK;  // QubitRegister.h:607,613
  if (h->world > 1) dist_reset_layout(h);
  if (destructive) {
    QCSIM_TRY(engine_wait(h));
    std::swap(h->psi, h->saved);
    CUDA_TRY(cudaFree(h->saved));
    h->saved = nullptr;
    if (h->world > 1) QCSIM_TRY(dist_buffers_changed(h));
  } else {
    CUDA_TRY(cudaMemcpyAsync(h->psi, h->saved, h->dim_local * sizeof(amp), cudaMemcpyDeviceToDevice, h->stream));
  }
  return QCSIM_OK;
}

int engine_inner_product(qcsim_sv* a, qcsim_sv* b, double* re_im) {
  if (a->n != b->n || a->world != b->world || a->device != b->device)
    return fail(QCSIM_ERR_BAD_ARG, "registers must have the same shape and device");
  QCSIM_TRY(engine_canonicalize(a));
  QCSIM_TRY(engine_canonicalize(b));
  QCSIM_TRY(engine_wait(b));
  const int g = grid_for(a->dim_local);
  k_inner_product<<<g, kThreads, 0, a->stream>>>(a->psi, b->psi, a->dim_local, a->d_partials);
  k_final_sum<<<1, kThreads, 0, a->stream>>>(a->d_partials, g, 2, a->d_scalars);
  CUDA_TRY(cudaGetLastError());
  a->stats.kernel_launches += 2;
  double* stage = (double*)a->h_pinned;
  CUDA_TRY(cudaMemcpyAsync(stage, a->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, a->stream));
  QCSIM_TRY(engine_wait(a));
  if (a->world > 1) QCSIM_TRY(dist_allreduce_host(a, stage, 2));
  re_im[0] = stage[0];
  re_im[1] = stage[1];
  return QCSIM_OK;
}

// ---- gate application: one kernel per op ---------------------------------------------------------

static FixedBits sorted_bits(const int* a, int na, const int* b, int nb) {
  FixedBits f;
  f.n = 0;
  f.pos[0] = f.pos[1] = f.pos[2] = 0;
  int tmp[6];
  int n = 0;
  for (int i = 0; i < na; ++i) tmp[n++] = a[i];
  for (int i = 0; i < nb; ++i) tmp[n++] = b[i];
  std::sort(tmp, tmp + n);
  for (int i = 0; i < n && i < 3; ++i) f.pos[f.n++] = tmp[i];
  return f;
}

// Launch `op` on the local slice.  All qubit indices in `op` are PHYSICAL local bit positions.
int engine_launch_local(qcsim_sv* h, const Op& op) {
  const int nl = h->n_local;
  amp* psi = h->psi;
  uint64_t or_ctrl = 0;
  for (int i = 0; i < op.n_ctrl; ++i) or_ctrl |= 1ULL << op.ctrl[i];
  switch (op.kind) {
    case OP_NOP: return QCSIM_OK;
    case OP_PAIR: {
      PairArgs A;
      A.fix = sorted_bits(op.ctrl, op.n_ctrl, op.tgt, op.n_tgt);
      if (op.n_tgt == 1) {
        A.or_lo = or_ctrl;
        A.or_hi = or_ctrl | (1ULL << op.tgt[0]);
      } else {
        A.or_lo = or_ctrl | (1ULL << op.tgt[0]);
        A.or_hi = or_ctrl | (1ULL << op.tgt[1]);
      }
      A.m00 = to_amp(op.m[0]);
      A.m01 = to_amp(op.m[1]);
      A.m10 = to_amp(op.m[2]);
      A.m11 = to_amp(op.m[3]);
      A.n_items = 1ULL << (nl - A.fix.n);
      if (op.n_tgt == 1 && op.tgt[0] == 0) {
        k_pair_q0<<<grid_for(A.n_items / 2 + 1), kThreads, 0, h->stream>>>(psi, A);
      } else if (A.fix.pos[0] >= 1 && A.n_items >= 2) {
        k_pair_v2<<<grid_for(A.n_items / 4 + 1), kThreads, 0, h->stream>>>(psi, A);
      } else {
        k_pair_v1<<<grid_for(A.n_items), kThreads, 0, h->stream>>>(psi, A);
      }
      CUDA_TRY(cudaGetLastError());
      count_pass(h, 2 * A.n_items);
      return QCSIM_OK;
    }
    case OP_DENSE2: {
      DenseArgs<2> A;
      A.fix = sorted_bits(op.ctrl, op.n_ctrl, op.tgt, 2);
      A.or_ctrl = or_ctrl;
      for (int j = 0; j < 4; ++j) A.off[j] = ((j & 1) ? (1ULL << op.tgt[0]) : 0) | ((j & 2) ? (1ULL << op.tgt[1]) : 0);
      for (int j = 0; j < 16; ++j) A.m[j] = to_amp(op.m[j]);
      A.n_items = 1ULL << (nl - A.fix.n);
      if (A.fix.pos[0] >= 1 && A.n_items >= 2)
        k_dense_v2<2><<<grid_for(A.n_items / 2), kThreads, 0, h->stream>>>(psi, A);
      else
        k_dense_v1<2><<<grid_for(A.n_items), kThreads, 0, h->stream>>>(psi, A);
      CUDA_TRY(cudaGetLastError());
      count_pass(h, 4 * A.n_items);
      return QCSIM_OK;
    }
    case OP_DENSE3: {
      DenseArgs<3> A;
      A.fix = sorted_bits(op.ctrl, op.n_ctrl, op.tgt, 3);
      A.or_ctrl = or_ctrl;
      for (int j = 0; j < 8; ++j)
        A.off[j] = ((j & 1) ? (1ULL << op.tgt[0]) : 0) | ((j & 2) ? (1ULL << op.tgt[1]) : 0) | ((j & 4) ? (1ULL << op.tgt[2]) : 0);
      for (int j = 0; j < 64; ++j) A.m[j] = to_amp(op.m[j]);
      A.n_items = 1ULL << (nl - A.fix.n);
      if (A.fix.pos[0] >= 1 && A.n_items >= 2)
        k_dense_v2<3><<<grid_for(A.n_items / 2), kThreads, 0, h->stream>>>(psi, A);
      else
        k_dense_v1<3><<<grid_for(A.n_items), kThreads, 0, h->stream>>>(psi, A);
      CUDA_TRY(cudaGetLastError());
      count_pass(h, 8 * A.n_items);
      return QCSIM_OK;
    }
    case OP_DIAG: {
      DiagArgs A;
      A.ctrl = sorted_bits(op.ctrl, op.n_ctrl, nullptr, 0);
      A.or_ctrl = or_ctrl;
      A.nsel = op.n_tgt;
      for (int k = 0; k < 3; ++k) A.selpos[k] = k < op.n_tgt ? op.tgt[k] : 0;
      for (int j = 0; j < 8; ++j) A.table[j] = j < (1 << op.n_tgt) ? to_amp(op.m[j]) : make_amp(1, 0);
      A.n_items = 1ULL << (nl - A.ctrl.n);
      if ((A.ctrl.n == 0 || A.ctrl.pos[0] >= 1) && A.n_items >= 2)
        k_diag_v2<<<grid_for(A.n_items / 4 + 1), kThreads, 0, h->stream>>>(psi, A);
      else
        k_diag_v1<<<grid_for(A.n_items), kThreads, 0, h->stream>>>(psi, A);
      CUDA_TRY(cudaGetLastError());
      count_pass(h, A.n_items);
      return QCSIM_OK;
    }
  }
  return fail(QCSIM_ERR_BAD_ARG, "unknown op kind");
}

// ---- dense operator (ApplyOperatorMatrix, QubitRegister.h:499-505): psi = M psi, small registers only ----------
// one warp per row; the matrix (16 * 4^n bytes) is read once from a temporary device copy
static __global__ void __launch_bounds__(kThreads) k_dense_operator(const amp* __restrict__ M, const amp* __restrict__ in, amp* __restrict__ out,
                                                                    uint64_t dim) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < dim; r += n_warps) {
    amp acc = make_amp(0, 0);
    const amp* row = M + r * dim;
    for (uint64_t c = lane; c < dim; c += 32) acc = cadd(acc, cmul(row[c], in[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc.x += __shfl_down_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_down_sync(0xffffffffu, acc.y, o);
    }
    if (lane == 0) out[r] = acc;
  }
}

int engine_apply_operator(qcsim_sv* h, const double* m) {
  const uint64_t dim = h->dim_local;
  amp *dM = nullptr, *dout = nullptr;
  cudaError_t ce = cudaMalloc(&dM, dim * dim * sizeof(amp));
  if (ce == cudaSuccess) ce = cudaMalloc(&dout, dim * sizeof(amp));
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(dM, m, dim * dim * sizeof(amp), cudaMemcpyHostToDevice, h->stream);
  if (ce == cudaSuccess) {
    const uint64_t blocks = std::min<uint64_t>((dim * 32 + kThreads - 1) / kThreads, (uint64_t)kNumSMs * 8);
    k_dense_operator<<<(unsigned)std::max<uint64_t>(blocks, 1), kThreads, 0, h->stream>>>(dM, h->psi, dout, dim);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h->psi, dout, dim * sizeof(amp), cudaMemcpyDeviceToDevice, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  cudaFree(dM);
  cudaFree(dout);
  if (ce != cudaSuccess)
    return fail(ce == cudaErrorMemoryAllocation ? QCSIM_ERR_OOM : QCSIM_ERR_CUDA, "ApplyOperatorMatrix: %s", cudaGetErrorString(ce));
  h->stats.kernel_launches += 1;
  return QCSIM_OK;
}

int engine_apply_now(qcsim_sv* h, const Op& op) {
  if (h->world > 1) return dist_apply(h, op);
  return engine_launch_local(h, op);
}

int engine_enqueue(qcsim_sv* h, const Op& op) {
  if (op.kind != OP_NOP) h->queue.push_back(op);
  if (h->queue.size() < 16384) return QCSIM_OK;
  // The queue is bounded.  A QFT gate stream that is still arriving is recognised as a whole
  // (planner.h: match_qft), so its prefix stays queued instead of being cut here.
  std::vector<Op> head;
  head.swap(h->queue);
  std::vector<Op> tail;
  const int rc = fusion_execute_partial(h, head, head.size() < 65536 ? &tail : nullptr);
  h->queue.insert(h->queue.begin(), tail.begin(), tail.end());
  return rc;
}

void engine_drop_queue(qcsim_sv* h) { h->queue.clear(); }

int engine_flush(qcsim_sv* h) {
  if (h->queue.empty()) return QCSIM_OK;
  std::vector<Op> q;
  q.swap(h->queue);
  return fusion_execute(h, q);
}

// Flush for an observable that is diagonal and acts on the qubits of `qmask` only (planner.h): gates it cannot see
// stay queued, so a caller that reads one probability per layer still gets whole-circuit fusion.
int engine_flush_for_diagonal_observable(qcsim_sv* h, uint64_t qmask) {
  static const int lazy = [] {
    const char* e = std::getenv("QCSIM_LAZY_OBSERVABLES");
    return e ? std::atoi(e) : 1;
  }();
  if (h->queue.empty()) return QCSIM_OK;
  if (!lazy || fusion_holds_qft(h->queue)) return engine_flush(h);  // a transform is recognised from its whole gate stream
  std::vector<Op> needed, rest;
  split_queue_for_diagonal_observable(h->queue, qmask, &needed, &rest);
  if (rest.size() < 4) return engine_flush(h);  // nothing worth keeping
  h->queue.swap(rest);
  if (needed.empty()) return QCSIM_OK;
  return fusion_execute(h, needed);
}

int engine_canonicalize(qcsim_sv* h) {
  if (h->world > 1) return dist_canonicalize(h);
  return QCSIM_OK;
}

// ---- QFT (QuantumFourierTransform.h:35-87, QubitsSwapper.h:23-34) ---------------------------------

static void hadamard_matrix(double* m) {  // SimpleGates.h:588-596
  const double v = 1. / std::sqrt(2.);
  const double hm[8] = {v, 0, v, 0, v, 0, -v, 0};
  std::memcpy(m, hm, sizeof hm);
}
static void cphase_matrix(double* m, double theta) {  // QuantumGate.h:251-269, m33 = std::polar(1., theta)
  std::memset(m, 0, 32 * sizeof(double));
  m[0] = m[10] = m[20] = 1.0;
  m[30] = std::cos(theta);
  m[31] = std::sin(theta);
}
static void swap_matrix(double* m) {  // QuantumGate.h:10-28
  std::memset(m, 0, 32 * sizeof(double));
  m[0] = 1.0;
  m[2 * (1 * 4 + 2)] = 1.0;
  m[2 * (2 * 4 + 1)] = 1.0;
  m[30] = 1.0;
}

// item-index bits of a round whose register bits are `reg_mask` (tile-local): the other tile bits in
// an order whose low three have distinct (position mod 3), see swz()
static void item_bit_order(int k, uint32_t reg_mask, uint32_t* words, int n_words) {
  int rest[16], nrest = 0, order[16], n = 0;
  for (int lb = 0; lb < k; ++lb)
    if (!((reg_mask >> lb) & 1u)) rest[nrest++] = lb;
  for (int res = 0; res < 3; ++res)
    for (int i = 0; i < nrest; ++i)
      if (rest[i] >= 0 && rest[i] % 3 == res) {
        order[n++] = rest[i];
        rest[i] = -1;
        break;
      }
  for (int i = 0; i < nrest; ++i)
    if (rest[i] >= 0) order[n++] = rest[i];
  for (int w = 0; w < n_words; ++w) words[w] = 0;
  for (int j = 0; j < n && j < 4 * n_words; ++j) words[j >> 2] |= (uint32_t)order[j] << (8 * (j & 3));
}

// The QFT / IQFT gates whose TARGET is one of the logical qubits [lo, hi], as radix-8 FFT passes
// (qft_kernels.cuh).  [lo, hi] is a slice of a transform that starts at logical qubit r_floor (the twiddles read
// every lower qubit down to r_floor, wherever it lives: local bits, or rank bits on a sharded register).
// phys_of maps logical qubits to physical index bits (nullptr: identity); the qubits [lo, hi] must sit on local
// positions, in any order.
static int engine_qft_passes_plain(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of);

int engine_qft_passes(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of) {
  int ident[64];
  if (!phys_of) {
    for (int q = 0; q < 64; ++q) ident[q] = q;
    phys_of = ident;
  }
  static const bool plain_only = std::getenv("QCSIM_QFT_PLAIN") != nullptr;
  if (plain_only || !fusion_pipe_available() || h->n_local < 11) return engine_qft_passes_plain(h, lo, hi, inverse, r_floor, phys_of);
  // TMA-staged passes (qft_pipe.cu): contiguous slices of the targets, as many qubits per pass as fit an 11-bit tile
  // next to qubits 0..2; top-down for the QFT, bottom-up for the IQFT
  int cur = inverse ? lo : hi;
  while (inverse ? cur <= hi : cur >= lo) {
    int phys[16], left = inverse ? hi - cur + 1 : cur - lo + 1;
    const int look = std::min(left, 11);
    for (int j = 0; j < look; ++j) phys[j] = phys_of[inverse ? cur + j : cur - j];
    for (int j = 0; j < look; ++j)
      if (phys[j] < 0 || phys[j] >= h->n_local) return fail(QCSIM_ERR_BAD_ARG, "internal: QFT target qubit is not on a local position");
    const int n = qft_pipe_pass_capacity(phys, look);
    const int a = inverse ? cur : cur - n + 1, b = inverse ? cur + n - 1 : cur;
    const int rc = qft_pipe_pass(h, a, b, inverse, r_floor, phys_of);
    if (rc == QCSIM_ERR_UNSUPPORTED) QCSIM_TRY(engine_qft_passes_plain(h, a, b, inverse, r_floor, phys_of));
    else QCSIM_TRY(rc);
    cur = inverse ? b + 1 : a - 1;
  }
  return QCSIM_OK;
}

// the same with plain tile passes (k_qft_pass: 12-bit tiles staged by the CTA itself): small registers, fallback
static int engine_qft_passes_plain(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of) {
  const int nl = h->n_local;
  struct Grp { int top, size; };
  std::vector<Grp> groups;  // top-down
  for (int top = hi; top >= lo;) {
    const int size = std::min(3, top - lo + 1);
    groups.push_back({top, size});
    top -= size;
  }
  const int Kmax = std::min(kMaxTileBits, nl);
  const int Lmin = std::min(3, nl);
  struct Pass { std::vector<int> tile; std::vector<Grp> groups; };
  std::vector<Pass> passes;
  for (size_t i = 0; i < groups.size();) {
    uint64_t bits = (1ULL << Lmin) - 1ULL;
    Pass p;
    while (i < groups.size() && (int)p.groups.size() < kMaxQftGroups) {
      uint64_t gb = 0;
      for (int j = 0; j < groups[i].size; ++j) {
        const int ph = phys_of[groups[i].top - j];
        if (ph < 0 || ph >= nl) return fail(QCSIM_ERR_BAD_ARG, "internal: QFT target qubit is not on a local position");
        gb |= 1ULL << ph;
      }
      if (__builtin_popcountll(bits | gb) > Kmax) break;
      bits |= gb;
      p.groups.push_back(groups[i]);
      ++i;
    }
    if (p.groups.empty()) return fail(QCSIM_ERR_BAD_ARG, "internal: QFT group does not fit a tile");
    for (int q = 0; q < nl && __builtin_popcountll(bits) < Kmax; ++q) bits |= 1ULL << q;  // pad: longer contiguous runs
    for (int q = 0; q < nl; ++q)
      if ((bits >> q) & 1ULL) p.tile.push_back(q);
    passes.push_back(p);
  }
  if (inverse) {
    std::reverse(passes.begin(), passes.end());
    for (Pass& p : passes) std::reverse(p.groups.begin(), p.groups.end());
  }
  const double pi = 3.14159265358979323846;
  for (const Pass& p : passes) {
    QftPassArgs A;
    std::memset(&A, 0, sizeof A);
    const int k = (int)p.tile.size();
    A.k = k;
    int L = 0;
    while (L < k && p.tile[L] == L) ++L;
    A.low_identity = L;
    A.n_groups = (int)p.groups.size();
    A.inverse = inverse ? 1 : 0;
    A.n_tiles = 1ULL << (nl - k);
    A.rank_bits = (uint64_t)h->rank << nl;
    for (int j = 0; j < kMaxTileBits; ++j) A.tpos[j] = j < k ? p.tile[j] : 0;
    A.sq = r_floor;  // logical start qubit of the whole transform
    A.n_phys = h->n;
    for (int q = 0; q < h->n; ++q) A.log_of[phys_of[q]] = (signed char)q;
    A.s = 1. / std::sqrt(2.);
    const double sign = inverse ? -1.0 : 1.0;
    A.ph2 = make_amp(std::cos(sign * pi / 2), std::sin(sign * pi / 2));  // std::polar(1., theta)
    A.ph4 = make_amp(std::cos(sign * pi / 4), std::sin(sign * pi / 4));
    int local_of[64];
    for (int q = 0; q < 64; ++q) local_of[q] = -1;
    for (int j = 0; j < k; ++j) local_of[p.tile[j]] = j;
    for (size_t g = 0; g < p.groups.size(); ++g) {
      QftGroup& G = A.groups[g];
      G.size = p.groups[g].size;
      G.top_qubit = p.groups[g].top;
      const int low_q = p.groups[g].top - G.size + 1;
      uint32_t reg_mask = 0;
      G.rb = 0;
      for (int j = 0; j < G.size; ++j) {
        const int lb = local_of[phys_of[low_q + j]];
        G.rb |= (uint32_t)lb << (8 * j);
        reg_mask |= 1u << lb;
      }
      item_bit_order(k, reg_mask, G.tb, 3);
    }
    const size_t smem = (sizeof(amp) << k) + sizeof(amp) * kMaxQftGroups;
    const uint64_t grid = std::min<uint64_t>(A.n_tiles, (uint64_t)kNumSMs * 3);
    // twiddle table (16 B per item) followed by the slot table (4 B per item)
    if (!h->d_qft_table) CUDA_TRY(cudaMalloc(&h->d_qft_table, (sizeof(amp) + sizeof(uint32_t)) * kMaxQftGroups * kQftItemsMax));
    uint32_t* const d_slots = reinterpret_cast<uint32_t*>(h->d_qft_table + kMaxQftGroups * kQftItemsMax);
    if (k == 12) {
      k_qft_item_table<<<kMaxQftGroups * kQftItemsMax / 256, 256, 0, h->stream>>>(h->d_qft_table, d_slots, A);
      h->stats.kernel_launches += 1;
    }
    k_qft_pass<<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, h->d_qft_table, d_slots, A);
    CUDA_TRY(cudaGetLastError());
    count_pass(h, h->dim_local);
    h->stats.fused_rounds += p.groups.size();
  }
  return QCSIM_OK;
}

// In-place permutation of LOCAL physical index bits: afterwards position p holds what position
// src_of[p] held (src_of is a permutation of 0..n_local-1, identity where nothing moves).  The
// permutation is written as a sequence of position swaps; consecutive swaps whose positions fit
// one tile (with the low bits that keep HBM accesses contiguous) run as ONE in-tile permutation pass.
int engine_permute_bits(qcsim_sv* h, const int* src_of) {
  const int nl = h->n_local;
  std::vector<std::pair<int, int>> swaps;
  {
    int cur[64];  // cur[p] = original position whose content is now at p
    for (int p = 0; p < nl; ++p) cur[p] = p;
    for (int p = 0; p < nl; ++p) {
      if (cur[p] == src_of[p]) continue;
      int q = -1;
      for (int r = p + 1; r < nl; ++r)
        if (cur[r] == src_of[p]) q = r;
      if (q < 0) return fail(QCSIM_ERR_BAD_ARG, "internal: not a permutation");
      swaps.push_back({p, q});
      std::swap(cur[p], cur[q]);
    }
  }
  if (swaps.empty()) return QCSIM_OK;
  const int Kmax = std::min(kMaxTileBits, nl);
  const uint64_t low = (1ULL << std::min(2, nl)) - 1ULL;
  size_t i = 0;
  while (i < swaps.size()) {
    uint64_t bits = low;
    int perm[64];  // perm[p] = position (at the start of this pass) whose content ends up at p
    for (int p = 0; p < 64; ++p) perm[p] = p;
    while (i < swaps.size()) {
      const uint64_t nb = bits | (1ULL << swaps[i].first) | (1ULL << swaps[i].second);
      if (__builtin_popcountll(nb) > Kmax) break;
      bits = nb;
      std::swap(perm[swaps[i].first], perm[swaps[i].second]);
      ++i;
    }
    for (int q = 0; q < nl && __builtin_popcountll(bits) < Kmax; ++q) bits |= 1ULL << q;  // pad: longer contiguous runs
    PermPassArgs A;
    std::memset(&A, 0, sizeof A);
    int local_of[64], k = 0;
    for (int q = 0; q < 64; ++q) local_of[q] = -1;
    for (int q = 0; q < nl; ++q)
      if ((bits >> q) & 1ULL) {
        A.tpos[k] = q;
        local_of[q] = k++;
      }
    A.k = k;
    int L = 0;
    while (L < k && A.tpos[L] == L) ++L;
    A.low_identity = L;
    A.n_tiles = 1ULL << (nl - k);
    for (int j = 0; j < kMaxTileBits; ++j) A.src_bit[j] = j < k ? local_of[perm[A.tpos[j]]] : j;
    const size_t smem = sizeof(amp) << k;
    const uint64_t grid = std::min<uint64_t>(A.n_tiles, (uint64_t)kNumSMs * 3);
    k_tile_permute<<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
    CUDA_TRY(cudaGetLastError());
    count_pass(h, h->dim_local);
  }
  return QCSIM_OK;
}

static uint64_t qft_gate_count(int m, bool do_swap) {
  return (uint64_t)m + (uint64_t)m * (m - 1) / 2 + (do_swap ? (uint64_t)(m / 2) : 0);
}

// generic gate-by-gate expansion (QuantumFourierTransform.h:35-87, QubitsSwapper.h:23-34); used by
// sharded registers in non-canonical layouts and by QCSIM_QFT_GENERIC=1
static std::vector<Op> qft_gate_ops(int sq, int eq, bool do_swap, bool inverse) {
  double hm[8], cp[32], sw[32];
  hadamard_matrix(hm);
  swap_matrix(sw);
  std::vector<Op> ops;
  auto H = [&](int q) { ops.push_back(classify(1, hm, 0, q, 0, 0)); };
  auto CP = [&](int q, int c, double phase) {
    cphase_matrix(cp, phase);
    ops.push_back(classify(2, cp, QCSIM_GATE_CONTROLLED | QCSIM_GATE_DIAGONAL, q, c, 0));
  };
  auto swaps = [&]() {
    int s = sq, e = eq;
    while (s < e) {
      ops.push_back(classify(2, sw, QCSIM_GATE_SWAP, s, e, 0));
      ++s;
      --e;
    }
  };
  const double pi_2 = 1.57079632679489661923;  // M_PI_2
  if (!inverse) {
    H(eq);
    for (int cur = eq; cur > sq; --cur) {
      double phase = pi_2;
      for (int ctrl = cur - 1; ctrl >= sq; --ctrl) {
        CP(cur, ctrl, phase);
        phase *= 0.5;
      }
      H(cur - 1);
    }
    if (do_swap) swaps();
  } else {
    if (do_swap) swaps();
    for (int cur = sq + 1; cur <= eq; ++cur) {
      H(cur - 1);
      double phase = -pi_2;
      for (int ctrl = cur - 1; ctrl >= sq; --ctrl) {
        CP(cur, ctrl, phase);
        phase *= 0.5;
      }
    }
    H(eq);
  }
  return ops;
}

static int engine_qft_generic(qcsim_sv* h, int sq, int eq, bool do_swap, bool inverse) {
  const std::vector<Op> ops = qft_gate_ops(sq, eq, do_swap, inverse);
  if (h->fusion) {
    for (const Op& op : ops) QCSIM_TRY(engine_enqueue(h, op));
    return QCSIM_OK;
  }
  QCSIM_TRY(engine_flush(h));
  return fusion_execute(h, ops);
}

int engine_reverse_bits(qcsim_sv* h, int sq, int eq) {
  const int m = eq - sq + 1;
  static const bool no_cobra = std::getenv("QCSIM_NO_COBRA") != nullptr;
  if (sq == 0 && m >= 2 * kRevT && !no_cobra) {  // whole-register style reversal: one pass (k_bit_reverse)
    const size_t smem = 2 * sizeof(amp) << (2 * kRevT);
    BitRevArgs A;
    A.m = m;
    A.n_local = h->n_local;
    A.n_work = 1ULL << (h->n_local - 2 * kRevT);
    const uint64_t grid = std::min<uint64_t>(A.n_work, (uint64_t)kNumSMs * 4);
    k_bit_reverse<<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
    CUDA_TRY(cudaGetLastError());
    count_pass(h, h->dim_local);
    return QCSIM_OK;
  }
  int src_of[64];
  for (int p = 0; p < 64; ++p) src_of[p] = p;
  for (int s = sq, e = eq; s < e; ++s, --e) {
    src_of[s] = e;
    src_of[e] = s;
  }
  return engine_permute_bits(h, src_of);
}

int engine_qft(qcsim_sv* h, uint64_t sq_, uint64_t eq_, bool do_swap, bool inverse) {
  // sub-register clamp as in QuantumSubAlgorithmOnSubregister (QuantumAlgorithm.h): eq = max(sq, min(N-1, eq))
  const uint64_t nm1 = (uint64_t)h->n - 1;
  if (sq_ > nm1) return fail(QCSIM_ERR_QUBIT_TOO_HIGH, "Qubit number is too high");
  const int sq = (int)sq_;
  const int eq = (int)std::max<uint64_t>(sq_, std::min<uint64_t>(nm1, eq_));
  h->stats.gates_applied += qft_gate_count(eq - sq + 1, do_swap);
  static const bool generic = std::getenv("QCSIM_QFT_GENERIC") != nullptr;
  if (generic) return engine_qft_generic(h, sq, eq, do_swap, inverse);
  QCSIM_TRY(engine_flush(h));  // queued gates come first
  return engine_qft_direct(h, sq, eq, do_swap, inverse);
}

// the transform itself, on qubits [sq, eq] (already validated, nothing queued before it)
int engine_qft_direct(qcsim_sv* h, int sq, int eq, bool do_swap, bool inverse) {
  NvtxRange nvtx_range("qcsim.qft");
  if (h->world > 1) {
    int handled = 0;
    QCSIM_TRY(dist_qft(h, sq, eq, do_swap, inverse, &handled));
    if (handled) return QCSIM_OK;
    // transform entirely on global qubits: expand and run through the sharded gate path
    return dist_execute(h, qft_gate_ops(sq, eq, do_swap, inverse));
  }
  if (inverse && do_swap) QCSIM_TRY(engine_reverse_bits(h, sq, eq));
  QCSIM_TRY(engine_qft_passes(h, sq, eq, inverse, sq, nullptr));
  if (!inverse && do_swap) QCSIM_TRY(engine_reverse_bits(h, sq, eq));
  return QCSIM_OK;
}

// ---- measurement scan ----------------------------------------------------------------------------

static int ensure_scan_buffers(qcsim_sv* h, uint64_t count) {
  if (!h->d_prefix_hi) {
    CUDA_TRY(cudaMalloc(&h->d_prefix_hi, h->n_chunks * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->d_chunk_K, h->n_chunks * sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&h->d_chunk_flags, h->n_chunks * sizeof(int)));
    CUDA_TRY(cudaMalloc(&h->d_acc_start, (h->n_chunks + 1) * sizeof(double)));
  }
  if (h->draws_capacity < count) {
    cudaFree(h->d_draws);
    cudaFree(h->d_outcomes);
    h->d_draws = nullptr;
    h->d_outcomes = nullptr;
    h->draws_capacity = 0;
    const uint64_t cap = std::max<uint64_t>(count, 1024);
    CUDA_TRY(cudaMalloc(&h->d_draws, cap * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->d_outcomes, cap * sizeof(unsigned long long)));
    h->draws_capacity = cap;
  }
  return QCSIM_OK;
}

// The reference's running sum at every chunk start of this slice (reduce_kernels.cuh), then every draw resolved
// against it.  On a sharded register the slices are chained: rank r starts from the sum rank r-1 ended with.
int engine_resolve_draws(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes) {
  NvtxRange nvtx_range("qcsim.measure_scan");
  if (count == 0) return QCSIM_OK;
  QCSIM_TRY(engine_canonicalize(h));
  QCSIM_TRY(ensure_scan_buffers(h, count));
  const int g = (int)std::min<uint64_t>(h->n_chunks, (uint64_t)kMaxPartials);
  k_chunk_sums<<<g, kThreads, 0, h->stream>>>(h->psi, h->dim_local, h->n_chunks, h->d_chunk_sums);
  dd offset = dd_make(0, 0);
  double start = 0.0;
  if (h->world > 1) QCSIM_TRY(dist_scan_offset(h, &offset));  // exact mass of the lower ranks (binade prediction)
  k_chunk_prefix<<<1, 1024, 0, h->stream>>>(h->d_chunk_sums, h->n_chunks, offset, h->d_prefix_hi, h->d_total);
  k_chunk_increments<<<g, kThreads, 0, h->stream>>>(h->psi, h->dim_local, h->n_chunks, h->d_prefix_hi, h->d_chunk_K, h->d_chunk_flags);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 3;
  h->stats.state_passes += 2;
  h->stats.bytes_moved += 32ULL * h->dim_local;
  if (h->world > 1) {
    QCSIM_TRY(dist_chained_walk(h));  // rank by rank: walk with the predecessor's final sum
  } else {
    k_sequential_walk<<<1, kWalkThreads, 0, h->stream>>>(h->psi, h->dim_local, h->n_chunks, h->d_prefix_hi, h->d_chunk_K, h->d_chunk_flags, start,
                                                     h->d_acc_start, nullptr);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_draws, probs, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  k_resolve_draws<<<(unsigned)((count + 127) / 128), 128, 0, h->stream>>>(h->psi, h->dim_local, h->n_chunks, h->d_acc_start, h->d_draws, count,
                                                                          h->d_outcomes);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  if (h->world > 1) QCSIM_TRY(dist_combine_outcomes(h, h->d_outcomes, count));  // local -> global index, one owner per draw
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "outcome type");
  CUDA_TRY(cudaMemcpyAsync(outcomes, h->d_outcomes, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
  QCSIM_TRY(engine_wait(h));
  return QCSIM_OK;
}

int engine_pick_state(qcsim_sv* h, double prob, uint64_t fallback, uint64_t* outcome) {
  uint64_t s = 0;
  QCSIM_TRY(engine_resolve_draws(h, &prob, 1, &s));
  *outcome = (s == ~0ULL) ? fallback : s;
  return QCSIM_OK;
}

// RepeatedMeasure (QubitRegister.h:227-273): `count` draws against ONE cumulative table.  The reference cuts its
// table at the first index whose running sum exceeds 1 - DBL_EPSILON (:250-254) and std::lower_bound returns end()
// for a draw above the last stored value (:268), i.e. the outcome "table size"; both are reproduced: the cut is one
// more draw resolved against the same running sum.
int engine_sample(qcsim_sv* h, const double* probs, uint64_t count, uint64_t* outcomes) {
  if (count == 0) return QCSIM_OK;
  std::vector<double> p(probs, probs + count);
  const double cut_threshold = std::nextafter(1.0 - 2.220446049250313e-16, 2.0);  // first acc > 1 - eps  <=>  first acc >= the next double
  p.push_back(cut_threshold);
  std::vector<uint64_t> o(count + 1);
  QCSIM_TRY(engine_resolve_draws(h, p.data(), count + 1, o.data()));
  const uint64_t cut = o[count];
  const uint64_t table_size = (cut == ~0ULL) ? h->dim : cut + 1;
  for (uint64_t i = 0; i < count; ++i) outcomes[i] = (o[i] == ~0ULL || o[i] >= table_size) ? table_size : o[i];
  return QCSIM_OK;
}

}  // namespace qcsim
