// dist.h -- sharded registers: the state is split over `world` GPUs on its top log2(world)
// qubits; one process per GPU.  See dist.cu.
#pragma once

#include "engine.h"

namespace qcsim {

int dist_init(qcsim_sv* h, const void* nccl_id);
void dist_shutdown(qcsim_sv* h);
int dist_unique_id(void* out128);
void dist_reset_layout(qcsim_sv* h);
int dist_buffers_changed(qcsim_sv* h);
void dist_map_mask(qcsim_sv* h, uint64_t mask, uint64_t want, uint64_t* pmask, uint64_t* pwant);
int dist_wait(qcsim_sv* h);  // bounded wait for the stream (a collective whose peers never arrive must not hang the host)
int dist_allreduce_host(qcsim_sv* h, double* vals, int count);
int dist_apply(qcsim_sv* h, const Op& op);
int dist_canonicalize(qcsim_sv* h);
int dist_scan_offset(qcsim_sv* h, dd* offset);
int dist_chained_walk(qcsim_sv* h);
int dist_combine_outcomes(qcsim_sv* h, unsigned long long* d_outcomes, uint64_t count);
// fast QFT on a sharded register; *handled = 0 when the caller must fall back to the gate-by-gate path
int dist_qft(qcsim_sv* h, int sq, int eq, bool do_swap, bool inverse, int* handled);
void dist_collect_stats(qcsim_sv* h);  // resolves the CUDA-event timings of finished exchanges into stats.exchange_ms

}  // namespace qcsim
