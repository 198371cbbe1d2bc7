// fusion.h -- gate-stream planner + fused shared-memory gate blocks (several gates per HBM pass).
#pragma once

#include <vector>

#include "engine.h"

namespace qcsim {

// Execute `ops` in order on the register.  QFT / IQFT gate streams are recognised and run as
// radix-8 passes (qft_kernels.cuh); of the rest, ops whose non-diagonal targets fit in one
// shared-memory tile are applied in a single pass over HBM (tile_kernels.cuh) and everything else
// runs as single-gate kernels.  Result is identical (to rounding) to applying the ops one by one.
int fusion_execute(qcsim_sv* h, const std::vector<Op>& ops);
bool fusion_holds_qft(const std::vector<Op>& ops);  // the list contains a QFT / IQFT gate stream or the start of one
// same; a trailing run of ops that is still a valid QFT prefix is returned in `deferred` instead of
// being executed (used when the bounded gate queue is flushed while a transform is still arriving)
int fusion_execute_partial(qcsim_sv* h, const std::vector<Op>& ops, std::vector<Op>* deferred);

// same, for ops whose qubit indices are already physical bit positions of the local slice
int fusion_execute_local(qcsim_sv* h, const std::vector<Op>& ops);
int dist_execute(qcsim_sv* h, const std::vector<Op>& ops);

int engine_launch_local(qcsim_sv* h, const Op& op);

// TMA tile geometry of a tile set -> tensor map + coordinate recipe for the handle's state (fusion.cu)
struct TmaTileGeom;
struct PipeGeom;
int fusion_fill_pipe_geom(qcsim_sv* h, const std::vector<int>& tile_sorted, const TmaTileGeom& g, PipeGeom* G);
bool fusion_pipe_available();

// QFT passes on the TMA pipeline (qft_pipe.cu)
int qft_pipe_init_device_kernels();
int qft_pipe_pass_capacity(const int* phys, int count);
int qft_pipe_pass(qcsim_sv* h, int lo, int hi, bool inverse, int r_floor, const int* phys_of);

}  // namespace qcsim
