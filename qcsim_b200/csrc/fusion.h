// fusion.h -- gate-stream planner + fused shared-memory gate blocks (several gates per HBM pass).
#pragma once

#include <vector>

#include "engine.h"

namespace qcsim {

// Execute `ops` in order on the register.  Consecutive ops whose non-diagonal targets fit in one
// shared-memory tile are applied in a single pass over HBM; everything else runs as single-gate
// kernels.  Result is identical (to rounding) to applying the ops one by one.
int fusion_execute(qcsim_sv* h, const std::vector<Op>& ops);
int fusion_execute_partial(qcsim_sv* h, const std::vector<Op>& ops, std::vector<Op>* deferred);

// same, for ops whose qubit indices are already physical bit positions of the local slice
int fusion_execute_local(qcsim_sv* h, const std::vector<Op>& ops);
int dist_execute(qcsim_sv* h, const std::vector<Op>& ops);

int engine_launch_local(qcsim_sv* h, const Op& op);

}  // namespace qcsim
