// dist.cu -- sharded registers (placeholder until the NCCL / peer-memory exchange lands).
#include "dist.h"
#include "fusion.h"

namespace qcsim {

int dist_init(qcsim_sv*, const void*) { return fail(QCSIM_ERR_UNSUPPORTED, "sharded registers are not built yet"); }
void dist_shutdown(qcsim_sv*) {}
int dist_unique_id(void*) { return fail(QCSIM_ERR_UNSUPPORTED, "sharded registers are not built yet"); }
void dist_reset_layout(qcsim_sv*) {}
int dist_buffers_changed(qcsim_sv*) { return QCSIM_OK; }
void dist_map_mask(qcsim_sv*, uint64_t mask, uint64_t want, uint64_t* pmask, uint64_t* pwant) {
  *pmask = mask;
  *pwant = want;
}
int dist_allreduce_host(qcsim_sv*, double*, int) { return QCSIM_OK; }
int dist_apply(qcsim_sv*, const Op&) { return fail(QCSIM_ERR_UNSUPPORTED, "sharded registers are not built yet"); }
int dist_canonicalize(qcsim_sv*) { return QCSIM_OK; }
int dist_pick_state(qcsim_sv*, double, uint64_t, uint64_t*) { return fail(QCSIM_ERR_UNSUPPORTED, "sharded registers are not built yet"); }

int dist_execute(qcsim_sv*, const std::vector<Op>&) { return fail(QCSIM_ERR_UNSUPPORTED, "sharded registers are not built yet"); }

int engine_nccl_unique_id(void* out) { return dist_unique_id(out); }

}  // namespace qcsim
