// fusion.cu -- see fusion.h
#include "fusion.h"

namespace qcsim {

int fusion_execute(qcsim_sv* h, const std::vector<Op>& ops) {
  for (const Op& op : ops) QCSIM_TRY(engine_apply_now(h, op));
  return QCSIM_OK;
}

}  // namespace qcsim
