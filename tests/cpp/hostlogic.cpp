// CPU-side test shim: exposes the engine's pure host logic (classify.h, planner.h) to ctypes so
// tests/test_planner.py can check it without a GPU.  Built with g++ by the test.
#include <cstring>
#include <vector>

#include "../../include/qcsim_b200.h"
#include "../../qcsim_b200/csrc/classify.h"
#include "../../qcsim_b200/csrc/planner.h"

using namespace qcsim;

extern "C" {

struct hl_op {
  int kind, n_ctrl, ctrl[3], n_tgt, tgt[3];
  double m[128];
};

static void export_op(const Op& op, hl_op* o) {
  o->kind = op.kind;
  o->n_ctrl = op.n_ctrl;
  o->n_tgt = op.n_tgt;
  for (int i = 0; i < 3; ++i) {
    o->ctrl[i] = op.ctrl[i];
    o->tgt[i] = op.tgt[i];
  }
  for (int i = 0; i < 64; ++i) {
    o->m[2 * i] = op.m[i].real();
    o->m[2 * i + 1] = op.m[i].imag();
  }
}

void hl_classify(const qcsim_gate* g, hl_op* out) { export_op(classify(g->nq, g->m, g->flags, g->q, g->c1, g->c2), out); }

// Plans `count` gates; writes per step: fused flag, tile mask, number of ops, then op indices into
// `order`.  Returns the number of steps.
int hl_plan(const qcsim_gate* gates, int count, int n_local, int K, int L, int* step_fused, unsigned long long* step_tile,
            int* step_nops, int* order, hl_op* ops_out) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) {
    ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    if (ops_out) export_op(ops.back(), &ops_out[i]);
  }
  const std::vector<PlanStep> steps = plan_passes(ops, n_local, K, L, 64, 512);
  int o = 0;
  for (size_t s = 0; s < steps.size(); ++s) {
    step_fused[s] = steps[s].fused ? 1 : 0;
    unsigned long long t = 0;
    for (int q : steps[s].pass.tile) t |= 1ULL << q;
    step_tile[s] = t;
    step_nops[s] = (int)steps[s].pass.ops.size();
    for (int i : steps[s].pass.ops) order[o++] = i;
  }
  return (int)steps.size();
}
}
