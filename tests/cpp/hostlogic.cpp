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

// ---- sharded planning (dist_plan.h) -----------------------------------------------------------
#include "../../qcsim_b200/csrc/dist_plan.h"

extern "C" {

// Plans `count` gates for `rank` from the layout phys_of[0..n) (updated in place).  Steps are
// written flat: step_kind (0 local, 1 exchange), step_nops / ops_out for local steps, ex_k /
// ex_g / ex_l (3 ints per step) for exchanges.  canonicalize != 0: plan the return to the identity
// layout instead (gates ignored).  Returns the number of steps, or -1 if a buffer is too small.
int hl_dist_plan(const qcsim_gate* gates, int count, int n, int n_local, int rank, int* phys_of, int canonicalize,
                 int max_steps, int max_ops, int* step_kind, int* step_nops, hl_op* ops_out, int* ex_k, int* ex_g, int* ex_l) {
  DistLayout L;
  L.reset(n, n_local);
  for (int q = 0; q < n; ++q) {
    L.phys_of[q] = phys_of[q];
    L.log_of[phys_of[q]] = q;
  }
  std::vector<DistStep> steps;
  if (canonicalize) {
    steps = dist_plan_canonicalize(L);
  } else {
    std::vector<Op> ops;
    for (int i = 0; i < count; ++i) ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    steps = dist_plan(L, ops, rank);
  }
  if ((int)steps.size() > max_steps) return -1;
  int o = 0;
  for (size_t s = 0; s < steps.size(); ++s) {
    step_kind[s] = steps[s].exchange ? 1 : 0;
    step_nops[s] = (int)steps[s].ops.size();
    ex_k[s] = steps[s].k;
    for (int j = 0; j < 3; ++j) {
      ex_g[3 * s + j] = steps[s].gpos[j];
      ex_l[3 * s + j] = steps[s].lpos[j];
    }
    for (const Op& op : steps[s].ops) {
      if (o >= max_ops) return -1;
      export_op(op, &ops_out[o++]);
    }
  }
  for (int q = 0; q < n; ++q) phys_of[q] = L.phys_of[q];
  return (int)steps.size();
}
}

// ---- QFT stream recognition (planner.h: match_qft) ----------------------------------------------
extern "C" {
// out = {length, sq, eq, do_swap, inverse}
void hl_match_qft(const qcsim_gate* gates, int count, int start, int min_qubits, int* out) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
  const QftMatch m = match_qft(ops, (size_t)start, min_qubits);
  out[0] = m.length;
  out[1] = m.sq;
  out[2] = m.eq;
  out[3] = m.do_swap ? 1 : 0;
  out[4] = m.inverse ? 1 : 0;
}
}

// ---- rounds of a fused pass (planner.h: schedule_rounds) -----------------------------------------
extern "C" {
// Plans `count` gates as ONE pass over the tile qubits in `tile_mask` and schedules its rounds.
// Per round: rbits (3 ints, tile-local), n_var + vq (3 ints), n_ops; op indices appended to `order`.
// Returns the number of rounds.
int hl_rounds(const qcsim_gate* gates, int count, unsigned long long tile_mask, int max_var, int* rbits, int* nvar, int* vq,
              int* nops, int* order, int* item_bits) {
  std::vector<Op> ops;
  PassPlan plan;
  for (int q = 0; q < 64; ++q)
    if ((tile_mask >> q) & 1ULL) plan.tile.push_back(q);
  for (int i = 0; i < count; ++i) {
    ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    if (ops.back().kind != OP_NOP) plan.ops.push_back(i);
  }
  const std::vector<RoundPlan> rounds = schedule_rounds(ops, plan, max_var);
  int o = 0;
  for (size_t r = 0; r < rounds.size(); ++r) {
    for (int j = 0; j < 3; ++j) rbits[3 * r + j] = rounds[r].rbits[j];
    nvar[r] = (int)rounds[r].vq.size();
    for (int j = 0; j < 3; ++j) vq[3 * r + j] = j < nvar[r] ? rounds[r].vq[j] : -1;
    for (int j = 0; j < 9; ++j) item_bits[9 * r + j] = rounds[r].item_bit[j];
    nops[r] = (int)rounds[r].ops.size();
    for (int i : rounds[r].ops) order[o++] = i;
  }
  return (int)rounds.size();
}
}

// ---- round matrices of a fused pass (planner.h: build_round_matrices) ----------------------------
extern "C" {
// Same planning as hl_rounds; additionally writes, round after round, the 2^nvar 8x8 matrices
// (row-major, re/im pairs) into `mats` (capacity max_mats matrices).  Returns the number of rounds,
// or -1 if `mats` is too small.
int hl_round_matrices(const qcsim_gate* gates, int count, unsigned long long tile_mask, int max_var, int* rbits, int* nvar,
                      int* vq, double* mats, int max_mats) {
  std::vector<Op> ops;
  PassPlan plan;
  for (int q = 0; q < 64; ++q)
    if ((tile_mask >> q) & 1ULL) plan.tile.push_back(q);
  for (int i = 0; i < count; ++i) {
    ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    if (ops.back().kind != OP_NOP) plan.ops.push_back(i);
  }
  const std::vector<RoundPlan> rounds = schedule_rounds(ops, plan, max_var);
  int used = 0;
  for (size_t r = 0; r < rounds.size(); ++r) {
    const int nv = (int)rounds[r].vq.size();
    if (used + (1 << nv) > max_mats) return -1;
    for (int j = 0; j < 3; ++j) rbits[3 * r + j] = rounds[r].rbits[j];
    nvar[r] = nv;
    for (int j = 0; j < 3; ++j) vq[3 * r + j] = j < nv ? rounds[r].vq[j] : -1;
    build_round_matrices(ops, plan, rounds[r], reinterpret_cast<cplx*>(mats) + (size_t)used * 64);
    used += 1 << nv;
  }
  return (int)rounds.size();
}
}

// ---- TMA-staged pass (planner.h: tma_tile_geometry; fusion.cu: launch_pass_pipe) ------------------
extern "C" {
// geometry of a tile set: out = {n_dims, n_enum, box_log2, dim_lo[5], dim_bits[5], box_bits[5], enum_pos[9], slot_qubit[12]}
static void export_geometry(const TmaTileGeom& g, int* out);
int hl_tma_geometry(unsigned long long tile_mask, int n_local, int* out) {
  std::vector<int> tile;
  for (int q = 0; q < 64; ++q)
    if ((tile_mask >> q) & 1ULL) tile.push_back(q);
  TmaTileGeom g;
  if (!tma_tile_geometry(tile, n_local, &g)) return 0;
  export_geometry(g, out);
  return 1;
}
}

static void export_geometry(const TmaTileGeom& g, int* out) {
  int o = 0;
  out[o++] = g.n_dims;
  out[o++] = g.n_enum;
  out[o++] = g.box_log2;
  for (int d = 0; d < 5; ++d) out[o++] = g.dim_lo[d];
  for (int d = 0; d < 5; ++d) out[o++] = g.dim_bits[d];
  for (int d = 0; d < 5; ++d) out[o++] = g.box_bits[d];
  for (int j = 0; j < 9; ++j) out[o++] = j < g.n_enum ? g.enum_pos[j] : 0;
  for (int j = 0; j < 12; ++j) out[o++] = j < g.k ? g.slot_qubit[j] : 0;
}

extern "C" {

// Exactly what launch_pass_pipe hands to k_tile_pipe for ONE pass over `tile_mask` (11 qubits): the tensor-map
// geometry with the dimension order chosen against bank conflicts (geom_out, layout of hl_tma_geometry), the tile in
// slot order, the rounds scheduled against the TMA swizzle, their descriptors (6 uint32 each: rb, tb[3], var,
// mat_off) and matrices.  Returns the number of rounds, or -1.
int hl_pipe_pass(const qcsim_gate* gates, int count, unsigned long long tile_mask, int n_local, unsigned* desc, double* mats, int max_mats,
                 int* geom_out, int layout_search) {
  std::vector<Op> ops;
  PassPlan plan;
  for (int q = 0; q < 64; ++q)
    if ((tile_mask >> q) & 1ULL) plan.tile.push_back(q);
  for (int i = 0; i < count; ++i) {
    ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    if (ops.back().kind != OP_NOP) plan.ops.push_back(i);
  }
  TmaTileGeom base, g;
  if (!tma_tile_geometry(plan.tile, n_local, &base)) return -1;
  const PassPlan sorted_plan = plan;
  const std::vector<RoundPlan> rounds = schedule_rounds_best_layout(ops, sorted_plan, base, 3, &g, &plan, nullptr, false, layout_search != 0);
  export_geometry(g, geom_out);
  int local_of[64];
  for (int q = 0; q < 64; ++q) local_of[q] = -1;
  for (int j = 0; j < g.k; ++j) local_of[plan.tile[j]] = j;
  int used = 0;
  for (size_t r = 0; r < rounds.size(); ++r) {
    const int nv = (int)rounds[r].vq.size();
    if (used + (1 << nv) > max_mats) return -1;
    build_round_matrices(ops, plan, rounds[r], reinterpret_cast<cplx*>(mats) + (size_t)used * 64);
    const RoundDescHost rd = make_round_desc(rounds[r], local_of, (uint32_t)used);
    unsigned* d = desc + 6 * r;
    d[0] = rd.rb; d[1] = rd.tb[0]; d[2] = rd.tb[1]; d[3] = rd.tb[2]; d[4] = rd.var; d[5] = rd.mat_off;
    used += 1 << nv;
  }
  return (int)rounds.size();
}
}

// ---- what fusion_execute_local would launch (fusion.cu), counted without a GPU --------------------------------
extern "C" {
// out = {fused launches, rounds, ops run as single-gate kernels, fused passes before launch splitting, matrices}
void hl_fusion_stats(const qcsim_gate* gates, int count, int n_local, int K, int L, int max_rounds, int max_mats, int swizzle_kind, int* out) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
  const std::vector<PlanStep> steps = plan_passes(ops, n_local, K, L, 1 << 20, 1 << 30);
  int launches = 0, rounds = 0, singles = 0, passes = 0, mats = 0;
  for (const PlanStep& st : steps) {
    if (!st.fused) {
      ++singles;
      continue;
    }
    const std::vector<RoundPlan> rplan = schedule_rounds(ops, st.pass, 3, swizzle_kind);
    size_t n_mats = 0;
    for (const RoundPlan& rp : rplan) n_mats += (size_t)1 << rp.vq.size();
    // launch splitting as in launch_pass_pipe
    int l = 0;
    size_t r = 0;
    while (r < rplan.size()) {
      int nr = 0;
      size_t mi = 0;
      while (r < rplan.size() && nr < max_rounds && mi + ((size_t)1 << rplan[r].vq.size()) <= (size_t)max_mats) {
        mi += (size_t)1 << rplan[r].vq.size();
        ++nr;
        ++r;
      }
      ++l;
    }
    const double cost_fused = std::max(32.0 * l, 24.0 * rplan.size());
    double cost_alone = 0;
    for (int idx : st.pass.ops) cost_alone += standalone_cost(ops[idx]);
    if (cost_fused >= cost_alone) {
      singles += (int)st.pass.ops.size();
      continue;
    }
    ++passes;
    launches += l;
    rounds += (int)rplan.size();
    mats += (int)n_mats;
  }
  out[0] = launches;
  out[1] = rounds;
  out[2] = singles;
  out[3] = passes;
  out[4] = mats;
}
}

// ---- bank-conflict census of the DMMA fragment accesses for a whole gate list (what launch_pass_pipe would run) ----
extern "C" {
// hist[d] = number of (round, access kind) pairs whose half-warp conflict degree is d (d = 1, 2, 4, 8); returns rounds
int hl_pipe_conflicts(const qcsim_gate* gates, int count, int n_local, int* hist, int* chained) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
  const std::vector<PlanStep> steps = plan_passes(ops, n_local, 11, 3, 1 << 20, 1 << 30);
  for (int d = 0; d < 16; ++d) hist[d] = 0;
  int rounds = 0;
  *chained = 0;
  auto bank = [](uint32_t s) { return (s ^ (s >> 3)) & 7u; };
  for (const PlanStep& st : steps) {
    if (!st.fused || (int)st.pass.tile.size() != 11) continue;
    TmaTileGeom g;
    if (!tma_tile_geometry(st.pass.tile, n_local, &g)) continue;
    PassPlan plan = st.pass;
    TmaTileGeom base = g;
    const std::vector<RoundPlan> rplan = schedule_rounds_best_layout(ops, st.pass, base, 3, &g, &plan);
    for (const RoundPlan& rp : rplan) {
      ++rounds;
      if (rp.chain_next) ++*chained;
      const uint32_t r0 = 1u << rp.rbits[0], r1 = 1u << rp.rbits[1], i0 = 1u << rp.item_bit[0], i1 = 1u << rp.item_bit[1], i2 = 1u << rp.item_bit[2];
      for (int kind = 0; kind < 2; ++kind) {
        int cnt[16] = {0}, worst = 0;
        for (int lane = 0; lane < 16; ++lane) {
          const uint32_t s = kind == 0 ? (((lane & 1) ? r0 : 0u) ^ ((lane & 2) ? r1 : 0u) ^ ((lane & 4) ? i0 : 0u) ^ ((lane & 8) ? i1 : 0u))
                                       : (((lane & 1) ? i1 : 0u) ^ ((lane & 2) ? i2 : 0u) ^ ((lane & 4) ? r0 : 0u) ^ ((lane & 8) ? r1 : 0u));
          const int half = kind == 0 ? (lane & 1) : ((lane >> 2) & 1);
          worst = std::max(worst, ++cnt[2 * bank(s) + half]);
        }
        hist[worst]++;
      }
    }
  }
  return rounds;
}
// Diagnostic dump of every DMMA round of the TMA pipeline plan (slot order, register / item bits, conflict degrees).
void hl_pipe_round_dump(const qcsim_gate* gates, int count, int n_local) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
  const std::vector<PlanStep> steps = plan_passes(ops, n_local, 11, 3, 1 << 20, 1 << 30);
  auto bank = [](uint32_t s) { return (s ^ (s >> 3)) & 7u; };
  for (const PlanStep& st : steps) {
    if (!st.fused || (int)st.pass.tile.size() != 11) continue;
    TmaTileGeom g;
    if (!tma_tile_geometry(st.pass.tile, n_local, &g)) continue;
    PassPlan plan = st.pass;
    TmaTileGeom base = g;
    const std::vector<RoundPlan> rplan = schedule_rounds_best_layout(ops, st.pass, base, 3, &g, &plan);
    printf("pass slots:");
    for (int j = 0; j < 11; ++j) printf(" %d", g.slot_qubit[j]);
    printf("  (boxed 2^%d, enum %d)\n", g.box_log2, g.n_enum);
    for (const RoundPlan& rp : rplan) {
      const uint32_t r0 = 1u << rp.rbits[0], r1 = 1u << rp.rbits[1], i0 = 1u << rp.item_bit[0], i1 = 1u << rp.item_bit[1], i2 = 1u << rp.item_bit[2];
      int deg[2];
      for (int kind = 0; kind < 2; ++kind) {
        int cnt[16] = {0}, worst = 0;
        for (int lane = 0; lane < 16; ++lane) {
          const uint32_t s = kind == 0 ? (((lane & 1) ? r0 : 0u) ^ ((lane & 2) ? r1 : 0u) ^ ((lane & 4) ? i0 : 0u) ^ ((lane & 8) ? i1 : 0u))
                                       : (((lane & 1) ? i1 : 0u) ^ ((lane & 2) ? i2 : 0u) ^ ((lane & 4) ? r0 : 0u) ^ ((lane & 8) ? r1 : 0u));
          const int half = kind == 0 ? (lane & 1) : ((lane >> 2) & 1);
          worst = std::max(worst, ++cnt[2 * bank(s) + half]);
        }
        deg[kind] = worst;
      }
      printf("   round r=(%d,%d,%d) items=(%d,%d,%d | %d %d %d %d %d) nvar=%d chain=%d  load x%d store x%d\n", rp.rbits[0], rp.rbits[1], rp.rbits[2], rp.item_bit[0],
             rp.item_bit[1], rp.item_bit[2], rp.item_bit[3], rp.item_bit[4], rp.item_bit[5], rp.item_bit[6], rp.item_bit[7], (int)rp.vq.size(), (int)rp.chain_next, deg[0], deg[1]);
    }
  }
}
}

// ---- lazy flush for a diagonal single-qubit observable (planner.h: split_queue_for_diagonal_observable) -------
extern "C" {
// needed[i] = 1 when gate i has to run before the observable on the qubits of qmask is read
void hl_split_for_observable(const qcsim_gate* gates, int count, unsigned long long qmask, int* needed) {
  std::vector<Op> ops;
  for (int i = 0; i < count; ++i) {
    ops.push_back(classify(gates[i].nq, gates[i].m, gates[i].flags, gates[i].q, gates[i].c1, gates[i].c2));
    ops.back().m[63] = cplx((double)i, 0.0);  // tag (the split copies ops; entry 63 is unused below 3 qubits ... restored by index)
  }
  // the split keeps order, so a two-pointer walk recovers the indices
  std::vector<Op> need, rest;
  split_queue_for_diagonal_observable(ops, qmask, &need, &rest);
  size_t a = 0, b = 0;
  for (int i = 0; i < count; ++i) {
    const bool in_need = a < need.size() && need[a].m[63] == cplx((double)i, 0.0);
    needed[i] = in_need ? 1 : 0;
    if (in_need) ++a;
    else ++b;
  }
}
}
