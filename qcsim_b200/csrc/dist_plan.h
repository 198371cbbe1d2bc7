// dist_plan.h -- planning for sharded registers.  Pure host code (no CUDA, no NCCL): unit-tested
// on the CPU with simulated shards (tests/test_dist_plan.py) before it ever meets a GPU.
//
// The state of n qubits is split over world = 2^g ranks.  PHYSICAL bit positions
// [0, n_local) address an amplitude inside a rank's slice, positions [n_local, n) are the bits of
// the rank number.  A DistLayout maps the LOGICAL qubits of the circuit onto physical positions;
// it starts as the identity (rank = top g qubits) and changes when
//   * a SWAP gate is applied: pure relabelling, no data moves;
//   * a gate needs non-diagonal access to a qubit that currently sits on a global position: the
//     planner emits an EXCHANGE step that swaps k global positions with k local positions (an
//     all-to-all among the 2^k ranks that differ in those bits).  The local partners are the
//     positions of the k local qubits whose next non-diagonal use is farthest away -- wherever they
//     sit: the exchange kernel moves strided runs of 2^(lowest partner position) amplitudes, so no
//     local "parking" pass is needed (positions below kMinExchangePos are avoided: short runs).
// Everything else is shard-local: controls on global positions are decided per rank (run or
// skip), diagonal selectors on global positions are folded into the table.  The layout is only
// brought back to the identity (dist_plan_canonicalize) when the caller looks at amplitudes.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

#include "classify.h"
#include "planner.h"

namespace qcsim {

struct DistLayout {
  int n = 0, n_local = 0;
  int phys_of[64];  // logical qubit -> physical position
  int log_of[64];   // physical position -> logical qubit

  void reset(int n_, int n_local_) {
    n = n_;
    n_local = n_local_;
    for (int q = 0; q < 64; ++q) phys_of[q] = log_of[q] = q;
  }
  bool is_identity() const {
    for (int q = 0; q < n; ++q)
      if (phys_of[q] != q) return false;
    return true;
  }
  void swap_logical(int a, int b) {
    std::swap(phys_of[a], phys_of[b]);
    log_of[phys_of[a]] = a;
    log_of[phys_of[b]] = b;
  }
  void swap_physical(int x, int y) { swap_logical(log_of[x], log_of[y]); }
  bool is_global(int logical) const { return phys_of[logical] >= n_local; }
};

constexpr int kMaxExchange = 3;     // global positions per exchange = log2 of the largest world (engine.h: kMaxWorld = 8)
constexpr int kMinExchangePos = 5;  // preferred lowest local partner position: runs of >= 2^5 amplitudes = 512 B

struct DistStep {
  bool exchange = false;
  std::vector<Op> ops;  // !exchange: ops on physical LOCAL positions, folded for this rank
  int k = 0;            // exchange: physical position gpos[j] <-> lpos[j] (any distinct local positions)
  int gpos[kMaxExchange] = {0, 0, 0};
  int lpos[kMaxExchange] = {0, 0, 0};
};

namespace detail {

inline bool is_plain_swap(const Op& op) {
  return op.kind == OP_PAIR && op.n_tgt == 2 && op.n_ctrl == 0 && op.m[0] == cplx(0, 0) && op.m[3] == cplx(0, 0) &&
         op.m[1] == cplx(1, 0) && op.m[2] == cplx(1, 0);
}

inline Op physical_swap_op(int x, int y) {
  Op op;
  std::memset(&op, 0, sizeof(op));
  op.kind = OP_PAIR;
  op.n_tgt = 2;
  op.tgt[0] = x;
  op.tgt[1] = y;
  op.m[1] = op.m[2] = cplx(1, 0);
  return op;
}

}  // namespace detail

// Logical op -> op on physical local positions for `rank`.  Returns false when the op does
// nothing on this rank (a control on a global position is 0 here, or the folded table is 1).
// Precondition: every non-diagonal target is on a local position.
inline bool fold_to_physical(const Op& in, const DistLayout& L, int rank, Op* out) {
  Op op = in;
  const int nl = L.n_local;
  auto rank_bit = [&](int p) { return (rank >> (p - nl)) & 1; };
  int nc = 0;
  for (int i = 0; i < in.n_ctrl; ++i) {
    const int p = L.phys_of[in.ctrl[i]];
    if (p >= nl) {
      if (!rank_bit(p)) return false;
    } else {
      op.ctrl[nc++] = p;
    }
  }
  op.n_ctrl = nc;
  for (int i = nc; i < 3; ++i) op.ctrl[i] = 0;
  if (in.kind == OP_DIAG) {
    int nt = 0;
    int keep_pos[3];
    int fixed_mask = 0, fixed_val = 0;
    for (int k = 0; k < in.n_tgt; ++k) {
      const int p = L.phys_of[in.tgt[k]];
      if (p >= nl) {
        fixed_mask |= 1 << k;
        if (rank_bit(p)) fixed_val |= 1 << k;
      } else {
        keep_pos[nt] = k;
        op.tgt[nt++] = p;
      }
    }
    cplx t[8];
    for (int j = 0; j < (1 << nt); ++j) {
      int idx = fixed_val;
      for (int b = 0; b < nt; ++b)
        if ((j >> b) & 1) idx |= 1 << keep_pos[b];
      t[j] = in.m[idx];
    }
    (void)fixed_mask;
    for (int j = 0; j < 8; ++j) op.m[j] = j < (1 << nt) ? t[j] : cplx(0, 0);
    op.n_tgt = nt;
    for (int i = nt; i < 3; ++i) op.tgt[i] = 0;
    bool all_one = true;
    for (int j = 0; j < (1 << nt); ++j) all_one = all_one && detail::is_one(op.m[j]);
    if (all_one) return false;
  } else {
    for (int k = 0; k < in.n_tgt; ++k) op.tgt[k] = L.phys_of[in.tgt[k]];
  }
  *out = op;
  return true;
}

// Plans `ops_in` (logical qubits, program order) for `rank`.  The exchange steps and the layout
// evolution depend only on the op list, never on the rank, so all ranks stay in lock step.
inline std::vector<DistStep> dist_plan(DistLayout& L, const std::vector<Op>& ops_in, int rank) {
  using namespace detail;
  const int nl = L.n_local, n = L.n;
  const int g = n - nl;
  std::vector<DistStep> steps;

  // 1. SWAP gates become relabellings: rewrite the ops that follow onto the pre-swap names
  int sigma[64];
  for (int q = 0; q < 64; ++q) sigma[q] = q;
  std::vector<Op> ops;
  ops.reserve(ops_in.size());
  for (const Op& o : ops_in) {
    if (o.kind == OP_NOP) continue;
    if (is_plain_swap(o)) {
      std::swap(sigma[o.tgt[0]], sigma[o.tgt[1]]);
      continue;
    }
    Op r = o;
    for (int i = 0; i < r.n_ctrl; ++i) r.ctrl[i] = sigma[o.ctrl[i]];
    for (int i = 0; i < r.n_tgt; ++i) r.tgt[i] = sigma[o.tgt[i]];
    ops.push_back(r);
  }
  const int N = (int)ops.size();

  // 2. Rounds of "everything that can run locally, then one exchange".  Walking the pending ops in program order, an
  //    op runs now when every qubit it acts on non-diagonally is local AND it commutes with every op deferred before
  //    it (same rule as plan_passes: two ops commute when on every shared qubit both act diagonally); otherwise it is
  //    deferred.  So an exchange is not a barrier for unrelated gates: later gates on other qubits are pulled in front
  //    of it, which keeps the fused passes full and lets one exchange serve every global qubit the deferred ops need.
  std::vector<OpMasks> mk(N);
  for (int i = 0; i < N; ++i) mk[i] = masks_of(ops[i]);
  std::vector<int> pending(N);
  for (int i = 0; i < N; ++i) pending[i] = i;
  const int kNever = 1 << 30;

  DistStep cur;
  auto flush_local = [&]() {
    if (!cur.ops.empty()) steps.push_back(cur);
    cur = DistStep();
  };

  while (!pending.empty()) {
    uint64_t global_mask = 0;
    for (int q = 0; q < n; ++q)
      if (L.is_global(q)) global_mask |= 1ULL << q;
    uint64_t blocked_nd = 0, blocked_dg = 0;
    std::vector<int> deferred;
    for (int i : pending) {
      const OpMasks& m = mk[i];
      const bool needs_global = (m.nd & global_mask) != 0;
      const bool conflict = ((m.nd | m.dg) & blocked_nd) != 0 || (m.nd & blocked_dg) != 0;
      if (!needs_global && !conflict) {
        Op phys;
        if (fold_to_physical(ops[i], L, rank, &phys)) cur.ops.push_back(phys);
      } else {
        deferred.push_back(i);
        blocked_nd |= m.nd;
        blocked_dg |= m.dg;
      }
    }
    flush_local();
    if (deferred.empty()) break;
    // first non-diagonal use of every logical qubit among the deferred ops
    std::vector<int> first_use(n, kNever);
    for (size_t d = 0; d < deferred.size(); ++d)
      for (int q = 0; q < n; ++q)
        if (((mk[deferred[d]].nd >> q) & 1ULL) && first_use[q] == kNever) first_use[q] = (int)d;
    const uint64_t mandatory = mk[deferred[0]].nd & global_mask;  // the first deferred op is one that needs a global qubit
    // incoming: global-resident qubits by first use; outgoing: local ones, farthest first use first; positions below
    // kMinExchangePos only when nothing else is left (short runs over NVLink); ties: the qubit highest up
    std::vector<std::pair<int, int>> incoming, outgoing;  // (first use, logical)
    for (int q = 0; q < n; ++q) {
      if (L.is_global(q)) incoming.push_back({((mandatory >> q) & 1ULL) ? -1 : first_use[q], q});
      else outgoing.push_back({first_use[q], q});
    }
    std::sort(incoming.begin(), incoming.end());
    const int min_pos = std::min(kMinExchangePos, std::max(0, nl - g - 1));
    std::sort(outgoing.begin(), outgoing.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
      const bool la = L.phys_of[a.second] < min_pos, lb = L.phys_of[b.second] < min_pos;
      if (la != lb) return lb;
      if (a.first != b.first) return a.first > b.first;
      return L.phys_of[a.second] > L.phys_of[b.second];
    });
    int k = 0;
    std::vector<int> in_q, out_q;
    for (size_t j = 0; j < incoming.size() && j < outgoing.size() && (int)j < g && (int)j < kMaxExchange; ++j) {
      const bool must = incoming[j].first < 0;
      if (!must && !(incoming[j].first < outgoing[j].first)) break;
      in_q.push_back(incoming[j].second);
      out_q.push_back(outgoing[j].second);
      ++k;
    }
    DistStep ex;
    ex.exchange = true;
    ex.k = k;
    for (int j = 0; j < k; ++j) {
      ex.gpos[j] = L.phys_of[in_q[j]];
      ex.lpos[j] = L.phys_of[out_q[j]];
    }
    steps.push_back(ex);
    for (int j = 0; j < k; ++j) L.swap_physical(ex.gpos[j], ex.lpos[j]);
    pending.swap(deferred);
  }

  // 3. account for the relabellings: logical q now names the data that was called sigma[q]
  int new_phys[64];
  for (int q = 0; q < n; ++q) new_phys[q] = L.phys_of[sigma[q]];
  for (int q = 0; q < n; ++q) {
    L.phys_of[q] = new_phys[q];
    L.log_of[new_phys[q]] = q;
  }
  return steps;
}

// Steps that bring the layout back to the identity: exchanges that put every global position's own
// qubit back (a position whose qubit sits on another global position first trades with a plain local
// one), then ONE local step of position swaps.
inline std::vector<DistStep> dist_plan_canonicalize(DistLayout& L) {
  using namespace detail;
  const int nl = L.n_local, n = L.n;
  std::vector<DistStep> steps;
  auto do_exchange = [&](const std::vector<int>& gp, const std::vector<int>& lp) {
    DistStep ex;
    ex.exchange = true;
    ex.k = (int)gp.size();
    for (int j = 0; j < ex.k; ++j) {
      ex.gpos[j] = gp[j];
      ex.lpos[j] = lp[j];
    }
    steps.push_back(ex);
    for (int j = 0; j < ex.k; ++j) L.swap_physical(ex.gpos[j], ex.lpos[j]);
  };
  for (int guard = 0; guard < 8; ++guard) {
    std::vector<int> gp, lp;
    bool any_wrong = false;
    for (int p = nl; p < n; ++p) {
      if (L.log_of[p] == p) continue;
      any_wrong = true;
      if (L.phys_of[p] < nl && (int)gp.size() < kMaxExchange) {  // its own qubit is local: bring it up
        gp.push_back(p);
        lp.push_back(L.phys_of[p]);
      }
    }
    if (!any_wrong) break;
    if (gp.empty()) {
      // every wrong global position wants a qubit that sits on another global position: break the cycle by
      // trading one of them against a local position that holds a local qubit (highest such position)
      int p = nl;
      while (L.log_of[p] == p) ++p;
      int x = nl - 1;
      while (x >= 0 && L.log_of[x] >= nl) --x;
      gp.push_back(p);
      lp.push_back(x);
    }
    do_exchange(gp, lp);
  }
  DistStep loc;
  for (int p = 0; p < nl; ++p) {
    if (L.log_of[p] == p) continue;
    loc.ops.push_back(physical_swap_op(p, L.phys_of[p]));
    L.swap_physical(p, L.phys_of[p]);
  }
  if (!loc.ops.empty()) steps.push_back(loc);
  return steps;
}

}  // namespace qcsim
