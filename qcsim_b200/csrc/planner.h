// planner.h -- gate-stream planner for fused gate blocks.  Pure host code (no CUDA), so it is
// unit-tested on the CPU (tests/test_planner.py compiles it with g++).
//
// Input: classified ops (classify.h) whose qubit indices are physical bit positions of the local
// slice.  Output: an ordered list of steps, each either one op run as its own kernel or a fused
// pass = (tile qubit set, ops absorbed into it).  Executing the steps in order is equivalent to
// executing the ops in program order: an op is only moved ahead of a skipped op when the two
// commute, i.e. on every qubit they share both act diagonally (control or diagonal selector).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "classify.h"

namespace qcsim {

struct OpMasks {
  uint64_t nd;  // qubits acted on non-diagonally (must be tile qubits to be fused)
  uint64_t dg;  // qubits acted on diagonally (controls, diagonal selectors): may be anywhere
};

inline OpMasks masks_of(const Op& op) {
  OpMasks m{0, 0};
  for (int i = 0; i < op.n_ctrl; ++i) m.dg |= 1ULL << op.ctrl[i];
  for (int i = 0; i < op.n_tgt; ++i) {
    if (op.kind == OP_DIAG) m.dg |= 1ULL << op.tgt[i];
    else m.nd |= 1ULL << op.tgt[i];
  }
  return m;
}

// HBM bytes per amplitude of the state if the op runs as its own kernel (gate_kernels.cuh)
inline double standalone_cost(const Op& op) {
  double frac = 1.0;
  for (int i = 0; i < op.n_ctrl; ++i)
    if (op.ctrl[i] >= 1) frac *= 0.5;  // a control on bit 0 cannot skip 32-byte sectors
  if (op.kind == OP_PAIR && op.n_tgt == 2) frac *= 0.5;
  return 32.0 * frac;
}

inline int pool_amps_of(const Op& op) {
  switch (op.kind) {
    case OP_PAIR: return 4;
    case OP_DENSE2: return 16;
    case OP_DENSE3: return 64;
    case OP_DIAG: return 8;
    default: return 0;
  }
}

struct PassPlan {
  std::vector<int> tile;  // tile qubits, ascending
  std::vector<int> ops;   // indices into the op list, program order
};

// ---- what a diagonal observable needs from the queue --------------------------------------------
// GetQubitProbability(q) (QubitRegister.h:211-225) is the expectation of a projector that is diagonal in the computational
// basis and acts on qubit q only.  A queued gate changes it only if it acts NON-diagonally on q, or has to run before
// such a gate (it does not commute with it).  Everything else commutes past the needed gates and past the projector:
// it can stay queued and fuse with the gates that arrive later.  Two gates commute here iff they share qubits only
// diagonally (controls / diagonal selectors) -- the rule plan_passes uses to reorder.
// needed / rest keep program order; running `needed` and leaving `rest` queued is the same circuit.
inline void split_queue_for_diagonal_observable(const std::vector<Op>& queue, uint64_t qmask, std::vector<Op>* needed, std::vector<Op>* rest) {
  const size_t n = queue.size();
  std::vector<char> take(n, 0);
  uint64_t s_nd = 0, s_dg = qmask;  // the needed set so far (scanning backwards), seeded with the observable
  for (size_t i = n; i-- > 0;) {
    const OpMasks m = masks_of(queue[i]);
    if (((m.nd | m.dg) & s_nd) != 0 || (m.nd & s_dg) != 0) {
      take[i] = 1;
      s_nd |= m.nd;
      s_dg |= m.dg;
    }
  }
  needed->clear();
  rest->clear();
  for (size_t i = 0; i < n; ++i) (take[i] ? needed : rest)->push_back(queue[i]);
}

struct PlanStep {
  bool fused;
  PassPlan pass;  // !fused: pass.ops holds the single op index
};

// K = tile bits, L = low qubits that are always tile qubits (contiguous 2^L-amplitude runs)
inline std::vector<PlanStep> plan_passes(const std::vector<Op>& ops, int n_local, int K, int L, int max_ops, int max_pool,
                                         int window = 2048, double min_saving = 40.0) {
  const int N = (int)ops.size();
  std::vector<PlanStep> steps;
  std::vector<char> done(N, 0);
  std::vector<OpMasks> mk(N);
  for (int i = 0; i < N; ++i) mk[i] = masks_of(ops[i]);
  const uint64_t all_qubits = (n_local >= 64) ? ~0ULL : ((1ULL << n_local) - 1ULL);
  int first = 0;
  while (first < N) {
    if (done[first] || ops[first].kind == OP_NOP) {
      done[first] = 1;
      ++first;
      continue;
    }
    PassPlan plan;
    uint64_t T = (1ULL << L) - 1ULL;
    int free_slots = K - L;
    uint64_t blocked_nd = 0, blocked_d = 0;
    double saved = 0;
    int pool = 0;
    for (int i = first; i < N && i < first + window; ++i) {
      if (done[i] || ops[i].kind == OP_NOP) continue;
      const OpMasks& m = mk[i];
      const bool conflict = ((m.nd | m.dg) & blocked_nd) != 0 || (m.nd & blocked_d) != 0;
      if (!conflict) {
        const uint64_t need = m.nd & ~T;
        const int c = __builtin_popcountll(need);
        const int pa = pool_amps_of(ops[i]);
        if (c <= free_slots && (int)plan.ops.size() < max_ops && pool + pa <= max_pool) {
          T |= need;
          free_slots -= c;
          pool += pa;
          plan.ops.push_back(i);
          saved += standalone_cost(ops[i]);
          continue;
        }
      }
      blocked_nd |= m.nd;
      blocked_d |= m.dg;
      if ((blocked_nd & all_qubits) == all_qubits) break;
    }
    if (plan.ops.size() < 2 || saved <= min_saving) {
      // not worth a fused pass: run the first pending op on its own
      PlanStep s;
      s.fused = false;
      s.pass.ops.push_back(first);
      steps.push_back(s);
      done[first] = 1;
      ++first;
      continue;
    }
    for (int q = 0; q < n_local && free_slots > 0; ++q)  // pad the tile with the lowest unused qubits
      if (!((T >> q) & 1ULL)) {
        T |= 1ULL << q;
        --free_slots;
      }
    for (int q = 0; q < n_local; ++q)
      if ((T >> q) & 1ULL) plan.tile.push_back(q);
    for (int i : plan.ops) done[i] = 1;
    PlanStep s;
    s.fused = true;
    s.pass = plan;
    steps.push_back(s);
  }
  return steps;
}

// ---- rounds inside a fused pass --------------------------------------------------------------------
// A round = 3 tile bits held in registers (8 amplitudes per thread) + the ops folded into its 8x8
// matrix + up to 3 "variant" qubits (controls / diagonal selectors that are not round bits; each
// doubles the number of matrices of the round).  Ops are taken in program order but may jump ahead
// of ops they commute with (same rule as plan_passes).
struct RoundPlan {
  int rbits[3];            // tile-local bit held by register bit j
  std::vector<int> vq;     // variant qubits (register qubit numbers), <= 3
  int item_bit[9];         // tile-local bit walked by item-index bit j (first k - 3 entries valid)
  std::vector<int> ops;    // indices into the op list, program order
  bool chain_next = false; // (swizzle_kind 2) the next round uses the same warp-index bits and none of them is a register
                           // bit of either round: every warp keeps working on its own amplitudes, no group barrier needed
};

// swizzle_kind selects the shared-memory slot swizzle the low item bits must dodge: 0 = swz() of
// common.cuh (every tile bit folds onto the low three, class = bit mod 3), 1 = the 128 B TMA swizzle
// (only slot bits 3..5 fold onto 0..2) with one item per lane, 2 = the TMA swizzle with the DMMA
// fragment mapping of tile_pipe.cuh (register-bit order and item bits 0..2 chosen together).
// What round formation decides, independent of the order of the tile qubits: register qubits (before padding with
// unused tile qubits), variant qubits, ops.  schedule_rounds can hand it out / take it back, so that the same rounds are
// placed on several shared-memory layouts (schedule_rounds_best_layout) without searching again.
struct FormedRound {
  uint64_t R = 0, V = 0;
  std::vector<int> ops;
};

inline std::vector<RoundPlan> schedule_rounds(const std::vector<Op>& all, const PassPlan& plan, int max_variant_bits = 3,
                                              int swizzle_kind = 0, int* conflict_cost = nullptr, bool always_chain = false,
                                              const std::vector<FormedRound>* preformed = nullptr, std::vector<FormedRound>* formed_out = nullptr) {
  int total_cost = 0;  // swizzle_kind 2: sum over rounds of (load degree + store degree); 2 per round = conflict-free
  const int k = (int)plan.tile.size();
  int local_of[64];
  for (int q = 0; q < 64; ++q) local_of[q] = -1;
  for (int j = 0; j < k; ++j) local_of[plan.tile[j]] = j;

  std::vector<int> remaining = plan.ops;
  std::vector<RoundPlan> rounds;
  std::vector<uint32_t> used_of, vtile_of;  // per round: register bits / in-tile variant bits (tile-local masks)
  std::vector<FormedRound> formed;
  if (preformed) {
    formed = *preformed;
    remaining.clear();
  }
  while (!remaining.empty()) {
    // Greedy fill from a seed: the seed's targets become register bits first, then ops are taken in
    // program order.  Several seeds are tried (the first pending op, and the next few ops that
    // commute with everything before them); the fill that absorbs the most ops wins.
    struct Fill {
      uint64_t R = 0, V = 0;
      std::vector<int> ops, left;
    };
    auto fill_from = [&](int seed_pos) {
      Fill f;
      uint64_t blocked_nd = 0, blocked_d = 0;
      if (seed_pos > 0) {  // reserve the seed's register bits before scanning
        const OpMasks ms = masks_of(all[remaining[seed_pos]]);
        f.R = ms.nd;
        f.V = ms.dg & ~f.R;
      }
      for (size_t pos = 0; pos < remaining.size(); ++pos) {
        const int idx = remaining[pos];
        const OpMasks m = masks_of(all[idx]);
        const bool conflict = ((m.nd | m.dg) & blocked_nd) != 0 || (m.nd & blocked_d) != 0;
        if (!conflict) {
          const uint64_t nR = f.R | m.nd;
          const uint64_t nV = (f.V | m.dg) & ~nR;
          if (__builtin_popcountll(nR) <= 3 && __builtin_popcountll(nV) <= max_variant_bits) {
            f.R = nR;
            f.V = nV;
            f.ops.push_back(idx);
            continue;
          }
        }
        f.left.push_back(idx);
        blocked_nd |= m.nd;
        blocked_d |= m.dg;
      }
      return f;
    };
    Fill best = fill_from(0);
    {
      uint64_t blocked_nd = 0, blocked_d = 0;
      int tried = 0;
      for (size_t pos = 0; pos < remaining.size() && pos < 64 && tried < 6; ++pos) {
        const OpMasks m = masks_of(all[remaining[pos]]);
        const bool free_to_lead = ((m.nd | m.dg) & blocked_nd) == 0 && (m.nd & blocked_d) == 0;
        if (pos > 0 && free_to_lead && m.nd != 0) {
          ++tried;
          Fill f = fill_from((int)pos);
          // the seed must actually have been absorbed, and the first pending op must not starve
          if (f.ops.size() > best.ops.size()) best = f;
        }
        blocked_nd |= m.nd;
        blocked_d |= m.dg;
      }
    }
    // Then every triple of the pending ops' target qubits as the register set: the greedy seeds above commit to the
    // first ops they meet, a fixed triple absorbs everything that fits it (11 % fewer rounds on the benchmark circuit).
    {
      constexpr size_t kTripleWindow = 192;
      auto fill_fixed = [&](uint64_t R0) {
        Fill f;
        f.R = R0;
        uint64_t blocked_nd = 0, blocked_d = 0;
        for (size_t pos = 0; pos < remaining.size(); ++pos) {
          const int idx = remaining[pos];
          if (pos >= kTripleWindow) {  // bounded look-ahead: 165 triples x the whole pass would be quadratic in deep passes
            f.left.push_back(idx);
            continue;
          }
          const OpMasks m = masks_of(all[idx]);
          const bool conflict = ((m.nd | m.dg) & blocked_nd) != 0 || (m.nd & blocked_d) != 0;
          if (!conflict && (m.nd & ~R0) == 0) {
            const uint64_t nV = (f.V | m.dg) & ~R0;
            if (__builtin_popcountll(nV) <= max_variant_bits) {
              f.V = nV;
              f.ops.push_back(idx);
              continue;
            }
          }
          f.left.push_back(idx);
          blocked_nd |= m.nd;
          blocked_d |= m.dg;
        }
        return f;
      };
      uint64_t cand = 0;
      for (size_t pos = 0; pos < remaining.size() && pos < kTripleWindow; ++pos) cand |= masks_of(all[remaining[pos]]).nd;
      int cq[64], nc = 0;
      for (int q = 0; q < 64; ++q)
        if ((cand >> q) & 1ULL) cq[nc++] = q;
      bool replaced = false;
      for (int a = 0; a < nc; ++a)
        for (int b = a + 1; b < nc; ++b)
          for (int c = b + 1; c < nc; ++c) {
            Fill f = fill_fixed((1ULL << cq[a]) | (1ULL << cq[b]) | (1ULL << cq[c]));
            if (f.ops.size() > best.ops.size()) {
              best = f;
              replaced = true;
            }
          }
      (void)replaced;  // best.R is the whole triple (a qubit none of the absorbed ops turns is an identity factor), best.V as counted
    }
    uint64_t R = best.R, V = best.V;
    // free register slots: promote variant qubits that live in the tile (halves the matrix count for free)
    for (int q = 0; q < 64 && __builtin_popcountll(R) < 3; ++q)
      if (((V >> q) & 1ULL) && local_of[q] >= 0) {
        R |= 1ULL << q;
        V &= ~(1ULL << q);
      }
    FormedRound fr;
    fr.R = R;
    fr.V = V;
    fr.ops = best.ops;
    formed.push_back(fr);
    std::vector<int> left = best.left;
    remaining.swap(left);
  }
  if (formed_out) *formed_out = formed;
  for (const FormedRound& fr : formed) {
    uint64_t R = fr.R;
    const uint64_t V = fr.V;
    RoundPlan rp;
    rp.ops = fr.ops;
    // pad the register set with unused tile bits (high slots first: the low ones serve the lanes)
    for (int j = k - 1; j >= 0 && __builtin_popcountll(R) < 3; --j) R |= 1ULL << plan.tile[j];
    int nr = 0;
    uint32_t used = 0;
    for (int q = 0; q < 64; ++q)
      if ((R >> q) & 1ULL) {
        rp.rbits[nr++] = local_of[q];
        used |= 1u << local_of[q];
      }
    for (int q = 0; q < 64; ++q)
      if ((V >> q) & 1ULL) rp.vq.push_back(q);
    used_of.push_back(used);
    uint32_t vt = 0;
    for (int lb = 0; lb < k; ++lb)
      if (!((used >> lb) & 1u) && ((V >> plan.tile[lb]) & 1ULL)) vt |= 1u << lb;
    vtile_of.push_back(vt);
    rounds.push_back(rp);
  }

  // item-index bits.  Variant qubits inside the tile go to the warp-index part of the item index
  // (bits 5..7: warp-uniform matrix choice; bit 8 pairs the two items a thread of k_tile_pass processes with
  // one set of matrix loads, so it must not select the matrix); the low item bits are chosen against the
  // shared-memory swizzle (see below).  forced_warp != 0: exactly these three tile bits are the warp bits.
  auto assign_items = [&](RoundPlan& rp, uint32_t used, uint32_t vt, uint32_t forced_warp) -> int {
    int round_cost = 0;
    const int ni = k - 3;
    int order[16];
    for (int j = 0; j < 16; ++j) order[j] = -1;
    int top = ni >= 9 ? 7 : ni - 1;
    std::vector<int> rest;
    for (int lb = 0; lb < k; ++lb) {
      if ((used >> lb) & 1u) continue;
      if (forced_warp ? ((forced_warp >> lb) & 1u) : ((vt >> lb) & 1u)) order[top--] = lb;
      else rest.push_back(lb);
    }
    int pos = 0;
    auto next_free = [&]() {
      while (order[pos] >= 0) ++pos;
      return pos;
    };
    if (swizzle_kind == 2) {
      // DMMA rounds on the TMA layout (tile_pipe.cuh).  Fragments move as 64-bit halves (LDS.64 / STS.64; odd
      // k-lanes / odd rows take the imaginary half first), so a half-warp's 16 lanes must hit 16 distinct
      // 8-byte columns of a 128 B row: column = 2 * bank(slot) + half.  The lanes of a half-warp differ in
      //   fragment loads : item bits 0,1 (4 columns of the panel) x register bits 0,1 (the 4 k-lanes; bit 0 also picks the half)
      //   result stores  : register bits 0,1 (4 rows; bit 0 also picks the half) x item bits 1,2
      // Pick the order of the three register bits and the item bits (i0,i1,i2) with the fewest conflicts under tswz
      // (only slot bits 0..5 move the bank: bank = (s ^ s >> 3) & 7).
      // Under tswz only slot bits 0..5 move the bank, each by one bit: code(s) = unit vector (s mod 3) for s < 6, else 0.
      // The 16 lanes map linearly (XOR) onto 16 columns; the half-selecting lane bit is independent of the rest, so the
      // conflict degree is 2^(3 - rank) with rank = number of distinct non-zero codes among the other three lane bits:
      //   loads : {r1, i0, i1}      stores : {r1, i1, i2}
      // r1 is one of the three register bits, the items come from `rest`; only the codes matter, so the search runs
      // over code classes (4^3 x 3 candidates) instead of slots.
      auto code = [](int s) -> uint32_t { return s < 3 ? 1u << s : (s < 6 ? 1u << (s - 3) : 0u); };
      auto cls = [](int s) { return s < 6 ? s % 3 : 3; };  // 3 = no bank bit
      const uint32_t cls_code[4] = {1u, 2u, 4u, 0u};
      std::vector<int> by_cls[4];
      for (int lb : rest) by_cls[cls(lb)].push_back(lb);  // ascending
      int best_cost = 1 << 30, best_r1 = 0, best_c[3] = {3, 3, 3};
      const int nrest = (int)rest.size();
      for (int j1 = 0; j1 < 3 && nrest >= 3; ++j1) {
        const uint32_t c1 = code(rp.rbits[j1]);
        for (int ca = 0; ca < 4; ++ca)
          for (int cb = 0; cb < 4; ++cb)
            for (int cc = 0; cc < 4; ++cc) {
              int need[4] = {0, 0, 0, 0};
              ++need[ca];
              ++need[cb];
              ++need[cc];
              if (need[0] > (int)by_cls[0].size() || need[1] > (int)by_cls[1].size() || need[2] > (int)by_cls[2].size() ||
                  need[3] > (int)by_cls[3].size())
                continue;
              const int cost = (1 << (3 - __builtin_popcount(c1 | cls_code[ca] | cls_code[cb]))) +
                               (1 << (3 - __builtin_popcount(c1 | cls_code[cb] | cls_code[cc])));
              if (cost < best_cost) {
                best_cost = cost;
                best_r1 = j1;
                best_c[0] = ca;
                best_c[1] = cb;
                best_c[2] = cc;
              }
            }
      }
      int best_r[3], best_i[3] = {-1, -1, -1};
      {  // register order: r1 = the chosen bit, r0 / r2 = the other two in their given order
        int o = 0, others[2];
        for (int j = 0; j < 3; ++j)
          if (j != best_r1) others[o++] = rp.rbits[j];
        best_r[0] = others[0];
        best_r[1] = rp.rbits[best_r1];
        best_r[2] = others[1];
      }
      if (best_cost < (1 << 30)) {
        size_t taken[4] = {0, 0, 0, 0};
        for (int j = 0; j < 3; ++j) {
          const int lb = by_cls[best_c[j]][taken[best_c[j]]++];
          for (int i = 0; i < nrest; ++i)
            if (rest[i] == lb) best_i[j] = i;
        }
        round_cost = best_cost;
      }
      for (int j = 0; j < 3; ++j) rp.rbits[j] = best_r[j];
      if (best_i[0] >= 0)
        for (int j = 0; j < 3; ++j) {
          order[next_free()] = rest[best_i[j]];
          rest[best_i[j]] = -1;
        }
    } else {
      for (int res = 0; res < 3; ++res)
        for (size_t i = 0; i < rest.size(); ++i)
          if (rest[i] >= 0 && rest[i] % 3 == res && (swizzle_kind == 0 || rest[i] < 6)) {
            order[next_free()] = rest[i];
            rest[i] = -1;
            break;
          }
    }
    for (int lb : rest)
      if (lb >= 0) order[next_free()] = lb;
    for (int j = 0; j < 9; ++j) rp.item_bit[j] = (j < ni && order[j] >= 0) ? order[j] : 0;
    return round_cost;
  };

  const int n_rounds = (int)rounds.size();
  if (swizzle_kind == 2 && k - 3 == 8) {
    // Chains of rounds that share their three warp-index bits: the bits are register bits of none of the rounds and
    // contain every round's in-tile variant bits, so a warp reads and writes the same 256 amplitudes round after
    // round -- the rounds of a chain are separated by __syncwarp() instead of the group barrier.
    const uint32_t all_bits = (1u << k) - 1u;
    int i = 0;
    while (i < n_rounds) {
      uint32_t F = all_bits & ~used_of[i], Nd = vtile_of[i];
      int j = i + 1;
      while (j < n_rounds) {
        const uint32_t F2 = F & ~used_of[j], N2 = Nd | vtile_of[j];
        if (__builtin_popcount(N2) > 3 || (N2 & ~F2) != 0 || __builtin_popcount(F2) < 3) break;
        F = F2;
        Nd = N2;
        ++j;
      }
      if (j - i >= 2) {
        uint32_t W = Nd;
        for (int lb = k - 1; lb >= 0 && __builtin_popcount(W) < 3; --lb)  // high slot bits first: the low ones serve the lanes
          if (((F >> lb) & 1u) && !((W >> lb) & 1u)) W |= 1u << lb;
        // a chain saves group barriers but pins the warp bits; when that costs bank conflicts (variant qubits on
        // the low slots leave too few bank-moving bits for the lanes) the rounds run unchained
        std::vector<RoundPlan> chained(rounds.begin() + i, rounds.begin() + j), apart(rounds.begin() + i, rounds.begin() + j);
        int cost_chained = 0, cost_apart = 0;
        for (int r = i; r < j; ++r) {
          cost_chained += assign_items(chained[r - i], used_of[r], vtile_of[r], W);
          chained[r - i].chain_next = r + 1 < j;
          cost_apart += assign_items(apart[r - i], used_of[r], vtile_of[r], 0);
        }
        const bool keep_chain = always_chain || cost_chained <= cost_apart;
        const std::vector<RoundPlan>& pick = keep_chain ? chained : apart;
        for (int r = i; r < j; ++r) rounds[r] = pick[r - i];
        total_cost += keep_chain ? cost_chained : cost_apart;
      } else {
        total_cost += assign_items(rounds[i], used_of[i], vtile_of[i], 0);
      }
      i = j;
    }
  } else {
    for (int r = 0; r < n_rounds; ++r) total_cost += assign_items(rounds[r], used_of[r], vtile_of[r], 0);
  }
  if (conflict_cost) *conflict_cost = total_cost;
  return rounds;
}

// ---- TMA tile geometry (tile_pipe.cuh) ---------------------------------------------------------------
// The state vector seen as a rank-5 tensor of doubles whose dimensions are ranges of index bits:
// dim 0 = bits [0,3) (8 amplitudes = 128 B, always boxed whole), dims 1.. = consecutive bit ranges
// [lo_i, lo_{i+1}) whose LOW box_bits[i] bits are tile qubits (the TMA box spans them) and whose
// remaining bits are addressed by the coordinate.  Tile qubits that no dimension boxes are
// "enumerated": one TMA op per value.  The shared-memory slot order of the tile follows: qubits
// 0..2, the boxed groups in dimension order, the enumerated qubits.
struct TmaTileGeom {
  int k = 0;
  int slot_qubit[16];  // index bit held by shared-memory slot bit j
  int n_dims = 0;      // dimensions in use (<= 5); the kernel always passes 5 coordinates
  int dim_lo[5];
  int dim_bits[5];     // log2 of the extent in amplitudes (dim 0: 3)
  int box_bits[5];     // log2 of the box in amplitudes (dim 0: 3)
  int n_enum = 0;
  int enum_pos[16];
  int box_log2 = 0;    // log2(amplitudes per TMA op)
};

// tile: ascending tile qubits.  Needs qubits 0,1,2 in the tile and at least 3 more.
inline bool tma_tile_geometry(const std::vector<int>& tile, int n_local, TmaTileGeom* g) {
  const int k = (int)tile.size();
  if (k < 6 || k > 16 || n_local < k) return false;
  for (int j = 0; j < 3; ++j)
    if (tile[j] != j) return false;
  struct Grp { int lo, w; };
  std::vector<Grp> groups;  // contiguous runs of tile qubits >= 3, cut at width 8 (box extents are <= 256)
  for (int j = 3; j < k; ++j) {
    if (!groups.empty() && groups.back().lo + groups.back().w == tile[j] && groups.back().w < 8) ++groups.back().w;
    else groups.push_back({tile[j], 1});
  }
  std::vector<Grp> boxed;
  std::vector<char> taken(groups.size(), 0);
  if (groups[0].lo == 3) {
    boxed.push_back(groups[0]);
    taken[0] = 1;
  } else {
    boxed.push_back({3, 0});  // dim 1 must start at bit 3; nothing of it is boxed
  }
  for (int pick = 0; pick < 3; ++pick) {  // up to three more groups, widest first
    int best = -1;
    for (size_t i = 0; i < groups.size(); ++i)
      if (!taken[i] && (best < 0 || groups[i].w > groups[best].w)) best = (int)i;
    if (best < 0) break;
    taken[best] = 1;
    boxed.push_back(groups[best]);
  }
  std::sort(boxed.begin(), boxed.end(), [](const Grp& a, const Grp& b) { return a.lo < b.lo; });
  g->k = k;
  g->n_dims = 1 + (int)boxed.size();
  g->dim_lo[0] = 0;
  g->dim_bits[0] = 3;
  g->box_bits[0] = 3;
  int ns = 0;
  for (int j = 0; j < 3; ++j) g->slot_qubit[ns++] = j;
  g->box_log2 = 3;
  for (size_t i = 0; i < boxed.size(); ++i) {
    const int hi = i + 1 < boxed.size() ? boxed[i + 1].lo : n_local;
    g->dim_lo[1 + i] = boxed[i].lo;
    g->dim_bits[1 + i] = hi - boxed[i].lo;
    g->box_bits[1 + i] = boxed[i].w;
    if (g->dim_bits[1 + i] < boxed[i].w || g->dim_bits[1 + i] > 31) return false;
    for (int b = 0; b < boxed[i].w; ++b) g->slot_qubit[ns++] = boxed[i].lo + b;
    g->box_log2 += boxed[i].w;
  }
  for (int d = g->n_dims; d < 5; ++d) {
    g->dim_lo[d] = n_local;
    g->dim_bits[d] = 0;
    g->box_bits[d] = 0;
  }
  g->n_enum = 0;
  for (size_t i = 0; i < groups.size(); ++i)
    if (!taken[i])
      for (int b = 0; b < groups[i].w; ++b) {
        g->enum_pos[g->n_enum++] = groups[i].lo + b;
        g->slot_qubit[ns++] = groups[i].lo + b;
      }
  if (ns != k || g->box_log2 + g->n_enum != k) return false;
  if (g->box_log2 < 6) return false;  // every op must land on a 1024 B boundary (128 B swizzle atom = 8 rows)
  return true;
}

// The same geometry with tensor dimensions 1.. in another order (every dimension carries its own stride, so any order
// is a valid tensor map).  order[i] = dimension of `base` that becomes dimension 1 + i, for i < base.n_dims - 1.
// Only the shared-memory slot order of the boxed qubits changes: the box is laid out dimension by dimension.
inline TmaTileGeom tma_reorder_dims(const TmaTileGeom& base, const int* order) {
  TmaTileGeom g = base;
  int ns = 3;
  for (int i = 0; i + 1 < base.n_dims; ++i) {
    const int d = order[i];
    g.dim_lo[1 + i] = base.dim_lo[d];
    g.dim_bits[1 + i] = base.dim_bits[d];
    g.box_bits[1 + i] = base.box_bits[d];
    for (int b = 0; b < base.box_bits[d]; ++b) g.slot_qubit[ns++] = base.dim_lo[d] + b;
  }
  for (int j = 0; j < base.n_enum; ++j) g.slot_qubit[ns++] = base.enum_pos[j];
  return g;
}

// Rounds of a pass on the TMA pipeline (swizzle_kind 2) with the dimension order whose slot layout gives the DMMA
// rounds the fewest shared-memory bank conflicts: slot bits 3..5 are the only ones besides 0..2 that move the bank
// (tswz), so which boxed qubits land there decides whether a round on high qubits can be conflict-free.
// plan_sorted.tile is ascending; *plan_slots receives the plan with the tile in the chosen slot order.
inline std::vector<RoundPlan> schedule_rounds_best_layout(const std::vector<Op>& all, const PassPlan& plan_sorted, const TmaTileGeom& base,
                                                          int max_variant_bits, TmaTileGeom* chosen, PassPlan* plan_slots, int* cost_out = nullptr,
                                                          bool always_chain = false, bool search = true) {
  const int k = base.k, nd = base.n_dims - 1;
  int boxed[4], n_boxed = 0, plain[4], n_plain = 0;
  for (int d = 1; d <= nd; ++d) {
    if (base.box_bits[d] > 0) boxed[n_boxed++] = d;
    else plain[n_plain++] = d;
  }
  std::vector<RoundPlan> best_rounds;
  int best_cost = 1 << 30;
  std::vector<FormedRound> formed;  // the rounds are formed once (on the ascending order) and placed on every candidate layout
  bool have_formed = false;
  auto consider = [&](const TmaTileGeom& g) {  // true: nothing can beat it
    PassPlan plan = plan_sorted;
    for (int j = 0; j < k; ++j) plan.tile[j] = g.slot_qubit[j];
    int cost = 0;
    std::vector<RoundPlan> rounds = schedule_rounds(all, plan, max_variant_bits, 2, &cost, always_chain, have_formed ? &formed : nullptr,
                                                    have_formed ? nullptr : &formed);
    have_formed = true;
    if (cost < best_cost) {
      best_cost = cost;
      best_rounds.swap(rounds);
      *chosen = g;
      *plan_slots = plan;
    }
    return best_cost <= 2 * (int)best_rounds.size();  // every round conflict-free
  };
  bool done = consider(base);  // the ascending order first: it wins ties
  int perm[4] = {0, 1, 2, 3};
  while (!done && search) {
    int order[4];
    for (int i = 0; i < n_boxed; ++i) order[i] = boxed[perm[i]];
    for (int i = 0; i < n_plain; ++i) order[n_boxed + i] = plain[i];
    const TmaTileGeom g = tma_reorder_dims(base, order);
    bool same = true;
    for (int j = 0; j < k; ++j) same = same && g.slot_qubit[j] == base.slot_qubit[j];
    if (!same) done = consider(g);
    if (!std::next_permutation(perm, perm + n_boxed)) break;
  }
  if (cost_out) *cost_out = best_cost;
  return best_rounds;
}

// ---- round matrices ---------------------------------------------------------------------------------
// Applies `op` to the 8 amplitudes spanned by a round's bits.  rpos[q] = register bit of qubit q
// (-1: not a round qubit, then vval[q] is its value for this matrix variant).
inline void apply_small(const Op& op, const int* rpos, const int* vval, cplx* vec) {
  uint32_t cmask = 0;
  for (int i = 0; i < op.n_ctrl; ++i) {
    const int q = op.ctrl[i];
    if (rpos[q] >= 0) cmask |= 1u << rpos[q];
    else if (!vval[q]) return;  // this variant has the control at 0: identity
  }
  auto live = [&](uint32_t x) { return (x & cmask) == cmask; };
  switch (op.kind) {
    case OP_PAIR: {
      const uint32_t b0 = 1u << rpos[op.tgt[0]];
      const uint32_t b1 = op.n_tgt == 2 ? 1u << rpos[op.tgt[1]] : 0u;
      for (uint32_t x = 0; x < 8; ++x) {
        if ((x & (b0 | b1)) || !live(x)) continue;
        const uint32_t ia = op.n_tgt == 1 ? x : (x | b0), ib = op.n_tgt == 1 ? (x | b0) : (x | b1);
        const cplx a = vec[ia], b = vec[ib];
        vec[ia] = op.m[0] * a + op.m[1] * b;
        vec[ib] = op.m[2] * a + op.m[3] * b;
      }
      break;
    }
    case OP_DENSE2:
    case OP_DENSE3: {
      const int nt = op.kind == OP_DENSE2 ? 2 : 3, d = 1 << nt;
      uint32_t tm = 0, off[8];
      for (int j = 0; j < nt; ++j) tm |= 1u << rpos[op.tgt[j]];
      for (int c = 0; c < d; ++c) {
        off[c] = 0;
        for (int j = 0; j < nt; ++j)
          if ((c >> j) & 1) off[c] |= 1u << rpos[op.tgt[j]];
      }
      for (uint32_t x = 0; x < 8; ++x) {
        if ((x & tm) || !live(x)) continue;
        cplx in[8], out[8];
        for (int c = 0; c < d; ++c) in[c] = vec[x | off[c]];
        for (int r = 0; r < d; ++r) {
          cplx acc = op.m[r * d] * in[0];
          for (int c = 1; c < d; ++c) acc += op.m[r * d + c] * in[c];
          out[r] = acc;
        }
        for (int c = 0; c < d; ++c) vec[x | off[c]] = out[c];
      }
      break;
    }
    case OP_DIAG:
      for (uint32_t x = 0; x < 8; ++x) {
        if (!live(x)) continue;
        int idx = 0;
        for (int s = 0; s < op.n_tgt; ++s) {
          const int q = op.tgt[s];
          const int bit = rpos[q] >= 0 ? (int)((x >> rpos[q]) & 1u) : vval[q];
          idx |= bit << s;
        }
        vec[x] *= op.m[idx];
      }
      break;
    default: break;
  }
}

// The 2^v matrices of a round (v = rp.vq.size()): matrix a is the product, in program order, of the
// round's ops with the variant qubits fixed to the bits of a.  Row-major 8x8, register bit j of the
// row/column index <-> tile bit rp.rbits[j].  `out` holds 64 << v entries.
inline void build_round_matrices(const std::vector<Op>& all, const PassPlan& plan, const RoundPlan& rp, cplx* out) {
  int rpos[64];
  for (int q = 0; q < 64; ++q) rpos[q] = -1;
  for (int j = 0; j < 3; ++j) rpos[plan.tile[rp.rbits[j]]] = j;
  const int nv = (int)rp.vq.size();
  for (int a = 0; a < (1 << nv); ++a) {
    int vval[64];
    for (int q = 0; q < 64; ++q) vval[q] = 0;
    for (int j = 0; j < nv; ++j) vval[rp.vq[j]] = (a >> j) & 1;
    cplx* M = out + (size_t)a * 64;
    for (int col = 0; col < 8; ++col) {
      cplx vec[8];
      for (int x = 0; x < 8; ++x) vec[x] = cplx(x == col ? 1.0 : 0.0, 0.0);
      for (int idx : rp.ops) apply_small(all[idx], rpos, vval, vec);
      for (int row = 0; row < 8; ++row) M[row * 8 + col] = vec[row];
    }
  }
}

// ---- round descriptor (what the tile kernels read per round) -----------------------------------------
// Plain-integer image of TileRoundDesc (tile_kernels.cuh); built here so the CPU tests see exactly what
// the kernels are given.
struct RoundDescHost {
  uint32_t rb;       // rb0 | rb1 << 8 | rb2 << 16 : tile bit held by register bit j
  uint32_t tb[3];    // tile bit walked by item-index bit j, one byte each
  uint32_t var;      // nvar | chain flag << 7 | (src | pos << 1) << (8 + 8 j): src 0 = tile bit, 1 = index bit outside the tile
  uint32_t mat_off;  // first matrix of the round
};

// local_of[q] = tile bit of qubit q (-1: outside the tile)
inline RoundDescHost make_round_desc(const RoundPlan& rp, const int* local_of, uint32_t mat_index) {
  RoundDescHost rd;
  rd.rb = (uint32_t)(rp.rbits[0] | (rp.rbits[1] << 8) | (rp.rbits[2] << 16));
  for (int w = 0; w < 3; ++w) rd.tb[w] = 0;
  for (int j = 0; j < 9; ++j) rd.tb[j >> 2] |= (uint32_t)rp.item_bit[j] << (8 * (j & 3));
  const int nv = (int)rp.vq.size();
  rd.var = (uint32_t)nv | (rp.chain_next ? 0x80u : 0u);  // bit 7: the next round needs no group barrier (same warp bits)
  for (int j = 0; j < nv; ++j) {
    const int q = rp.vq[j];
    const uint32_t e = local_of[q] >= 0 ? (uint32_t)(local_of[q] << 1) : (uint32_t)((q << 1) | 1);
    rd.var |= e << (8 + 8 * j);
  }
  rd.mat_off = mat_index;
  return rd;
}

// ---- QFT recognition ---------------------------------------------------------------------------------
// QCSim's own QuantumFourierTransform (QuantumFourierTransform.h:35-87) reaches the engine as a
// stream of H / ControlledPhaseShift / SWAP gates.  The exact stream -- same order, the reference's
// own phases (M_PI_2 halved repeatedly through std::polar) -- is recognised here so that it runs
// as radix-8 FFT passes (qft_kernels.cuh) instead of hundreds of diagonal gates.
struct QftMatch {
  int length = 0;  // ops consumed (0: no match)
  int sq = 0, eq = 0;
  bool do_swap = false, inverse = false;
};

namespace detail {
inline bool is_hadamard_op(const Op& op, int* q) {
  if (op.kind != OP_PAIR || op.n_tgt != 1 || op.n_ctrl != 0) return false;
  const double s = 1. / std::sqrt(2.);
  if (!(op.m[0] == cplx(s, 0) && op.m[1] == cplx(s, 0) && op.m[2] == cplx(s, 0) && op.m[3] == cplx(-s, 0))) return false;
  *q = op.tgt[0];
  return true;
}
inline bool is_cphase_op(const Op& op, int a, int b, double theta) {
  if (op.kind != OP_DIAG || op.n_tgt != 0 || op.n_ctrl != 2) return false;
  if (!((op.ctrl[0] == a && op.ctrl[1] == b) || (op.ctrl[0] == b && op.ctrl[1] == a))) return false;
  return op.m[0] == cplx(std::cos(theta), std::sin(theta));  // std::polar(1., theta), QuantumGate.h:262-265
}
inline bool is_swap_op(const Op& op, int a, int b) {
  if (op.kind != OP_PAIR || op.n_tgt != 2 || op.n_ctrl != 0) return false;
  if (!(op.m[0] == cplx(0, 0) && op.m[3] == cplx(0, 0) && op.m[1] == cplx(1, 0) && op.m[2] == cplx(1, 0))) return false;
  return (op.tgt[0] == a && op.tgt[1] == b) || (op.tgt[0] == b && op.tgt[1] == a);
}
// swaps (sq,eq), (sq+1,eq-1) ... starting at ops[i]; returns ops consumed or -1
inline int match_swaps(const std::vector<Op>& ops, size_t i, int sq, int eq) {
  int used = 0;
  for (int s = sq, e = eq; s < e; ++s, --e, ++used)
    if (i + used >= ops.size() || !is_swap_op(ops[i + used], s, e)) return -1;
  return used;
}
}  // namespace detail

// Tries to match a QFT / IQFT of at least `min_qubits` qubits starting exactly at ops[i].
// *ran_off_end (optional) is set when the stream was still a valid QFT prefix at the end of `ops`.
inline QftMatch match_qft(const std::vector<Op>& ops, size_t i, int min_qubits = 4, bool* ran_off_end = nullptr) {
  using namespace detail;
  QftMatch none;
  if (ran_off_end) *ran_off_end = false;
  auto off = [&](size_t p) {  // true when position p is past the end (and remembers it)
    if (p >= ops.size() && ran_off_end) *ran_off_end = true;
    return p >= ops.size();
  };
  const double pi_2 = 1.57079632679489661923;  // M_PI_2
  const size_t N = ops.size();
  int q0 = 0;
  // ---- forward: H(eq) [CP(eq, eq-1 .. sq)] H(eq-1) ... H(sq) [swaps]
  if (i < N && is_hadamard_op(ops[i], &q0)) {
    const int eq = q0;
    // the run of controlled phases after the first H fixes sq
    size_t j = i + 1;
    double phase = pi_2;
    int ctrl = eq - 1;
    while (!off(j) && ctrl >= 0 && is_cphase_op(ops[j], eq, ctrl, phase)) {
      ++j;
      --ctrl;
      phase *= 0.5;
    }
    const int sq = ctrl + 1;
    if (eq - sq + 1 >= min_qubits) {
      bool ok = true;
      size_t p = i + 1;
      for (int cur = eq; cur > sq && ok; --cur) {
        double ph = pi_2;
        for (int c = cur - 1; c >= sq && ok; --c) {
          ok = !off(p) && is_cphase_op(ops[p], cur, c, ph);
          ++p;
          ph *= 0.5;
        }
        int hq = -1;
        ok = ok && !off(p) && is_hadamard_op(ops[p], &hq) && hq == cur - 1;
        ++p;
      }
      if (ok) {
        QftMatch m;
        m.sq = sq;
        m.eq = eq;
        m.inverse = false;
        const int sw = match_swaps(ops, p, sq, eq);
        m.do_swap = sw > 0;
        m.length = (int)(p - i) + (sw > 0 ? sw : 0);
        return m;
      }
    }
  }
  // ---- inverse: [swaps] H(sq) CP(sq+1, sq) H(sq+1) CP(sq+2, sq+1) CP(sq+2, sq) ... H(eq)
  {
    size_t p = i;
    int n_sw = 0, sw_s = -1, sw_e = -1;
    // leading swaps, if any, fix (sq, eq)
    if (p < N && ops[p].kind == OP_PAIR && ops[p].n_tgt == 2 && ops[p].n_ctrl == 0) {
      sw_s = std::min(ops[p].tgt[0], ops[p].tgt[1]);
      sw_e = std::max(ops[p].tgt[0], ops[p].tgt[1]);
      const int used = match_swaps(ops, p, sw_s, sw_e);
      if (used <= 0) return none;
      n_sw = used;
      p += used;
    }
    int sq = 0;
    if (!(p < N && is_hadamard_op(ops[p], &sq))) return none;
    if (n_sw && sq != sw_s) return none;
    ++p;
    int cur = sq + 1;
    for (;;) {
      // CP(cur, cur-1 .. sq) with phases -pi/2, -pi/4 ...
      size_t pp = p;
      double ph = -pi_2;
      bool ok = true;
      for (int c = cur - 1; c >= sq && ok; --c) {
        ok = !off(pp) && is_cphase_op(ops[pp], cur, c, ph);
        ++pp;
        ph *= 0.5;
      }
      int hq = -1;
      ok = ok && !off(pp) && is_hadamard_op(ops[pp], &hq) && hq == cur;
      if (!ok) break;
      p = pp + 1;
      ++cur;
    }
    const int eq = cur - 1;
    if (eq - sq + 1 < min_qubits) return none;
    if (n_sw && eq != sw_e) return none;
    QftMatch m;
    m.sq = sq;
    m.eq = eq;
    m.inverse = true;
    m.do_swap = n_sw > 0;
    m.length = (int)(p - i);
    return m;
  }
}

}  // namespace qcsim
