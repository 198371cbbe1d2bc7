// fusion.cu -- gate-stream planner and launcher for fused gate blocks (see fusion.h, tile_kernels.cuh).
//
// Planning is greedy and order-preserving:
//   * walk the pending ops in program order and grow a tile-qubit set T (low L qubits fixed);
//     an op is absorbed into the current pass when its non-diagonal targets fit into T and it
//     commutes with every op that was skipped before it (two ops commute when, on every qubit
//     they share, both act diagonally -- as a control or a diagonal selector);
//   * absorbed ops keep their relative order and are cut into rounds of <= 3 target bits;
//   * a pass that would not save HBM traffic over running its ops one by one is not fused.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fusion.h"
#include "planner.h"
#include "tile_kernels.cuh"

namespace qcsim {

namespace {

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : dflt;
}

}  // namespace

// Build the parameter-block descriptor for one pass and launch it.
static int launch_pass(qcsim_sv* h, const std::vector<Op>& all, const PassPlan& plan, int L) {
  const int k = (int)plan.tile.size();
  int local_of[64];
  for (int q = 0; q < 64; ++q) local_of[q] = -1;
  for (int j = 0; j < k; ++j) local_of[plan.tile[j]] = j;

  std::vector<TileRound> rounds;
  std::vector<TileOp> tops;
  std::vector<amp> pool;

  auto finish_round = [&](TileRound& rd, int nrb) {
    // pad to exactly kRoundBits distinct ascending local bits
    int bits[4];
    int n = nrb;
    for (int i = 0; i < nrb; ++i) bits[i] = rd.rb[i];
    for (int cand = 0; n < kRoundBits && cand < k; ++cand) {
      bool used = false;
      for (int i = 0; i < n; ++i) used |= (bits[i] == cand);
      if (!used) bits[n++] = cand;
    }
    std::sort(bits, bits + n);
    for (int i = 0; i < n; ++i) rd.rb[i] = bits[i];
  };

  // ---- cut into rounds: greedy on the union of target bits
  struct Pending {
    int op;
    int lt[3];
    int nlt;
  };
  std::vector<std::vector<Pending>> round_ops;
  std::vector<std::vector<int>> round_bits;
  {
    std::vector<Pending> cur;
    std::vector<int> bits;
    for (int idx : plan.ops) {
      const Op& op = all[idx];
      Pending p;
      p.op = idx;
      p.nlt = 0;
      if (op.kind != OP_DIAG)
        for (int i = 0; i < op.n_tgt; ++i) p.lt[p.nlt++] = local_of[op.tgt[i]];
      std::vector<int> u = bits;
      for (int i = 0; i < p.nlt; ++i)
        if (std::find(u.begin(), u.end(), p.lt[i]) == u.end()) u.push_back(p.lt[i]);
      if ((int)u.size() > kRoundBits) {
        round_ops.push_back(cur);
        round_bits.push_back(bits);
        cur.clear();
        bits.clear();
        for (int i = 0; i < p.nlt; ++i) bits.push_back(p.lt[i]);
      } else {
        bits = u;
      }
      cur.push_back(p);
    }
    if (!cur.empty()) {
      round_ops.push_back(cur);
      round_bits.push_back(bits);
    }
  }

  for (size_t r = 0; r < round_ops.size(); ++r) {
    TileRound rd;
    std::memset(&rd, 0, sizeof rd);
    const int nrb = (int)round_bits[r].size();
    for (int i = 0; i < nrb; ++i) rd.rb[i] = round_bits[r][i];
    finish_round(rd, nrb);
    int reg_of[16];
    for (int j = 0; j < 16; ++j) reg_of[j] = -1;
    for (int i = 0; i < kRoundBits; ++i) reg_of[rd.rb[i]] = i;
    rd.op_begin = (int)tops.size();
    for (const Pending& p : round_ops[r]) {
      const Op& op = all[p.op];
      TileOp t;
      std::memset(&t, 0, sizeof t);
      t.moff = (int)pool.size();
      // controls: round bit / other tile bit / outside the tile
      for (int i = 0; i < op.n_ctrl; ++i) {
        const int q = op.ctrl[i];
        const int lb = local_of[q];
        if (lb < 0) t.gctrl |= 1ULL << q;
        else if (reg_of[lb] >= 0) t.rctrl |= 1u << reg_of[lb];
        else t.lctrl |= 1u << lb;
      }
      auto push = [&](const cplx& z) { pool.push_back(make_amp(z.real(), z.imag())); };
      switch (op.kind) {
        case OP_PAIR:
          if (op.n_tgt == 1) {
            const cplx *m = op.m;
            const bool real = m[0].imag() == 0 && m[1].imag() == 0 && m[2].imag() == 0 && m[3].imag() == 0;
            const bool rim = m[0].imag() == 0 && m[3].imag() == 0 && m[1].real() == 0 && m[2].real() == 0;
            const bool isx = m[0] == cplx(0, 0) && m[3] == cplx(0, 0) && m[1] == cplx(1, 0) && m[2] == cplx(1, 0);
            t.kind = isx ? TK_PAIR1_X : real ? TK_PAIR1_REAL : rim ? TK_PAIR1_RIM : TK_PAIR1;
            t.r0 = reg_of[p.lt[0]];
            for (int i = 0; i < 4; ++i) push(op.m[i]);
          } else {
            const bool issw = op.m[0] == cplx(0, 0) && op.m[3] == cplx(0, 0) && op.m[1] == cplx(1, 0) && op.m[2] == cplx(1, 0);
            t.kind = issw ? TK_PAIR2_SWAP : TK_PAIR2;
            int a = reg_of[p.lt[0]], b = reg_of[p.lt[1]];
            if (a < b) {
              t.r0 = a;
              t.r1 = b;
              for (int i = 0; i < 4; ++i) push(op.m[i]);
            } else {  // swap the roles of the two amplitudes of the pair
              t.r0 = b;
              t.r1 = a;
              push(op.m[3]);
              push(op.m[2]);
              push(op.m[1]);
              push(op.m[0]);
            }
          }
          break;
        case OP_DENSE2: {
          t.kind = TK_DENSE2;
          int a = reg_of[p.lt[0]], b = reg_of[p.lt[1]];
          const bool flip = a > b;  // matrix bit0 <-> r0, bit1 <-> r1 with r0 < r1
          t.r0 = flip ? b : a;
          t.r1 = flip ? a : b;
          auto perm = [&](int i) { return flip ? (((i & 1) << 1) | ((i >> 1) & 1)) : i; };
          for (int r2 = 0; r2 < 4; ++r2)
            for (int c = 0; c < 4; ++c) push(op.m[perm(r2) * 4 + perm(c)]);
          break;
        }
        case OP_DENSE3: {
          t.kind = TK_DENSE3;
          int rr[3] = {reg_of[p.lt[0]], reg_of[p.lt[1]], reg_of[p.lt[2]]};
          // register index x has bit rr[j] <-> matrix bit j
          auto perm = [&](int x) {
            int mi = 0;
            for (int j = 0; j < 3; ++j)
              if ((x >> rr[j]) & 1) mi |= 1 << j;
            return mi;
          };
          for (int r2 = 0; r2 < 8; ++r2)
            for (int c = 0; c < 8; ++c) push(op.m[perm(r2) * 8 + perm(c)]);
          break;
        }
        case OP_DIAG: {
          t.kind = op.n_tgt == 0 ? TK_PHASE : TK_DIAG;
          t.nsel = op.n_tgt;
          for (int i = 0; i < op.n_tgt; ++i) {
            const int q = op.tgt[i];
            const int lb = local_of[q];
            if (lb < 0) {
              t.sel_src[i] = 2;
              t.sel_pos[i] = q;
            } else if (reg_of[lb] >= 0) {
              t.sel_src[i] = 0;
              t.sel_pos[i] = reg_of[lb];
            } else {
              t.sel_src[i] = 1;
              t.sel_pos[i] = lb;
            }
          }
          for (int i = 0; i < 8; ++i) push(i < (1 << op.n_tgt) ? op.m[i] : cplx(1, 0));
          break;
        }
        default: break;
      }
      tops.push_back(t);
    }
    rd.op_end = (int)tops.size();
    rounds.push_back(rd);
  }

  // ---- descriptor -> kernel parameter block
  if (rounds.size() > (size_t)kMaxTileRounds || tops.size() > (size_t)kMaxTileOps || pool.size() > (size_t)kMaxTilePool)
    return fail(QCSIM_ERR_BAD_ARG, "internal: pass too large (%zu rounds, %zu ops, %zu pool)", rounds.size(), tops.size(), pool.size());
  static thread_local TilePassArgs A;  // 12 KiB: keep it off the stack; the launch copies it
  A.k = k;
  A.n_rounds = (int)rounds.size();
  A.low_identity = L;
  A.n_tiles = 1ULL << (h->n_local - k);
  for (int j = 0; j < kMaxTileBits; ++j) A.tpos[j] = j < k ? plan.tile[j] : 0;
  for (size_t r = 0; r < rounds.size(); ++r) {
    A.rounds[r].x = (unsigned)(rounds[r].rb[0] | (rounds[r].rb[1] << 8) | (rounds[r].rb[2] << 16));
    A.rounds[r].y = (unsigned)(rounds[r].op_begin | (rounds[r].op_end << 16));
  }
  for (size_t o = 0; o < tops.size(); ++o) pack_tile_op(tops[o], &A.ops[2 * o]);
  for (size_t i = 0; i < pool.size(); ++i) A.pool[i] = pool[i];
  const size_t smem = (size_t)sizeof(amp) << k;
  static const int variant = env_int("QCSIM_TILE_VARIANT", 0);  // 0: NI=1/128 regs, 1: NI=1/80 regs, 2: NI=2/128 regs
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(k_tile_pass<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  const int by_smem = std::max(1, (int)((220 * 1024) / (smem + 1024)));
  const int per_sm = std::min(by_smem, variant == 1 ? 3 : 2);
  const uint64_t grid = std::min<uint64_t>(A.n_tiles, (uint64_t)kNumSMs * per_sm);
  if (variant == 2 && k == 12) k_tile_pass<2, 2><<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
  else if (variant == 1) k_tile_pass<1, 3><<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
  else k_tile_pass<1, 2><<<(unsigned)grid, kTileThreads, smem, h->stream>>>(h->psi, A);
  CUDA_TRY(cudaGetLastError());
  h->stats.kernel_launches += 1;
  h->stats.state_passes += 1;
  h->stats.bytes_moved += 32ULL * h->dim_local;
  return QCSIM_OK;
}

int fusion_execute(qcsim_sv* h, const std::vector<Op>& ops_in) {
  if (h->world > 1) return dist_execute(h, ops_in);
  return fusion_execute_local(h, ops_in);
}

// All qubit indices in `ops` are physical bit positions of the local slice.
int fusion_execute_local(qcsim_sv* h, const std::vector<Op>& ops) {
  const int nl = h->n_local;
  const int N = (int)ops.size();
  static const int K_env = env_int("QCSIM_TILE_BITS", kMaxTileBits);
  static const int L_env = env_int("QCSIM_TILE_LOW", 4);
  static const int no_fuse = env_int("QCSIM_NO_FUSION", 0);
  const int K = std::max(kRoundBits, std::min({K_env, kMaxTileBits, nl}));
  const int L = std::max(1, std::min(L_env, K - kRoundBits));
  if (no_fuse || nl < 6 || N < 2) {
    for (const Op& op : ops) QCSIM_TRY(engine_launch_local(h, op));
    return QCSIM_OK;
  }

  const std::vector<PlanStep> steps = plan_passes(ops, nl, K, L, kMaxTileOps, kMaxTilePool);
  static const int debug = env_int("QCSIM_DEBUG_PLAN", 0);
  if (debug) {
    int nf = 0, absorbed = 0;
    for (const PlanStep& st : steps)
      if (st.fused) {
        ++nf;
        absorbed += (int)st.pass.ops.size();
      }
    std::fprintf(stderr, "[qcsim plan] %d ops -> %zu steps (%d fused passes holding %d ops), K=%d L=%d\n", N, steps.size(), nf,
                 absorbed, K, L);
  }
  for (const PlanStep& st : steps) {
    if (!st.fused) {
      QCSIM_TRY(engine_launch_local(h, ops[st.pass.ops[0]]));
      continue;
    }
    int Lrun = 0;  // the low run of identity-mapped tile bits may be longer than L
    while (Lrun < (int)st.pass.tile.size() && st.pass.tile[Lrun] == Lrun) ++Lrun;
    QCSIM_TRY(launch_pass(h, ops, st.pass, Lrun));
  }
  return QCSIM_OK;
}

int fusion_reserve(qcsim_sv*) { return QCSIM_OK; }
void fusion_release(qcsim_sv*) {}

}  // namespace qcsim
