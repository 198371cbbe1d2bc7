// classify.h -- host-side lowering of a (matrix, flags, q, c1, c2) gate application to a kernel shape.
//
// Step 1 reproduces the reference's dispatch (QubitRegisterCalculator.h:39-227): given the
// virtual flags, only the matrix block that the selected reference kernel reads is kept and the
// rest is taken as identity -- so a flagged gate behaves exactly as it does in QCSim, and a
// flag-less gate (AppliedGate, Compute/Uncompute replay: QubitRegister.h:488-497, 554-590) is
// taken verbatim.
// Step 2 inspects that effective matrix for structure (qubits that act as pure controls,
// diagonal, |01><->|10> pair action) and emits the cheapest in-place shape.  This is exact:
// the reference's specialised and generic kernels agree up to the sign of zero.
#pragma once

#include <complex>
#include <cstdint>
#include <cstring>

#include "../../include/qcsim_b200.h"

namespace qcsim {

typedef std::complex<double> cplx;

enum OpKind { OP_NOP = 0, OP_PAIR = 1, OP_DENSE2 = 2, OP_DENSE3 = 3, OP_DIAG = 4 };

struct Op {
  OpKind kind;
  int n_ctrl;
  int ctrl[3];   // qubits that must be 1
  int n_tgt;
  int tgt[3];    // PAIR: 1 or 2 qubits; DENSE: 2/3 qubits (matrix bit k <-> tgt[k]); DIAG: selector qubits
  // PAIR with n_tgt == 1: pair = (tgt0 = 0, tgt0 = 1).  n_tgt == 2: pair = (|tgt0=1,tgt1=0>, |tgt0=0,tgt1=1>)
  cplx m[64];    // PAIR: 2x2; DENSE: 4x4 / 8x8 row-major; DIAG: table[2^n_tgt]
};

namespace detail {

inline bool is_one(const cplx& z) { return z.real() == 1.0 && z.imag() == 0.0; }
inline bool is_zero(const cplx& z) { return z.real() == 0.0 && z.imag() == 0.0; }

// effective d x d matrix per the reference's flag dispatch
inline void effective_matrix(int nq, const double* m, int flags, cplx* E) {
  const int d = 1 << nq;
  auto in = [&](int r, int c) { return cplx(m[2 * (r * d + c)], m[2 * (r * d + c) + 1]); };
  auto ident = [&]() {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) E[r * d + c] = (r == c) ? cplx(1, 0) : cplx(0, 0);
  };
  auto full = [&]() {
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) E[r * d + c] = in(r, c);
  };
  auto block2 = [&](int a, int b, bool diag, bool anti) {  // rows/cols a,b
    ident();
    E[a * d + a] = anti ? cplx(0, 0) : in(a, a);
    E[b * d + b] = anti ? cplx(0, 0) : in(b, b);
    E[a * d + b] = diag ? cplx(0, 0) : in(a, b);
    E[b * d + a] = diag ? cplx(0, 0) : in(b, a);
  };
  const bool diag = flags & QCSIM_GATE_DIAGONAL, anti = flags & QCSIM_GATE_ANTIDIAGONAL;
  if (nq == 1) {  // QubitRegisterCalculator.h:39-81
    if (diag) block2(0, 1, true, false);
    else if (anti) block2(0, 1, false, true);
    else full();
    return;
  }
  if (nq == 2) {  // :137-156
    if (flags & (QCSIM_GATE_SWAP | QCSIM_GATE_ISWAP | QCSIM_GATE_ISWAPDAG)) {  // :230-355
      ident();
      const cplx f = (flags & QCSIM_GATE_SWAP) ? cplx(1, 0) : (flags & QCSIM_GATE_ISWAP) ? cplx(0, 1) : cplx(0, -1);
      E[1 * 4 + 1] = E[2 * 4 + 2] = cplx(0, 0);
      E[1 * 4 + 2] = E[2 * 4 + 1] = f;
    } else if (flags & QCSIM_GATE_CONTROLLED) {  // :357-417 read only rows/cols 2,3
      block2(2, 3, diag, !diag && anti);
    } else {
      full();
    }
    return;
  }
  // nq == 3, :179-199
  if (flags & QCSIM_GATE_SWAP) {  // Fredkin :181-194
    ident();
    E[5 * 8 + 5] = E[6 * 8 + 6] = cplx(0, 0);
    E[5 * 8 + 6] = E[6 * 8 + 5] = cplx(1, 0);
  } else if (flags & QCSIM_GATE_CONTROLLED) {
    if (flags & QCSIM_GATE_TWO_CONTROLS) {  // :601-609, 642-707 read only rows/cols 6,7
      block2(6, 7, diag, !diag && anti);
    } else {  // :610-638 read rows/cols 4..7
      ident();
      for (int r = 4; r < 8; ++r)
        for (int c = 4; c < 8; ++c) E[r * 8 + c] = in(r, c);
    }
  } else {
    full();
  }
}

}  // namespace detail

// qubits[k] is the register qubit of matrix bit k
inline Op classify(int nq, const double* m, int flags, uint64_t q, uint64_t c1, uint64_t c2) {
  using namespace detail;
  Op op;
  std::memset(&op, 0, sizeof(op));
  cplx E[64];
  effective_matrix(nq, m, flags, E);
  int d = 1 << nq;
  int qubits[3] = {(int)q, (int)c1, (int)c2};
  int nrem = nq;
  int rem[3] = {0, 1, 2};  // matrix bits still in play

  // peel off pure controls: matrix is identity whenever bit b of the row or of the column is 0
  // (one round suffices: removing a control cannot turn another non-control into a control
  //  because the bit-b==0 rows stay part of the test for every other bit)
  bool is_ctrl[3] = {false, false, false};
  for (int b = 0; b < nq; ++b) {
    bool ctrl = true;
    for (int r = 0; r < d && ctrl; ++r)
      for (int c = 0; c < d; ++c) {
        if (((r >> b) & 1) && ((c >> b) & 1)) continue;
        const cplx want = (r == c) ? cplx(1, 0) : cplx(0, 0);
        if (E[r * d + c] != want) {
          ctrl = false;
          break;
        }
      }
    is_ctrl[b] = ctrl;
  }
  // if every qubit looks like a control the gate is diag(1,..,1,x): keep bit 0 as the target
  // so a phase gate still has something to act on (handled by the DIAG path below)
  int nctrl_bits = 0;
  for (int b = 0; b < nq; ++b) nctrl_bits += is_ctrl[b];
  if (nctrl_bits == nq) is_ctrl[0] = false;

  int ctrl_mask_bits = 0;
  nrem = 0;
  for (int b = 0; b < nq; ++b) {
    if (is_ctrl[b]) {
      op.ctrl[op.n_ctrl++] = qubits[b];
      ctrl_mask_bits |= 1 << b;
    } else {
      rem[nrem++] = b;
    }
  }
  // reduced matrix over the remaining bits, control bits fixed to 1
  const int dr = 1 << nrem;
  cplx R[64];
  auto expand = [&](int j) {
    int idx = ctrl_mask_bits;
    for (int k = 0; k < nrem; ++k)
      if ((j >> k) & 1) idx |= 1 << rem[k];
    return idx;
  };
  for (int r = 0; r < dr; ++r)
    for (int c = 0; c < dr; ++c) R[r * dr + c] = E[expand(r) * d + expand(c)];
  for (int k = 0; k < nrem; ++k) op.tgt[k] = qubits[rem[k]];
  op.n_tgt = nrem;

  // diagonal?
  bool diagonal = true, identity = true;
  for (int r = 0; r < dr; ++r)
    for (int c = 0; c < dr; ++c) {
      if (r != c && !is_zero(R[r * dr + c])) diagonal = false;
      if (!(R[r * dr + c] == ((r == c) ? cplx(1, 0) : cplx(0, 0)))) identity = false;
    }
  if (identity) {
    op.kind = OP_NOP;
    return op;
  }
  if (diagonal) {
    op.kind = OP_DIAG;
    for (int r = 0; r < dr; ++r) op.m[r] = R[r * dr + r];
    // a selector whose 0-half of the table is all ones is really a control (phase gates)
    for (int k = 0; k < op.n_tgt;) {
      bool zero_half_is_one = true;
      const int sz = 1 << op.n_tgt;
      for (int j = 0; j < sz; ++j)
        if (!((j >> k) & 1) && !is_one(op.m[j])) zero_half_is_one = false;
      if (!zero_half_is_one) {
        ++k;
        continue;
      }
      cplx t[8];
      int o = 0;
      for (int j = 0; j < sz; ++j)
        if ((j >> k) & 1) t[o++] = op.m[j];
      for (int j = 0; j < o; ++j) op.m[j] = t[j];
      op.ctrl[op.n_ctrl++] = op.tgt[k];
      for (int j = k; j + 1 < op.n_tgt; ++j) op.tgt[j] = op.tgt[j + 1];
      --op.n_tgt;
    }
    return op;
  }
  if (nrem == 1) {
    op.kind = OP_PAIR;
    std::memcpy(op.m, R, 4 * sizeof(cplx));
    return op;
  }
  if (nrem == 2) {
    // identity on |00> and |11>, acting only inside span{|01>, |10>}: SWAP / iSWAP / Fredkin core
    bool pairlike = true;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        const bool inner = (r == 1 || r == 2) && (c == 1 || c == 2);
        if (inner) continue;
        if (!(R[r * 4 + c] == ((r == c) ? cplx(1, 0) : cplx(0, 0)))) pairlike = false;
      }
    if (pairlike) {
      op.kind = OP_PAIR;
      op.m[0] = R[1 * 4 + 1];
      op.m[1] = R[1 * 4 + 2];
      op.m[2] = R[2 * 4 + 1];
      op.m[3] = R[2 * 4 + 2];
      return op;
    }
    op.kind = OP_DENSE2;
    std::memcpy(op.m, R, 16 * sizeof(cplx));
    return op;
  }
  op.kind = OP_DENSE3;
  std::memcpy(op.m, R, 64 * sizeof(cplx));
  return op;
}

}  // namespace qcsim
