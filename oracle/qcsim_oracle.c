/* oracle/qcsim_oracle.c -- CPU restatement ("port") of QCSim's statevector hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under qcsim_b200/ may include, link or load this file;
 * it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg have a checker
 * that can be built anywhere gcc exists (the GPU box has no /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port bit-for-bit (up to the sign of
 * zero) against the reference's own headers compiled into oracle/_ref/libqcsim_ref_sse2.so for
 * every gate class, every kernel branch, QFT/IQFT and all measurement entry points, and
 * against the committed fixtures in tests/golden/ (generated from that compiled reference by
 * tests/golden/make_golden.py).
 *
 * All file:line citations are relative to /root/reference/QCSim/.
 *
 * Conventions (QubitRegisterCalculator.h:427,749): amplitude index bit q is qubit q; gate
 * matrix row/col bit0 = `qubit`, bit1 = `controllingQubit1`, bit2 = `controllingQubit2`.
 * Matrices cross this interface row-major as (re, im) pairs.
 * Arithmetic follows the reference's -msse2 build: complex product (ac-bd, ad+bc) with four
 * separately rounded multiplies, sums left to right, no FMA (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

typedef struct {
  int n;          /* qubits */
  uint64_t dim;   /* 2^n */
  cplx* psi;      /* registerStorage  (QubitRegister.h:718) */
  cplx* scratch;  /* resultsStorage   (QubitRegister.h:719) */
} orc_reg;

/* the reference's virtual flags (SimpleGates.h:27-60), packed the same way as ref_gate_flags */
enum {
  F_CONTROLLED = 1, F_TWO_CONTROLS = 2, F_DIAGONAL = 4, F_ANTIDIAGONAL = 8,
  F_SWAP = 16, F_ISWAP = 32, F_ISWAPDAG = 64
};

static inline cplx cmul(cplx a, cplx b) {
  cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
  return r;
}
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static inline double cnorm(cplx a) { return a.re * a.re + a.im * a.im; } /* std::norm */
static inline cplx M(const double* m, int d, int r, int c) {
  cplx z = { m[2 * (r * d + c)], m[2 * (r * d + c) + 1] };
  return z;
}
static void swap_buffers(orc_reg* r) { cplx* t = r->psi; r->psi = r->scratch; r->scratch = t; }

/* ---- lifecycle ------------------------------------------------------------------------- */

orc_reg* orc_create(int n) { /* QubitRegister.h:17-37: zero vector, a[0] = 1 */
  orc_reg* r = (orc_reg*)malloc(sizeof(orc_reg));
  r->n = n;
  r->dim = 1ULL << n;
  r->psi = (cplx*)calloc(r->dim, sizeof(cplx));
  r->scratch = (cplx*)calloc(r->dim, sizeof(cplx));
  r->psi[0].re = 1.0;
  return r;
}
void orc_destroy(orc_reg* r) { free(r->psi); free(r->scratch); free(r); }
void orc_get_state(const orc_reg* r, double* out) { memcpy(out, r->psi, r->dim * sizeof(cplx)); }
void orc_set_state(orc_reg* r, const double* in) { memcpy(r->psi, in, r->dim * sizeof(cplx)); }
void orc_set_basis_state(orc_reg* r, uint64_t s) { /* QubitRegister.h:74-80 */
  if (s >= r->dim) return;
  memset(r->psi, 0, r->dim * sizeof(cplx));
  r->psi[s].re = 1.0;
}

/* ---- one-qubit gates: QubitRegisterCalculator.h:39-135 ---------------------------------- */

static void one_qubit(orc_reg* r, const double* m, int flags, uint64_t qb) {
  const int64_t dim = (int64_t)r->dim;
  cplx* a = r->psi;
  if (flags & F_DIAGONAL) { /* :41-52, in place, a[s] *= s&bit ? m11 : m00 */
    const cplx v0 = M(m, 2, 0, 0), v1 = M(m, 2, 1, 1);
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) a[s] = cmul(a[s], (s & qb) ? v1 : v0);
    return;
  }
  cplx* o = r->scratch;
  if (flags & F_ANTIDIAGONAL) { /* :57-67 */
    const cplx v10 = M(m, 2, 1, 0), v01 = M(m, 2, 0, 1);
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) o[s] = (s & qb) ? cmul(v10, a[s & ~qb]) : cmul(v01, a[s | qb]);
  } else { /* :68-79 generic */
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) {
      const int row = (s & qb) ? 1 : 0;
      o[s] = cadd(cmul(M(m, 2, row, 0), a[s & ~qb]), cmul(M(m, 2, row, 1), a[s | qb]));
    }
  }
  swap_buffers(r);
}

/* ---- two-qubit gates: dispatcher :137-177, kernels :230-597 ----------------------------- */

static void two_qubit(orc_reg* r, const double* m, int flags, uint64_t qb, uint64_t cb) {
  const int64_t dim = (int64_t)r->dim;
  cplx* a = r->psi;
  const uint64_t both = qb | cb;
  if (flags & (F_SWAP | F_ISWAP | F_ISWAPDAG)) { /* :230-355, in place on (q=0,c=1) <-> (q=1,c=0) */
    const int kind = (flags & F_SWAP) ? 0 : (flags & F_ISWAP) ? 1 : 2;
    const cplx iv = { 0.0, kind == 1 ? 1.0 : -1.0 };
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) {
      if ((s & qb) != 0 || (s & cb) == 0) continue;
      const uint64_t t = s ^ both;
      cplx x = a[s];
      a[s] = a[t];
      a[t] = x;
      if (kind) { a[s] = cmul(a[s], iv); a[t] = cmul(a[t], iv); }
    }
    return;
  }
  if (flags & F_CONTROLLED) {
    if (flags & F_DIAGONAL) { /* :357-371 in place, only rows 2,3 are read */
      const cplx v2 = M(m, 4, 2, 2), v3 = M(m, 4, 3, 3);
#pragma omp parallel for
      for (int64_t s = 0; s < dim; ++s)
        if (s & cb) a[s] = cmul(a[s], (s & qb) ? v3 : v2);
      return;
    }
    cplx* o = r->scratch;
    if (flags & F_ANTIDIAGONAL) { /* :373-393 */
      const cplx v32 = M(m, 4, 3, 2), v23 = M(m, 4, 2, 3);
#pragma omp parallel for
      for (int64_t s = 0; s < dim; ++s) {
        if (!(s & cb)) { o[s] = a[s]; continue; }
        o[s] = (s & qb) ? cmul(v32, a[s & ~qb]) : cmul(v23, a[s | qb]);
      }
    } else { /* :395-417 generic controlled: lower-right 2x2 block */
#pragma omp parallel for
      for (int64_t s = 0; s < dim; ++s) {
        if (!(s & cb)) { o[s] = a[s]; continue; }
        const int row = 2 | ((s & qb) ? 1 : 0);
        o[s] = cadd(cmul(M(m, 4, row, 2), a[(s & ~qb) | cb]), cmul(M(m, 4, row, 3), a[s | both]));
      }
    }
    swap_buffers(r);
    return;
  }
  { /* :419-465 dense 4x4 */
    cplx* o = r->scratch;
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) {
      const int row = ((s & cb) ? 2 : 0) | ((s & qb) ? 1 : 0);
      const uint64_t base = s & ~both;
      cplx acc = cmul(M(m, 4, row, 0), a[base]);
      acc = cadd(acc, cmul(M(m, 4, row, 1), a[base | qb]));
      acc = cadd(acc, cmul(M(m, 4, row, 2), a[base | cb]));
      acc = cadd(acc, cmul(M(m, 4, row, 3), a[base | both]));
      o[s] = acc;
    }
    swap_buffers(r);
  }
}

/* ---- three-qubit gates: dispatcher :179-227, kernels :599-939 --------------------------- */
/* qb = 1<<qubit (matrix bit0), q2 = 1<<controllingQubit1 (bit1), cb = 1<<controllingQubit2 (bit2) */

static void three_qubit(orc_reg* r, const double* m, int flags, uint64_t qb, uint64_t q2, uint64_t cb) {
  const int64_t dim = (int64_t)r->dim;
  cplx* a = r->psi;
  const uint64_t all = qb | q2 | cb;
  if (flags & F_SWAP) { /* Fredkin :181-194: swap q<->q2 where cb set, acting on (q2=1,q=0) */
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) {
      if ((s & cb) == 0 || (s & q2) == 0 || (s & qb) != 0) continue;
      const uint64_t t = s ^ (qb | q2);
      cplx x = a[s];
      a[s] = a[t];
      a[t] = x;
    }
    return;
  }
  if (flags & F_CONTROLLED) {
    if (flags & F_TWO_CONTROLS) { /* :601-609 */
      if (flags & F_DIAGONAL) { /* :642-656 in place, entries 66 / 77 */
        const cplx v6 = M(m, 8, 6, 6), v7 = M(m, 8, 7, 7);
#pragma omp parallel for
        for (int64_t s = 0; s < dim; ++s)
          if ((s & cb) && (s & q2)) a[s] = cmul(a[s], (s & qb) ? v7 : v6);
        return;
      }
      cplx* o = r->scratch;
      if (flags & F_ANTIDIAGONAL) { /* :658-680 (Toffoli) */
        const cplx v76 = M(m, 8, 7, 6), v67 = M(m, 8, 6, 7);
#pragma omp parallel for
        for (int64_t s = 0; s < dim; ++s) {
          if (!(s & cb) || !(s & q2)) { o[s] = a[s]; continue; }
          o[s] = (s & qb) ? cmul(v76, a[s & ~qb]) : cmul(v67, a[s | qb]);
        }
      } else { /* :682-707 generic CC: 2x2 block rows/cols 6,7 */
#pragma omp parallel for
        for (int64_t s = 0; s < dim; ++s) {
          if (!(s & cb) || !(s & q2)) { o[s] = a[s]; continue; }
          const int row = 6 | ((s & qb) ? 1 : 0);
          o[s] = cadd(cmul(M(m, 8, row, 6), a[(s & ~qb) | cb | q2]), cmul(M(m, 8, row, 7), a[s | all]));
        }
      }
      swap_buffers(r);
      return;
    }
    { /* :610-638 one control (cb), 4x4 block rows/cols 4..7 on (q2, q) */
      cplx* o = r->scratch;
#pragma omp parallel for
      for (int64_t s = 0; s < dim; ++s) {
        if (!(s & cb)) { o[s] = a[s]; continue; }
        const int row = 4 | ((s & q2) ? 2 : 0) | ((s & qb) ? 1 : 0);
        const uint64_t base = (s & ~(qb | q2)) | cb;
        cplx acc = cmul(M(m, 8, row, 4), a[base]);
        acc = cadd(acc, cmul(M(m, 8, row, 5), a[base | qb]));
        acc = cadd(acc, cmul(M(m, 8, row, 6), a[base | q2]));
        acc = cadd(acc, cmul(M(m, 8, row, 7), a[base | qb | q2]));
        o[s] = acc;
      }
      swap_buffers(r);
      return;
    }
  }
  { /* :710-762 dense 8x8 */
    cplx* o = r->scratch;
#pragma omp parallel for
    for (int64_t s = 0; s < dim; ++s) {
      const int row = ((s & cb) ? 4 : 0) | ((s & q2) ? 2 : 0) | ((s & qb) ? 1 : 0);
      const uint64_t base = s & ~all;
      cplx acc = cmul(M(m, 8, row, 0), a[base]);
      for (int c = 1; c < 8; ++c) {
        const uint64_t idx = base | ((c & 1) ? qb : 0) | ((c & 2) ? q2 : 0) | ((c & 4) ? cb : 0);
        acc = cadd(acc, cmul(M(m, 8, row, c), a[idx]));
      }
      o[s] = acc;
    }
    swap_buffers(r);
  }
}

/* QubitRegister::ApplyGate (QubitRegister.h:434-486) incl. CheckQubits (:677-690).
 * returns 0, or -1 = qubit too high, -2 = controlling qubit too high, -3 = duplicate qubits */
int orc_apply(orc_reg* r, int nq, const double* m, int flags, uint64_t q, uint64_t c1, uint64_t c2) {
  const uint64_t n = (uint64_t)r->n;
  if (n <= q) return -1;
  if (nq == 2) {
    if (n <= c1) return -2;
    if (q == c1) return -3;
  } else if (nq == 3) {
    if (n <= c1 || n <= c2) return -2;
    if (q == c1 || q == c2 || c1 == c2) return -3;
  }
  if (nq == 1) one_qubit(r, m, flags, 1ULL << q);
  else if (nq == 2) two_qubit(r, m, flags, 1ULL << q, 1ULL << c1);
  else if (nq == 3) three_qubit(r, m, flags, 1ULL << q, 1ULL << c1, 1ULL << c2);
  else return -4;
  return 0;
}

/* ---- reductions and measurement --------------------------------------------------------- */

double orc_norm2(const orc_reg* r) {
  double s = 0;
  for (uint64_t i = 0; i < r->dim; ++i) s += cnorm(r->psi[i]);
  return s;
}

/* GetQubitProbability (QubitRegisterCalculator.h:1088-1101) */
double orc_qubit_probability(const orc_reg* r, uint64_t q) {
  const uint64_t bit = 1ULL << q;
  double acc = 0;
  for (uint64_t s = bit; s < r->dim; ++s)
    if (s & bit) acc += cnorm(r->psi[s]);
  return acc;
}

/* the sequential scan shared by every Measure* entry point: first i with prob <= running sum */
static uint64_t scan_pick(const orc_reg* r, double prob, uint64_t fallback) {
  double acc = 0;
  for (uint64_t i = 0; i < r->dim; ++i) {
    acc += cnorm(r->psi[i]);
    if (prob <= acc) return i;
  }
  return fallback;
}

/* MeasureAll (QubitRegister.h:169-195): fallback = last state; collapse = setToBasisState */
uint64_t orc_measure_all(orc_reg* r, double prob) {
  const uint64_t s = scan_pick(r, prob, r->dim - 1);
  orc_set_basis_state(r, s);
  return s;
}
/* MeasureNoCollapse() (QubitRegister.h:619-642): fallback = 0 */
uint64_t orc_measure_all_nocollapse(const orc_reg* r, double prob) { return scan_pick(r, prob, 0); }

/* Measure(first,last) / MeasureQubit (QubitRegisterCalculator.h:948-998, 1124-1169) */
uint64_t orc_measure(orc_reg* r, uint64_t first, uint64_t last, double prob) {
  const uint64_t low = (1ULL << first) - 1;
  const uint64_t mask = (1ULL << (last + 1)) - 1 - low;
  const uint64_t picked = scan_pick(r, prob, 0) & mask;
  double acc = 0;
  for (uint64_t s = picked; s < r->dim; ++s)
    if ((s & mask) == picked) acc += cnorm(r->psi[s]);
  const double scale = 1. / sqrt(acc);
  for (uint64_t s = 0; s < r->dim; ++s) {
    const double f = ((s & mask) == picked) ? scale : 0.0;
    r->psi[s].re *= f;
    r->psi[s].im *= f;
  }
  return picked >> first;
}
/* MeasureNoCollapse(first,last) (QubitRegisterCalculator.h:1061-1086, 1227-1254) */
uint64_t orc_measure_nocollapse(const orc_reg* r, uint64_t first, uint64_t last, double prob) {
  const uint64_t low = (1ULL << first) - 1;
  const uint64_t mask = (1ULL << (last + 1)) - 1 - low;
  return (scan_pick(r, prob, 0) & mask) >> first;
}

/* ---- QFT / IQFT (QuantumFourierTransform.h:35-87) + QubitsSwapper::Swap (QubitsSwapper.h:23-34) */

static void hadamard_matrix(double* m) { /* SimpleGates.h:588-596 */
  const double v = 1. / sqrt(2.);
  const double h[8] = { v, 0, v, 0, v, 0, -v, 0 };
  memcpy(m, h, sizeof h);
}
static void cphase_matrix(double* m, double theta) { /* QuantumGate.h:251-269: identity, m33 = polar(1, theta) */
  memset(m, 0, 32 * sizeof(double));
  m[0] = m[10] = m[20] = 1.0;
  m[30] = cos(theta);
  m[31] = sin(theta);
}
static void swap_matrix(double* m) { /* QuantumGate.h:10-28 */
  memset(m, 0, 32 * sizeof(double));
  m[0] = 1.0; m[2 * (1 * 4 + 2)] = 1.0; m[2 * (2 * 4 + 1)] = 1.0; m[30] = 1.0;
}

int orc_qft(orc_reg* r, uint64_t sq_, uint64_t eq_, int do_swap, int inverse) {
  /* sub-register clamp: QuantumAlgorithm.h, QuantumSubAlgorithmOnSubregister ctor */
  const uint64_t nm1 = (uint64_t)r->n - 1;
  const int sq = (int)sq_;
  const int eq = (int)(sq_ > (eq_ < nm1 ? eq_ : nm1) ? sq_ : (eq_ < nm1 ? eq_ : nm1));
  double h[8], cp[32], sw[32];
  hadamard_matrix(h);
  swap_matrix(sw);
  const double pi_2 = 1.57079632679489661923; /* M_PI_2 */
  if (!inverse) {
    orc_apply(r, 1, h, 0, (uint64_t)eq, 0, 0);
    for (int cur = eq; cur > sq; --cur) {
      double phase = pi_2;
      for (int ctrl = cur - 1; ctrl >= sq; --ctrl) {
        cphase_matrix(cp, phase);
        orc_apply(r, 2, cp, F_CONTROLLED | F_DIAGONAL, (uint64_t)cur, (uint64_t)ctrl, 0);
        phase *= 0.5;
      }
      orc_apply(r, 1, h, 0, (uint64_t)(cur - 1), 0, 0);
    }
  }
  if (do_swap) {
    /* after the ladder for QFT, before it for IQFT */
    uint64_t s = (uint64_t)sq, e = (uint64_t)eq;
    while (s < e) { orc_apply(r, 2, sw, F_SWAP, s, e, 0); ++s; --e; }
  }
  if (inverse) {
    for (int cur = sq + 1; cur <= eq; ++cur) {
      orc_apply(r, 1, h, 0, (uint64_t)(cur - 1), 0, 0);
      double phase = -pi_2;
      for (int ctrl = cur - 1; ctrl >= sq; --ctrl) {
        cphase_matrix(cp, phase);
        orc_apply(r, 2, cp, F_CONTROLLED | F_DIAGONAL, (uint64_t)cur, (uint64_t)ctrl, 0);
        phase *= 0.5;
      }
    }
    orc_apply(r, 1, h, 0, (uint64_t)eq, 0, 0);
  }
  return 0;
}
