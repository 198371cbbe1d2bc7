// oracle/ref_driver.cpp -- C-callable driver around the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY (never linked or loaded by the product path).
// This translation unit #includes QCSim's own headers straight from /root/reference/QCSim
// (QubitRegister.h, QubitRegisterCalculator.h, SimpleGates.h, QuantumGate.h,
// QuantumFourierTransform.h, QubitsSwapper.h, QuantumAlgorithm.h, GroverAlgorithm.h,
// DraperAdder.h) and exposes them through a flat C interface for ctypes.  Eigen is replaced
// by oracle/eigen_shim (container API only).  Built by oracle/Makefile into
// oracle/_ref/libqcsim_ref_{sse2,avx2}.so; no reference source is copied into this repo.
//
// Every function below drives reference code; none re-implements it.

#include <map>
#include <climits>
#include <memory>
#include <cstring>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "QubitRegister.h"
#include "QuantumFourierTransform.h"
#include "GroverAlgorithm.h"
#include "DraperAdder.h"
#include "NControlledNotWithAncilla.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Vec = Eigen::VectorXcd;
using Mat = Eigen::MatrixXcd;
using Gate = QC::Gates::QuantumGateWithOp<Mat>;

// Subclass only to reach the protected RNG / storage members; no behaviour is overridden.
class RefRegister : public QC::QubitRegister<Vec, Mat> {
public:
  using Base = QC::QubitRegister<Vec, Mat>;
  explicit RefRegister(size_t n) : Base(n, 12345u) {}

  // Make the next `1. - uniformZeroOne(rng)` evaluate to exactly `prob`
  // (QubitRegister.h:171,210,621,707).  Exact whenever prob is a multiple of 2^-53 in (0,1],
  // which is what every real draw is; the test harness only injects such values.
  void injectDraw(double prob) {
    const double a = 1. - prob;
    uniformZeroOne = std::uniform_real_distribution<double>(a, a);
  }
  void reseed(uint64_t seed) {
    rng.seed(seed);
    uniformZeroOne = std::uniform_real_distribution<double>(0, 1);
  }
  Vec& storage() { return registerStorage; }
  double nextDraw() { return 1. - uniformZeroOne(rng); }  // the expression at QubitRegister.h:171
  size_t measureNoCollapseRange(size_t a, size_t b) { return Base::MeasureNoCollapse(a, b); }
};

// Expose the protected circuit bodies of two reference algorithms (no behaviour change).
class GroverProbe : public Grover::GroverAlgorithmWithGatesOracle<Vec, Mat> {
public:
  explicit GroverProbe(size_t n) : Grover::GroverAlgorithmWithGatesOracle<Vec, Mat>(n, 12345u) {}
  void run() { ExecuteWithoutMeasurement(); }
};
class DraperProbe : public Adders::DraperAdder<Vec, Mat> {
public:
  explicit DraperProbe(size_t n) : Adders::DraperAdder<Vec, Mat>(n, 12345u) {}
  void run() { ExecuteWithoutMeasurement(); }
};

thread_local std::string g_err;

std::unique_ptr<Gate> makeGate(int id, const double* p) {
  using namespace QC::Gates;
  const double p0 = p ? p[0] : 0, p1 = p ? p[1] : 0, p2 = p ? p[2] : 0, p3 = p ? p[3] : 0;
  switch (id) {
    case 0: return std::make_unique<HadamardGate<Mat>>();
    case 1: return std::make_unique<HyGate<Mat>>();
    case 2: return std::make_unique<SGate<Mat>>();
    case 3: return std::make_unique<SDGGate<Mat>>();
    case 4: return std::make_unique<TGate<Mat>>();
    case 5: return std::make_unique<TDGGate<Mat>>();
    case 6: return std::make_unique<PhaseShiftGate<Mat>>(p0);
    case 7: return std::make_unique<PauliXGate<Mat>>();
    case 8: return std::make_unique<PauliYGate<Mat>>();
    case 9: return std::make_unique<PauliZGate<Mat>>();
    case 10: return std::make_unique<SquareRootNOTGate<Mat>>();
    case 11: return std::make_unique<SquareRootNOTDagGate<Mat>>();
    case 12: return std::make_unique<SplitterGate<Mat>>();
    case 13: return std::make_unique<RxGate<Mat>>(p0);
    case 14: return std::make_unique<RyGate<Mat>>(p0);
    case 15: return std::make_unique<RzGate<Mat>>(p0);
    case 16: return std::make_unique<UGate<Mat>>(p0, p1, p2, p3);
    case 20: return std::make_unique<SwapGate<Mat>>();
    case 21: return std::make_unique<iSwapGate<Mat>>();
    case 22: return std::make_unique<iSwapDagGate<Mat>>();
    case 23: return std::make_unique<DecrementGate<Mat>>();
    case 24: return std::make_unique<CNOTGate<Mat>>();
    case 25: return std::make_unique<ControlledYGate<Mat>>();
    case 26: return std::make_unique<ControlledZGate<Mat>>();
    case 27: return std::make_unique<ControlledHadamardGate<Mat>>();
    case 28: return std::make_unique<ControlledSquareRootNOTGate<Mat>>();
    case 29: return std::make_unique<ControlledSquareRootNOTDagGate<Mat>>();
    case 30: return std::make_unique<ControlledPhaseGate<Mat>>();
    case 31: return std::make_unique<ControlledPhaseShiftGate<Mat>>(p0);
    case 32: return std::make_unique<ControlledUGate<Mat>>(p0, p1, p2, p3);
    case 33: return std::make_unique<ControlledRxGate<Mat>>(p0);
    case 34: return std::make_unique<ControlledRyGate<Mat>>(p0);
    case 35: return std::make_unique<ControlledRzGate<Mat>>(p0);
    case 40: return std::make_unique<ToffoliGate<Mat>>();
    case 41: return std::make_unique<FredkinGate<Mat>>();
    case 42: return std::make_unique<CCZGate<Mat>>();
    default: return nullptr;
  }
}

Mat fromRowMajor(int nq, const double* m) {
  const int d = 1 << nq;
  Mat out(d, d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) out(i, j) = std::complex<double>(m[2 * (i * d + j)], m[2 * (i * d + j) + 1]);
  return out;
}

template <class F> int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::invalid_argument& e) {
    g_err = e.what();
    return -1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -2;
  }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void* ref_create(int n_qubits) { return new RefRegister(static_cast<size_t>(n_qubits)); }
void ref_destroy(void* h) { delete static_cast<RefRegister*>(h); }
void ref_set_multithreading(void* h, int on) { static_cast<RefRegister*>(h)->SetMultithreading(on != 0); }

// 2x2 / 4x4 / 8x8 matrix of a reference gate class, row-major (re, im) pairs.
int ref_gate_matrix(int gate_id, const double* params, double* out, int* nq_out) {
  auto g = makeGate(gate_id, params);
  if (!g) return -1;
  const Mat& m = g->getRawOperatorMatrix();
  const int d = static_cast<int>(m.rows());
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      out[2 * (i * d + j)] = m(i, j).real();
      out[2 * (i * d + j) + 1] = m(i, j).imag();
    }
  *nq_out = static_cast<int>(g->getQubitsNumber());
  return 0;
}

// bit0 controlled, bit1 control-qubit(1) i.e. two controls, bit2 diagonal, bit3 antidiagonal,
// bit4 swap, bit5 iswap, bit6 iswapdag: the reference's virtual flags (SimpleGates.h:12-61).
int ref_gate_flags(int gate_id) {
  auto g = makeGate(gate_id, nullptr);
  if (!g) return -1;
  return (g->isControlled() ? 1 : 0) | (g->isControlQubit(1) ? 2 : 0) | (g->isDiagonal() ? 4 : 0) |
         (g->isAntidiagonal() ? 8 : 0) | (g->isSwapGate() ? 16 : 0) | (g->IsISwapGate() ? 32 : 0) |
         (g->IsISwapDagGate() ? 64 : 0);
}

// reg.ApplyGate(<reference gate class>, q, c1, c2)  -- flagged dispatch (QubitRegister.h:434)
int ref_apply_named(void* h, int gate_id, const double* params, uint64_t q, uint64_t c1, uint64_t c2) {
  auto g = makeGate(gate_id, params);
  if (!g) {
    g_err = "unknown gate id";
    return -3;
  }
  return guarded([&] { static_cast<RefRegister*>(h)->ApplyGate(*g, q, c1, c2); });
}

// reg.ApplyGate(AppliedGate(matrix, q, c1, c2)) -- flag-less generic kernels (QubitRegister.h:488)
int ref_apply_matrix(void* h, int nq, const double* m, uint64_t q, uint64_t c1, uint64_t c2) {
  QC::Gates::AppliedGate<Mat> g(fromRowMajor(nq, m), q, c1, c2);
  return guarded([&] { static_cast<RefRegister*>(h)->ApplyGate(g); });
}

// Second, independent reference path: dense 2^n x 2^n operator (SimpleGates.h:211-232 etc.)
int ref_apply_named_dense(void* h, int gate_id, const double* params, uint64_t q, uint64_t c1, uint64_t c2) {
  auto g = makeGate(gate_id, params);
  if (!g) return -3;
  RefRegister* r = static_cast<RefRegister*>(h);
  return guarded([&] { r->ApplyOperatorMatrix(g->getOperatorMatrix(r->getNrQubits(), q, c1, c2)); });
}

void ref_get_state(void* h, double* out) {
  const Vec& v = static_cast<RefRegister*>(h)->getRegisterStorage();
  std::memcpy(out, v.data(), sizeof(double) * 2 * static_cast<size_t>(v.size()));
}

// setRegisterStorageFastNoNormalize semantics: raw copy, no normalisation (QubitRegister.h:521)
void ref_set_state(void* h, const double* in) {
  Vec& v = static_cast<RefRegister*>(h)->storage();
  std::memcpy(v.data(), in, sizeof(double) * 2 * static_cast<size_t>(v.size()));
}

void ref_set_basis_state(void* h, uint64_t s) { static_cast<RefRegister*>(h)->setToBasisState(s); }
void ref_set_equal_superposition(void* h) { static_cast<RefRegister*>(h)->setToEqualSuperposition(); }
void ref_set_cat_state(void* h) { static_cast<RefRegister*>(h)->setToCatState(); }
void ref_normalize(void* h) { static_cast<RefRegister*>(h)->Normalize(); }

double ref_norm2(void* h) { return static_cast<RefRegister*>(h)->getRegisterStorage().squaredNorm(); }
double ref_qubit_probability(void* h, uint64_t q) { return static_cast<RefRegister*>(h)->GetQubitProbability(q); }

uint64_t ref_measure_all(void* h, double prob) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->injectDraw(prob);
  return r->MeasureAll();
}
uint64_t ref_measure(void* h, uint64_t first, uint64_t last, double prob) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->injectDraw(prob);
  return r->Measure(first, last);
}
uint64_t ref_measure_all_nocollapse(void* h, double prob) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->injectDraw(prob);
  return r->MeasureNoCollapse();
}
uint64_t ref_measure_nocollapse(void* h, uint64_t first, uint64_t last, double prob) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->injectDraw(prob);
  return r->measureNoCollapseRange(first, last);
}

// RepeatedMeasure / RepeatedMeasureUnordered, whole register or [first, last] (QubitRegister.h:227-429), after
// rng.seed(seed).  variant: 0 ordered, 1 unordered; first > last: the whole-register overloads.  Writes up to `cap`
// (outcome, count) pairs sorted by outcome; returns the number of distinct outcomes.
int ref_repeated_measure(void* h, uint64_t seed, uint64_t nr_times, int variant, uint64_t first, uint64_t last, uint64_t* keys, uint64_t* counts,
                         int cap) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->reseed(seed);
  std::map<size_t, size_t> out;
  if (first > last) {
    if (variant == 0) out = r->RepeatedMeasure(nr_times);
    else for (const auto& kv : r->RepeatedMeasureUnordered(nr_times)) out[kv.first] = kv.second;
  } else {
    if (variant == 0) out = r->RepeatedMeasure(first, last, nr_times);
    else for (const auto& kv : r->RepeatedMeasureUnordered(first, last, nr_times)) out[kv.first] = kv.second;
  }
  int i = 0;
  for (const auto& kv : out) {
    if (i < cap) {
      keys[i] = kv.first;
      counts[i] = kv.second;
    }
    ++i;
  }
  return i;
}

// stateFidelity (QubitRegister.h:527-534) against a caller-supplied state
double ref_state_fidelity(void* h, const double* state, uint64_t dim) {
  RefRegister* r = static_cast<RefRegister*>(h);
  Vec v(dim);
  std::memcpy(v.data(), state, sizeof(double) * 2 * dim);
  return r->stateFidelity(v);
}

// ExpectationValue (QubitRegister.h:646-660) of a product of gates given as (nq, row-major matrix, q, c1, c2) records
int ref_expectation_value(void* h, int n_gates, const int* nq, const double* mats, const uint64_t* qubits, double* out_re_im) {
  RefRegister* r = static_cast<RefRegister*>(h);
  return guarded([&] {
    std::vector<QC::Gates::AppliedGate<Mat>> gates;
    size_t off = 0;
    for (int i = 0; i < n_gates; ++i) {
      gates.emplace_back(fromRowMajor(nq[i], mats + off), qubits[3 * i], qubits[3 * i + 1], qubits[3 * i + 2]);
      off += 2 * (size_t(1) << nq[i]) * (size_t(1) << nq[i]);
    }
    const std::complex<double> v = r->ExpectationValue(gates);
    out_re_im[0] = v.real();
    out_re_im[1] = v.imag();
  });
}

// SaveState / RestoreState / RestoreStateDestructive (QubitRegister.h:600-616)
void ref_save_state(void* h) { static_cast<RefRegister*>(h)->SaveState(); }
void ref_restore_state(void* h, int destructive) {
  RefRegister* r = static_cast<RefRegister*>(h);
  if (destructive) r->RestoreStateDestructive();
  else r->RestoreState();
}

// ApplyOperatorMatrix (QubitRegister.h:499-505): dense 2^n x 2^n operator, row-major
int ref_apply_operator_matrix(void* h, const double* m) {
  RefRegister* r = static_cast<RefRegister*>(h);
  return guarded([&] {
    const size_t d = r->getNrBasisStates();
    Mat M(d, d);
    for (size_t i = 0; i < d; ++i)
      for (size_t j = 0; j < d; ++j) M(i, j) = std::complex<double>(m[2 * (i * d + j)], m[2 * (i * d + j) + 1]);
    r->ApplyOperatorMatrix(M);
  });
}

// Replay a circuit file (include/qcsim_b200.h "circuit files": the same file the engine replays through
// qcsim_sv_apply_circuit_file).  Records that name a reference gate class (gate_id >= 0) are applied as that class,
// with its virtual flags; the others as AppliedGate(matrix) -- the reference's ApplyGates path (QubitRegister.h:493-497).
// Returns the number of gates applied, or < 0.
long ref_apply_circuit_file(void* h, const char* path) {
  RefRegister* r = static_cast<RefRegister*>(h);
  long applied = -3;
  const int rc = guarded([&] {
    FILE* f = std::fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    char magic[8];
    uint32_t nq = 0, reserved = 0;
    uint64_t count = 0;
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "QCSIMC1\0", 8) == 0 && std::fread(&nq, 4, 1, f) == 1 &&
              std::fread(&reserved, 4, 1, f) == 1 && std::fread(&count, 8, 1, f) == 1;
    if (!ok || nq != r->getNrQubits()) {
      std::fclose(f);
      throw std::runtime_error("not a circuit file for this register");
    }
    for (uint64_t i = 0; i < count; ++i) {
      int32_t head[4];
      uint64_t q[3];
      double params[4], m[128];
      ok = std::fread(head, 4, 4, f) == 4 && head[0] >= 1 && head[0] <= 3 && std::fread(q, 8, 3, f) == 3 && std::fread(params, 8, 4, f) == 4;
      const size_t nm = ok ? (size_t)2 << (2 * head[0]) : 0;
      ok = ok && std::fread(m, 8, nm, f) == nm;
      if (!ok) {
        std::fclose(f);
        throw std::runtime_error("truncated circuit file");
      }
      std::unique_ptr<Gate> g = head[2] >= 0 ? makeGate(head[2], params) : nullptr;
      if (g) r->ApplyGate(*g, q[0], q[1], q[2]);
      else {
        QC::Gates::AppliedGate<Mat> ag(fromRowMajor(head[0], m), q[0], q[1], q[2]);
        r->ApplyGate(ag);
      }
    }
    std::fclose(f);
    applied = (long)count;
  });
  return rc == 0 ? applied : rc;
}

// rng.seed(seed) then `count` draws of `1. - uniformZeroOne(rng)`: pins qcsim_b200/rng.py
void ref_draws(void* h, uint64_t seed, int count, double* out) {
  RefRegister* r = static_cast<RefRegister*>(h);
  r->reseed(seed);
  for (int i = 0; i < count; ++i) out[i] = r->nextDraw();
}

// QuantumFourierTransform::QFT / IQFT on [sq, eq] (QuantumFourierTransform.h:35-87)
int ref_qft(void* h, uint64_t sq, uint64_t eq, int do_swap, int inverse) {
  RefRegister* r = static_cast<RefRegister*>(h);
  return guarded([&] {
    QC::SubAlgo::QuantumFourierTransform<Vec, Mat> f(r->getNrQubits(), sq, eq);
    if (inverse)
      f.IQFT(*r, do_swap != 0);
    else
      f.QFT(*r, do_swap != 0);
  });
}

// Gate recording (QubitRegister.h:536-590)
void ref_compute_start(void* h) { static_cast<RefRegister*>(h)->ComputeStart(); }
void ref_compute_end(void* h) { static_cast<RefRegister*>(h)->ComputeEnd(); }
void ref_compute(void* h) { static_cast<RefRegister*>(h)->Compute(); }
void ref_uncompute(void* h) { static_cast<RefRegister*>(h)->Uncompute(); }

// NControlledNotWithAncilla::Execute on the caller's register (NControlledNotWithAncilla.h:24-96)
int ref_ncnot(void* h, const uint64_t* ctrl, int n_ctrl, uint64_t target, uint64_t start_ancilla, int clear_ancilla) {
  RefRegister* r = static_cast<RefRegister*>(h);
  return guarded([&] {
    QC::SubAlgo::NControlledNotWithAncilla<Vec, Mat> nc(INT_MAX);
    std::vector<size_t> c(ctrl, ctrl + n_ctrl);
    nc.SetControlQubits(c);
    nc.SetTargetQubit(target);
    nc.SetStartAncillaQubits(start_ancilla);
    nc.SetClearAncillaAtTheEnd(clear_ancilla != 0);
    nc.Execute(*r);
  });
}

// GroverAlgorithmWithGatesOracle (GroverAlgorithm.h:128-242): Init + round(pi/4 sqrt(2^N))
// iterations (the reference's own count), no measurement; final amplitudes (2^(2N-1)) go to `out`.
int ref_grover_gates(int n_search, uint64_t marked, double* out, uint64_t* n_qubits_out) {
  return guarded([&] {
    GroverProbe g(static_cast<size_t>(n_search));
    g.setCorrectQuestionState(marked);
    g.run();
    const Vec& v = g.getRegisterStorage();
    if (n_qubits_out) *n_qubits_out = g.getNrQubits();
    if (out) std::memcpy(out, v.data(), sizeof(double) * 2 * static_cast<size_t>(v.size()));
  });
}

// DraperAdder known-answer circuit (DraperAdder.h:38-56) on |n1>|n2>: sub-register
// QFT(no swap) + controlled phases + IQFT(no swap); final amplitudes (2^(2 n_bits)) go to `out`.
int ref_draper_add(int n_bits, uint64_t n1, uint64_t n2, double* out) {
  return guarded([&] {
    DraperProbe a(static_cast<size_t>(n_bits));
    a.setToBasisState(n1 | (n2 << n_bits));
    a.run();
    const Vec& v = a.getRegisterStorage();
    std::memcpy(out, v.data(), sizeof(double) * 2 * static_cast<size_t>(v.size()));
  });
}

}  // extern "C"
