// facade_test.cpp -- drives QCSim's OWN algorithm headers (QuantumFourierTransform.h,
// GroverAlgorithm.h, DraperAdder.h, NControlledNotWithAncilla.h, QuantumAlgorithm.h -- compiled
// from /root/reference/QCSim, unmodified) on top of the drop-in register of
// qcsim_b200/cpp/QubitRegister.h.  Built here (where the reference tree exists) by
// tests/cpp/build_facade.sh into tests/cpp/facade_test[_fast].bin; the binary travels to the GPU
// box and tests/test_gpu_facade.py compares what it prints / dumps with the oracle.
// With -DFACADE_FAST the include path also shadows QuantumFourierTransform.h with the one-call version.
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "QubitRegister.h"
#include "QubitRegisterDebug.h"  // QCSim's own header, unchanged: reads the protected registerStorage(i)
#include "QuantumFourierTransform.h"
#include "GroverAlgorithm.h"
#include "DraperAdder.h"
#include "NControlledNotWithAncilla.h"

using Vec = Eigen::VectorXcd;
using Mat = Eigen::MatrixXcd;

class TestRegister : public QC::QubitRegister<Vec, Mat> {
public:
  using Base = QC::QubitRegister<Vec, Mat>;
  explicit TestRegister(size_t n) : Base(n, 12345u) {}
  void reseed(uint64_t seed) {  // same as the oracle driver's RefRegister::reseed
    rng.seed(seed);
    uniformZeroOne = std::uniform_real_distribution<double>(0, 1);
  }
  size_t measureNoCollapseRange(size_t a, size_t b) { return Base::MeasureNoCollapse(a, b); }
};
class GroverProbe : public Grover::GroverAlgorithmWithGatesOracle<Vec, Mat> {
public:
  explicit GroverProbe(size_t n) : Grover::GroverAlgorithmWithGatesOracle<Vec, Mat>(n, 12345u) {}
  void run() { ExecuteWithoutMeasurement(); }
  auto& R() { return reg; }  // the drop-in register inside the reference algorithm (QuantumAlgorithm.h:204)
};
class DraperProbe : public Adders::DraperAdder<Vec, Mat> {
public:
  explicit DraperProbe(size_t n) : Adders::DraperAdder<Vec, Mat>(n, 12345u) {}
  void run() { ExecuteWithoutMeasurement(); }
};

static Vec read_state(const char* path, size_t dim) {
  Vec v(dim);
  FILE* f = std::fopen(path, "rb");
  if (!f || std::fread(&v(0), 16, dim, f) != dim) {
    std::fprintf(stderr, "cannot read %s\n", path);
    std::exit(2);
  }
  std::fclose(f);
  return v;
}
static void append_state(const char* path, const Vec& v) {
  FILE* f = std::fopen(path, "ab");
  std::fwrite(&v(0), 16, (size_t)v.size(), f);
  std::fclose(f);
}

template <class F> static void expect_throw(const char* what, F&& fn) {
  try {
    fn();
    std::printf("%s: no exception\n", what);
  } catch (const std::invalid_argument& e) {
    std::printf("%s: invalid_argument: %s\n", what, e.what());
  } catch (const std::exception& e) {
    std::printf("%s: exception: %s\n", what, e.what());
  }
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  const std::string mode = argv[1];
  try {
    if (mode == "qft") {  // qft n sq eq do_swap fusion in.bin out.bin : QFT, dump, IQFT, dump
      const size_t n = std::stoul(argv[2]), sq = std::stoul(argv[3]), eq = std::stoul(argv[4]);
      const bool do_swap = std::stoi(argv[5]) != 0, fusion = std::stoi(argv[6]) != 0;
      TestRegister reg(n);
      Vec psi = read_state(argv[7], 1ULL << n);
      reg.setRegisterStorageFastNoNormalize(psi);
      reg.SetFusion(fusion);
      QC::SubAlgo::QuantumFourierTransform<Vec, Mat> f(n, sq, eq);
      f.QFT(reg, do_swap);
      append_state(argv[8], reg.getRegisterStorage());
      f.IQFT(reg, do_swap);
      append_state(argv[8], reg.getRegisterStorage());
      std::printf("norm2 %.17g\n", reg.Norm2());
    } else if (mode == "grover") {  // grover n_search marked out.bin
      GroverProbe g(std::stoul(argv[2]));
      g.setCorrectQuestionState(std::stoul(argv[3]));
      g.run();
      append_state(argv[4], g.getRegisterStorage());
      std::printf("qubits %zu\n", g.getNrQubits());
    } else if (mode == "grover_ranges") {  // grover_ranges n_search marked out.bin : 16 sampled ranges of 4096 amplitudes + P(marked)
      const size_t ns = std::stoul(argv[2]), marked = std::stoul(argv[3]);
      GroverProbe g(ns);
      g.setCorrectQuestionState(marked);
      g.run();
      const size_t nq = g.getNrQubits(), dim = size_t(1) << nq, cnt = 4096;
      Vec buf(cnt);
      for (size_t j = 0; j < 16; ++j) {
        const size_t first = ((dim / 16) * j + 4096 * j) & ~size_t(4095);
        g.R().DownloadRange(&buf(0), first < dim - cnt ? first : dim - cnt, cnt);
        append_state(argv[4], buf);
      }
      std::printf("qubits %zu\n", nq);
      std::printf("norm2 %.17g\n", g.R().Norm2());
      // every search qubit (the low ns bits) reads its bit of the marked string with probability ~1 after the
      // reference's round(pi/4 sqrt(2^ns)) iterations
      for (size_t q = 0; q < ns; ++q) std::printf("pq %zu %.17g\n", q, g.R().GetQubitProbability(q));
    } else if (mode == "draper") {  // draper n_bits n1 n2 out.bin
      const size_t nb = std::stoul(argv[2]), n1 = std::stoul(argv[3]), n2 = std::stoul(argv[4]);
      DraperProbe a(nb);
      a.setToBasisState(n1 | (n2 << nb));
      a.run();
      append_state(argv[5], a.getRegisterStorage());
      std::printf("measured %zu\n", a.Measure());  // deterministic known answer: |n1>|n1+n2 mod 2^nb>
    } else if (mode == "ncnot") {  // ncnot n out.bin : n-controlled NOT ladder with Compute/Uncompute on a superposition
      const size_t n = std::stoul(argv[2]);
      TestRegister reg(n);
      QC::Gates::HadamardGate<Mat> h;
      const size_t nc = (n + 1) / 2;  // controls 0..nc-1, target nc, ancillas nc+1..
      for (size_t q = 0; q < nc; ++q) reg.ApplyGate(h, q);
      QC::SubAlgo::NControlledNotWithAncilla<Vec, Mat> ncn(INT_MAX);
      std::vector<size_t> c;
      for (size_t q = 0; q < nc; ++q) c.push_back(q);
      ncn.SetControlQubits(c);
      ncn.SetTargetQubit(nc);
      ncn.SetStartAncillaQubits(nc + 1);
      ncn.SetClearAncillaAtTheEnd(true);
      ncn.Execute(reg);
      append_state(argv[3], reg.getRegisterStorage());
    } else if (mode == "measure") {  // measure n seed in.bin out.bin
      const size_t n = std::stoul(argv[2]);
      const uint64_t seed = std::stoull(argv[3]);
      TestRegister reg(n);
      Vec psi = read_state(argv[4], 1ULL << n);
      reg.setRegisterStorageFastNoNormalize(psi);
      reg.reseed(seed);
      for (int i = 0; i < 4; ++i) std::printf("nocollapse %zu\n", reg.MeasureNoCollapse());
      std::printf("range_nocollapse %zu\n", reg.measureNoCollapseRange(1, n - 2));
      std::printf("p0 %.17g\n", reg.GetQubitProbability(0));
      std::printf("qubit %zu\n", reg.MeasureQubit(n - 1));
      append_state(argv[5], reg.getRegisterStorage());
      std::printf("range %zu\n", reg.Measure(0, 2));
      append_state(argv[5], reg.getRegisterStorage());
      std::printf("all %zu\n", reg.MeasureAll());
      append_state(argv[5], reg.getRegisterStorage());
    } else if (mode == "misc") {  // error conventions, silent no-ops, state helpers, clone, expectation value
      TestRegister reg(4);
      QC::Gates::HadamardGate<Mat> h;
      QC::Gates::CNOTGate<Mat> cx;
      QC::Gates::ToffoliGate<Mat> ccx;
      QC::Gates::PauliZGate<Mat> z;
      expect_throw("1q too high", [&] { reg.ApplyGate(h, 4); });
      expect_throw("2q ctrl too high", [&] { reg.ApplyGate(cx, 1, 7); });
      expect_throw("2q same", [&] { reg.ApplyGate(cx, 2, 2); });
      expect_throw("3q ctrl too high", [&] { reg.ApplyGate(ccx, 0, 1, 9); });
      expect_throw("3q same", [&] { reg.ApplyGate(ccx, 0, 1, 1); });
      reg.setToBasisState(99);  // silently ignored (QubitRegister.h:76)
      reg.setRawAmplitude(99, 1.);
      std::printf("amp_out_of_range %.1f\n", std::abs(reg.getBasisStateAmplitude(99)));
      std::printf("amp0 %.17g\n", reg.getBasisStateAmplitude(0).real());
      reg.setToCatState();
      std::printf("cat %.17g %.17g\n", reg.getBasisStateAmplitude(0).real(), reg.getBasisStateAmplitude(15).real());
      reg.setToEqualSuperposition();
      std::printf("equal %.17g\n", reg.getBasisStateAmplitude(7).real());
      reg.setToBasisState(0);
      reg.ApplyGate(h, 0);
      std::vector<QC::Gates::AppliedGate<Mat>> zs;
      zs.emplace_back(z.getRawOperatorMatrix(), 0);
      std::printf("expect_z_plus %.17g\n", reg.ExpectationValue(zs).real());  // <+|Z|+> = 0  (Tests.cpp:665-723)
      reg.setToBasisState(1);
      std::printf("expect_z_one %.17g\n", reg.ExpectationValue(zs).real());   // -1
      auto c = reg.Clone();
      c->ApplyGate(h, 1);
      std::printf("clone_indep %.17g %.17g\n", reg.getBasisStateAmplitude(1).real(), c->getBasisStateAmplitude(1).real());
      reg.SaveState();
      reg.ApplyGate(h, 2);
      reg.RestoreState();
      std::printf("restored %.17g\n", reg.getBasisStateAmplitude(1).real());
      Vec want(16);
      for (int i = 0; i < 16; ++i) want(i) = 0;
      want(1) = 1;
      std::printf("fidelity %.17g\n", reg.stateFidelity(want));
      reg.setRawAmplitude(0, std::complex<double>(0, 2));  // state (2i, 1, 0, ...): divided by 2i, then normalised (QubitRegister.h:133-165)
      reg.AdjustPhaseAndNormalize();
      std::printf("adjusted %.17g %.17g %.17g\n", reg.getBasisStateAmplitude(0).real(), reg.getBasisStateAmplitude(0).imag(), reg.getBasisStateAmplitude(1).imag());
      auto hist = reg.RepeatedMeasure(50);
      std::printf("repeated %zu %zu\n", hist.size(), hist.count(1) ? hist[1] : 0);
      std::printf("threads_ok %d\n", TestRegister::GetNumberOfThreads() > 0);
      {  // QCSim's QubitRegisterDebug (QubitRegisterDebug.h:20-43) on the drop-in register
        QC::QubitRegisterDebug<Vec, Mat> dbg(3, 12345u);
        dbg.ApplyGate(h, 1);
        const std::string path = "/tmp/qcsim_b200_facade_debug.txt";
        const bool ok = dbg.writeToFile(path, true, false);
        std::ifstream in(path);
        size_t idx = 0, lines = 0;
        double val = 0, sumsq = 0;
        while (in >> idx >> val) { ++lines; sumsq += val * val; }
        std::printf("debug_dump %d %zu %.12f\n", ok ? 1 : 0, lines, sumsq);
      }
    } else {
      return 1;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "facade_test failed: %s\n", e.what());
    return 3;
  }
  return 0;
}
