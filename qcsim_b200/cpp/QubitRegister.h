// QubitRegister.h -- drop-in replacement for QCSim's QC::QubitRegister on top of libqcsim_b200.so.
//
// Put this directory BEFORE the QCSim source directory on the include path: QCSim's algorithm
// headers (QuantumAlgorithm.h, QuantumFourierTransform.h, GroverAlgorithm.h, DraperAdder.h ...)
// `#include "QubitRegister.h"` / "QubitRegisterDebug.h" and then compile against this class
// unchanged, while QCSim's own gate headers (QuantumGate.h, SimpleGates.h) keep being used as they
// are.  Same public members, same exceptions, same silent no-ops as the reference class
// (reference file:line cited per member, relative to /root/reference/QCSim/).
//
// What changes (BASELINE.json north_star): the register storage is a device buffer owned by the
// engine handle; QubitRegisterCalculator's OpenMP loops are replaced by calls into the C ABI
// (include/qcsim_b200.h); the random number generator stays here, on the host, and draws are
// passed down, so seeded runs give the reference's outcomes.  There is no CPU fallback: without a
// CUDA device the constructor throws.
//
// Host-visible storage: getRegisterStorage() returns a reference to a lazily refreshed host
// mirror; above QCSIM_B200_MIRROR_LIMIT_QUBITS (default 31 -> 32 GiB) it throws std::length_error
// -- use DownloadRange() instead.
#pragma once

#define _USE_MATH_DEFINES
#include <math.h>
#include <Eigen/Eigen>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <complex>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "QuantumGate.h"  // QCSim's own gate definitions (QuantumGate.h -> SimpleGates.h)

#include "qcsim_b200.h"

#ifndef QCSIM_B200_MIRROR_LIMIT_QUBITS
#define QCSIM_B200_MIRROR_LIMIT_QUBITS 31
#endif

namespace QC {

	// QubitRegisterCalculator's public statics that client code touches (QubitRegisterCalculator.h:1256-1278).
	// The loops themselves live in the CUDA library now.
	template<class VectorClass = Eigen::VectorXcd, class MatrixClass = Eigen::MatrixXcd> class QubitRegisterCalculator {
	public:
		using GateClass = Gates::QuantumGateWithOp<MatrixClass>;

		QubitRegisterCalculator() = default;
		virtual ~QubitRegisterCalculator() = default;

		static int GetNumberOfThreads()
		{
			const size_t threads = std::thread::hardware_concurrency();
			return static_cast<int>(threads ? threads : 1);
		}

		void SetMultithreading(bool enable = true) { enableMultithreading = enable; }  // kept for source compatibility; the GPU path ignores it
		bool GetMultithreading() const { return enableMultithreading; }

		constexpr static size_t OneQubitOmpLimit = 8192;
		constexpr static size_t TwoQubitOmpLimit = OneQubitOmpLimit;
		constexpr static size_t ThreeQubitOmpLimit = OneQubitOmpLimit;

	private:
		bool enableMultithreading = true;
	};

	template<class VectorClass = Eigen::VectorXcd, class MatrixClass = Eigen::MatrixXcd> class QubitRegister : public QubitRegisterCalculator<VectorClass, MatrixClass>
	{
	public:
		using GateClass = Gates::QuantumGateWithOp<MatrixClass>;
		using BaseClass = QubitRegisterCalculator<VectorClass, MatrixClass>;

		// QubitRegister.h:17-37
		QubitRegister(size_t N = 3, unsigned int addseed = 0)
			: NrQubits(N), NrBasisStates(1ULL << NrQubits), uniformZeroOne(0, 1), recordGates(false)
		{
			assert(N > 0);
			CreateHandle(N);
			SeedFromClock(addseed);
		}

		// QubitRegister.h:40-57: the reference swaps the caller's vector in; here it is uploaded and
		// the caller's vector is left holding what it held (documented difference)
		QubitRegister(size_t N, VectorClass& v, unsigned int addseed = 0)
			: NrQubits(N), NrBasisStates(1ULL << NrQubits), uniformZeroOne(0, 1), recordGates(false)
		{
			assert(N > 0);
			CreateHandle(N);
			if (static_cast<size_t>(v.size()) == NrBasisStates) Upload(v);
			SeedFromClock(addseed);
		}

		QubitRegister(const QubitRegister&) = delete;
		QubitRegister& operator=(const QubitRegister&) = delete;

		QubitRegister(QubitRegister&& o) noexcept
			: NrQubits(o.NrQubits), NrBasisStates(o.NrBasisStates), handle(o.handle), mirror(std::move(o.mirror)), mirrorValid(o.mirrorValid),
			registerStorage{ this }, rng(o.rng), uniformZeroOne(0, 1), computeGates(std::move(o.computeGates)), recordGates(o.recordGates)
		{
			o.handle = nullptr;
		}

		~QubitRegister() override
		{
			if (handle) qcsim_sv_destroy(handle);
		}

		size_t getNrQubits() const { return NrQubits; };
		size_t getNrBasisStates() const { return NrBasisStates; };

		// :62-66 (silently 0 when out of range)
		std::complex<double> getBasisStateAmplitude(size_t State) const
		{
			if (State >= NrBasisStates) return 0;
			double v[2] = { 0, 0 };
			Check(qcsim_sv_get_amplitude(handle, State, v));
			return std::complex<double>(v[0], v[1]);
		}

		double getBasisStateProbability(size_t State) const  // :68-72
		{
			if (State >= NrBasisStates) return 0;
			return std::norm(getBasisStateAmplitude(State));
		}

		void setToBasisState(size_t State)  // :74-80
		{
			if (State >= NrBasisStates) return;
			Touch();
			Check(qcsim_sv_set_basis_state(handle, State));
		}

		void setToQubitState(size_t q)  // :82-88
		{
			if (q >= NrQubits) return;
			setToBasisState(1ULL << q);
		}

		void setToCatState()  // :91-98
		{
			Clear();
			static const double OneOverSqrt2 = 1. / sqrt(2.);
			Check(qcsim_sv_set_amplitude(handle, 0, OneOverSqrt2, 0));
			Check(qcsim_sv_set_amplitude(handle, NrBasisStates - 1, OneOverSqrt2, 0));
		}

		void Reset() { setToBasisState(0); }  // :100-103

		void setToEqualSuperposition()  // :106-109
		{
			Touch();
			Check(qcsim_sv_fill(handle, 1. / sqrt(static_cast<double>(NrBasisStates)), 0));
		}

		void setRawAmplitude(size_t State, std::complex<double> val)  // :112-117
		{
			if (State >= NrBasisStates) return;
			Touch();
			Check(qcsim_sv_set_amplitude(handle, State, val.real(), val.imag()));
		}

		void Clear()  // :119-122
		{
			Touch();
			Check(qcsim_sv_fill(handle, 0, 0));
		}

		void Normalize()  // :124-130 (no-op when the norm is < 1e-20)
		{
			Touch();
			Check(qcsim_sv_normalize(handle));
		}

		// :133-165
		void AdjustPhaseAndNormalize()
		{
			std::complex<double> v0 = getBasisStateAmplitude(0);
			if (abs(v0) < 1E-5) v0 = getBasisStateAmplitude(NrBasisStates >> 1);
			if (abs(v0) < 1E-5) v0 = getBasisStateAmplitude(NrBasisStates - 1);
			if (abs(v0) >= 1E-5)
			{
				// psi /= v0 as a diagonal one-qubit gate diag(1/v0, 1/v0): one in-place pass on the device
				const std::complex<double> f = 1. / v0;
				const double m[8] = { f.real(), f.imag(), 0, 0, 0, 0, f.real(), f.imag() };
				Touch();
				Check(qcsim_sv_apply(handle, 1, m, QCSIM_GATE_DIAGONAL, 0, 0, 0));
			}
			Normalize();
		}

		size_t MeasureAll()  // :169-195
		{
			const double prob = 1. - uniformZeroOne(rng);
			uint64_t out = 0;
			Touch();
			Check(qcsim_sv_measure_all(handle, prob, &out));
			return static_cast<size_t>(out);
		}

		size_t MeasureQubit(size_t qubit) { return Measure(qubit, qubit); }  // :198-206

		size_t Measure(size_t firstQubit, size_t secondQubit)  // :208-224
		{
			const double prob = 1. - uniformZeroOne(rng);
			uint64_t out = 0;
			Touch();
			Check(qcsim_sv_measure(handle, firstQubit, secondQubit, prob, &out));
			return static_cast<size_t>(out);
		}

		// :227-273: nrTimes draws against ONE cumulative table.  The table is the reference's sequential running sum,
		// reproduced bit for bit on the device, including its cut at 1 - epsilon and the end() outcome (:250-254, 268);
		// a single shot takes the reference's MeasureNoCollapse shortcut (:231-236)
		std::map<size_t, size_t> RepeatedMeasure(size_t nrTimes = 1000)
		{
			std::map<size_t, size_t> measurements;
			if (nrTimes == 0) return measurements;
			if (nrTimes == 1) { ++measurements[MeasureNoCollapse()]; return measurements; }
			for (uint64_t s : SampleStates(nrTimes)) ++measurements[static_cast<size_t>(s)];
			return measurements;
		}

		std::unordered_map<size_t, size_t> RepeatedMeasureUnordered(size_t nrTimes = 1000)  // :276-322
		{
			std::unordered_map<size_t, size_t> measurements;
			if (nrTimes == 0) return measurements;
			if (nrTimes == 1) { ++measurements[MeasureNoCollapse()]; return measurements; }
			for (uint64_t s : SampleStates(nrTimes)) ++measurements[static_cast<size_t>(s)];
			return measurements;
		}

		std::map<size_t, size_t> RepeatedMeasure(size_t firstQubit, size_t secondQubit, size_t nrTimes = 1000)  // :325-375
		{
			std::map<size_t, size_t> measurements;
			if (nrTimes == 0) return measurements;
			const size_t mask = MeasuredMask(firstQubit, secondQubit);
			if (nrTimes == 1)  // :334-339: the already shifted outcome is masked and shifted once more, as in the reference
			{
				const size_t meas = MeasureNoCollapse(firstQubit, secondQubit);
				++measurements[(meas & mask) >> firstQubit];
				return measurements;
			}
			for (uint64_t s : SampleStates(nrTimes)) ++measurements[(static_cast<size_t>(s) & mask) >> firstQubit];
			return measurements;
		}

		std::unordered_map<size_t, size_t> RepeatedMeasureUnordered(size_t firstQubit, size_t secondQubit, size_t nrTimes = 1000)  // :378-429
		{
			std::unordered_map<size_t, size_t> measurements;
			if (nrTimes == 0) return measurements;
			const size_t mask = MeasuredMask(firstQubit, secondQubit);
			if (nrTimes == 1)
			{
				const size_t meas = MeasureNoCollapse(firstQubit, secondQubit);
				++measurements[(meas & mask) >> firstQubit];
				return measurements;
			}
			for (uint64_t s : SampleStates(nrTimes)) ++measurements[(static_cast<size_t>(s) & mask) >> firstQubit];
			return measurements;
		}

		// :434-486.  The kernel is selected by the gate's virtual flags exactly as
		// QubitRegisterCalculator does (:39-227): they cross the ABI as QCSIM_GATE_* bits.
		void ApplyGate(const GateClass& gate, size_t qubit, size_t controllingQubit1 = 0, size_t controllingQubit2 = 0)
		{
			const size_t gateQubits = gate.getQubitsNumber();
			CheckQubits(gate, qubit, controllingQubit1, controllingQubit2, gateQubits);
			assert(gateQubits > 0 && gateQubits <= 3);

			const MatrixClass& gateMatrix = gate.getRawOperatorMatrix();
			const int d = 1 << gateQubits;
			double m[128];
			for (int r = 0; r < d; ++r)
				for (int c = 0; c < d; ++c)
				{
					const std::complex<double> z = gateMatrix(r, c);  // Eigen is column-major; the ABI is row-major
					m[2 * (r * d + c)] = z.real();
					m[2 * (r * d + c) + 1] = z.imag();
				}
			int flags = 0;
			if (gate.isControlled()) flags |= QCSIM_GATE_CONTROLLED;
			if (gateQubits == 3 && gate.isControlQubit(1)) flags |= QCSIM_GATE_TWO_CONTROLS;
			if (gate.isDiagonal()) flags |= QCSIM_GATE_DIAGONAL;
			if (gate.isAntidiagonal()) flags |= QCSIM_GATE_ANTIDIAGONAL;
			if (gate.isSwapGate()) flags |= QCSIM_GATE_SWAP;
			if (gate.IsISwapGate()) flags |= QCSIM_GATE_ISWAP;
			if (gate.IsISwapDagGate()) flags |= QCSIM_GATE_ISWAPDAG;
			Touch();
			Check(qcsim_sv_apply(handle, static_cast<int>(gateQubits), m, flags, qubit, controllingQubit1, controllingQubit2));

			if (recordGates)
				computeGates.emplace_back(Gates::AppliedGate<MatrixClass>(gate.getRawOperatorMatrix(), qubit, controllingQubit1, controllingQubit2));
		}

		void ApplyGate(const Gates::AppliedGate<MatrixClass>& gate)  // :488-491
		{
			ApplyGate(gate, gate.getQubit1(), gate.getQubit2(), gate.getQubit3());
		}

		void ApplyGates(const std::vector<Gates::AppliedGate<MatrixClass>>& gates)  // :493-497
		{
			for (const auto& gate : gates)
				ApplyGate(gate);
		}

		// :499-505: registerStorage = m * registerStorage for a dense 2^n x 2^n operator -- a device GEMV, small registers
		// only (QCSIM_MAX_OPERATOR_QUBITS); this is what Compute / Uncompute replay for recorded gates on more than
		// three qubits (:563, 581) and what the dense-oracle algorithms (Shor, Grover, phase estimation) apply
		void ApplyOperatorMatrix(const MatrixClass& m)
		{
			const size_t d = NrBasisStates;
			if (static_cast<size_t>(m.rows()) != d || static_cast<size_t>(m.cols()) != d) throw std::invalid_argument("qcsim_b200: operator matrix must be 2^n x 2^n");
			std::vector<double> rowMajor(2 * d * d);
			for (size_t r = 0; r < d; ++r)
				for (size_t c = 0; c < d; ++c)
				{
					const std::complex<double> z = m(r, c);  // Eigen is column-major; the ABI is row-major
					rowMajor[2 * (r * d + c)] = z.real();
					rowMajor[2 * (r * d + c) + 1] = z.imag();
				}
			Touch();
			Check(qcsim_sv_apply_operator(handle, rowMajor.data()));

			if (recordGates)
				computeGates.emplace_back(Gates::AppliedGate<MatrixClass>(m));
		}

		const VectorClass& getRegisterStorage() const  // :507-510
		{
			if (!mirrorValid)
			{
				if (NrQubits > QCSIM_B200_MIRROR_LIMIT_QUBITS) throw std::length_error("qcsim_b200: register too large for a host mirror, use DownloadRange");
				if (static_cast<size_t>(mirror.size()) != NrBasisStates) mirror.resize(NrBasisStates);
				Check(qcsim_sv_download(handle, reinterpret_cast<double*>(&mirror(0)), 0, NrBasisStates));
				mirrorValid = true;
			}
			return mirror;
		}

		void setRegisterStorage(const VectorClass& vals)  // :512-518
		{
			if (NrBasisStates != static_cast<size_t>(vals.size())) return;
			Upload(vals);
			Normalize();
		}

		void setRegisterStorageFastNoNormalize(VectorClass& vals)  // :521-524 (uploads; `vals` is left untouched)
		{
			if (NrBasisStates != static_cast<size_t>(vals.size())) return;
			Upload(vals);
		}

		double stateFidelity(const VectorClass& state) const  // :527-534
		{
			if (NrBasisStates != static_cast<size_t>(state.size())) return 0;
			qcsim_sv* other = NewHandle(NrQubits);  // same device(s) as this register
			int rc = qcsim_sv_upload(other, reinterpret_cast<const double*>(&state(0)), 0, NrBasisStates);
			double p[2] = { 0, 0 };
			if (rc == QCSIM_OK) rc = qcsim_sv_inner_product(handle, other, p);  // conj(register) . state
			qcsim_sv_destroy(other);
			Check(rc);
			return p[0] * p[0] + p[1] * p[1];
		}

		void ComputeStart()  // :536-540
		{
			recordGates = true;
			computeGates.clear();
		}

		void ComputeEnd() { recordGates = false; }  // :542-545
		void ComputeClear() { computeGates.clear(); }  // :547-550

		void Compute()  // :554-569
		{
			const bool recordSave = recordGates;
			recordGates = false;
			for (const Gates::AppliedGate<MatrixClass>& gate : computeGates)
			{
				if (gate.getQubitsNumber() > 3)
					ApplyOperatorMatrix(gate.getRawOperatorMatrix());
				else
					ApplyGate(gate);
			}
			recordGates = recordSave;
		}

		void Uncompute()  // :573-590
		{
			const bool recordSave = recordGates;
			recordGates = false;
			for (auto it = computeGates.crbegin(); it != computeGates.crend(); ++it)
			{
				if (it->getQubitsNumber() > 3)
					ApplyOperatorMatrix(it->getRawOperatorMatrix().adjoint());
				else
				{
					Gates::AppliedGate<MatrixClass> gate(it->getRawOperatorMatrix().adjoint(), it->getQubit1(), it->getQubit2(), it->getQubit3());
					ApplyGate(gate);
				}
			}
			recordGates = recordSave;
		}

		double GetQubitProbability(size_t qubit) const  // :592-598
		{
			double p = 0;
			Check(qcsim_sv_qubit_probability(handle, qubit, &p));
			return p;
		}

		void SaveState() { Check(qcsim_sv_save_state(handle)); }  // :600-603

		void RestoreState()  // :605-609
		{
			Touch();
			Check(qcsim_sv_restore_state(handle, 0));
		}

		void RestoreStateDestructive()  // :611-616
		{
			Touch();
			Check(qcsim_sv_restore_state(handle, 1));
		}

		size_t MeasureNoCollapse()  // :619-642
		{
			const double prob = 1. - uniformZeroOne(rng);
			uint64_t out = 0;
			Check(qcsim_sv_measure_all_nocollapse(handle, prob, &out));
			return static_cast<size_t>(out);
		}

		// :646-660: <psi| G_k ... G_1 |psi> with the gates applied to a device-side copy
		std::complex<double> ExpectationValue(const std::vector<Gates::AppliedGate<MatrixClass>>& gates)
		{
			if (gates.empty()) return 1.;
			std::unique_ptr<QubitRegister> work = Clone();
			work->recordGates = false;
			work->ApplyGates(gates);
			double p[2] = { 0, 0 };
			Check(qcsim_sv_inner_product(handle, work->handle, p));
			return std::complex<double>(p[0], p[1]);
		}

		std::unique_ptr<QubitRegister<VectorClass, MatrixClass>> Clone() const  // :662-674
		{
			qcsim_sv* h2 = nullptr;
			Check(qcsim_sv_clone(handle, &h2));
			std::unique_ptr<QubitRegister> qr(new QubitRegister(NrQubits, h2));
			qr->computeGates = computeGates;
			qr->recordGates = recordGates;
			return qr;
		}

		// ---- extensions (not in the reference) ------------------------------------------------------

		// deterministic seeding for reproducible runs (the reference seeds from the clock, :26-34)
		void Seed(uint64_t s)
		{
			std::seed_seq seed{ uint32_t(s & 0xffffffff), uint32_t(s >> 32) };
			rng.seed(seed);
		}

		// true: ApplyGate only queues; the queue is cut into fused shared-memory gate blocks (several
		// gates per pass over HBM) and flushed by the next call that observes the state
		void SetFusion(bool enable) { Check(qcsim_sv_set_fusion(handle, enable ? 1 : 0)); }
		void Flush() { Check(qcsim_sv_sync(handle)); }

		// QuantumFourierTransform::QFT / IQFT (QuantumFourierTransform.h:35-87) as one engine call
		void ApplyQFT(size_t startQubit, size_t endQubit, bool doSwap, bool inverse)
		{
			Touch();
			Check(qcsim_sv_qft(handle, startQubit, endQubit, doSwap ? 1 : 0, inverse ? 1 : 0));
		}

		void DownloadRange(std::complex<double>* out, size_t first, size_t count) const
		{
			Check(qcsim_sv_download(handle, reinterpret_cast<double*>(out), first, count));
		}

		double Norm2() const
		{
			double v = 0;
			Check(qcsim_sv_norm2(handle, &v));
			return v;
		}

		qcsim_sv* Handle() const { return handle; }

	protected:
		QubitRegister(size_t N, qcsim_sv* adopt)
			: NrQubits(N), NrBasisStates(1ULL << NrQubits), handle(adopt), uniformZeroOne(0, 1), recordGates(false)
		{
			SeedFromClock(0);
		}

		// :677-690 -- same tests, same order, same messages
		inline void CheckQubits(const GateClass& /*gate*/, size_t qubit, size_t controllingQubit1, size_t controllingQubit2, size_t gateQubits) const
		{
			if (NrQubits == 0) throw std::invalid_argument("Qubit number is zero");
			else if (NrQubits <= qubit) throw std::invalid_argument("Qubit number is too high");
			else if (gateQubits == 2) {
				if (NrQubits <= controllingQubit1) throw std::invalid_argument("Controlling qubit number is too high");
				else if (qubit == controllingQubit1) throw std::invalid_argument("Qubit and controlling qubit are the same");
			}
			else if (gateQubits == 3)
			{
				if (NrQubits <= controllingQubit1 || NrQubits <= controllingQubit2) throw std::invalid_argument("Controlling qubit number is too high");
				else if (qubit == controllingQubit1 || qubit == controllingQubit2 || controllingQubit1 == controllingQubit2) throw std::invalid_argument("Qubits must be different");
			}
		}

		size_t MeasureNoCollapse(size_t qubit) { return MeasureNoCollapse(qubit, qubit); }  // :695-698

		size_t MeasureNoCollapse(size_t firstQubit, size_t secondQubit)  // :705-713
		{
			const double prob = 1. - uniformZeroOne(rng);
			uint64_t out = 0;
			Check(qcsim_sv_measure_nocollapse(handle, firstQubit, secondQubit, prob, &out));
			return static_cast<size_t>(out);
		}

		static void Check(int rc)
		{
			if (rc == QCSIM_OK) return;
			const char* msg = qcsim_last_error();
			const std::string text = msg ? msg : "qcsim_b200 error";
			switch (rc)
			{
			case QCSIM_ERR_QUBIT_TOO_HIGH:
			case QCSIM_ERR_CTRL_TOO_HIGH:
			case QCSIM_ERR_SAME_QUBITS:
				throw std::invalid_argument(text);
			case QCSIM_ERR_BAD_STATE:
				return;  // the reference ignores out-of-range basis states silently
			case QCSIM_ERR_OOM:
				throw std::bad_alloc();
			default:
				throw std::runtime_error("qcsim_b200: " + text);
			}
		}

		static int DefaultDevice()
		{
			const char* s = std::getenv("QCSIM_B200_DEVICE");
			return s ? std::atoi(s) : 0;
		}

		// QCSIM_B200_DEVICES=0,1,2,3 shards the register over those GPUs of this process (qcsim_sv_create_multi);
		// registers too small to shard (fewer than 8 qubits per device) stay on the first device
		void CreateHandle(size_t N)
		{
			handle = NewHandle(N);
			// QCSIM_B200_FUSION=1: ApplyGate only queues; the queue runs as fused gate blocks when the state is observed
			if (const char* f = std::getenv("QCSIM_B200_FUSION"))
				if (std::atoi(f) != 0) Check(qcsim_sv_set_fusion(handle, 1));
		}

		static qcsim_sv* NewHandle(size_t N)
		{
			qcsim_sv* out = nullptr;
			std::vector<int> ids;
			if (const char* s = std::getenv("QCSIM_B200_DEVICES"))
			{
				std::string item;
				for (const char* p = s;; ++p)
				{
					if (*p == ',' || *p == 0)
					{
						if (!item.empty()) ids.push_back(std::atoi(item.c_str()));
						item.clear();
						if (*p == 0) break;
					}
					else item.push_back(*p);
				}
			}
			size_t log2w = 0;
			while ((size_t(1) << (log2w + 1)) <= ids.size()) ++log2w;
			if (ids.size() > 1 && N >= log2w + 8)
				Check(qcsim_sv_create_multi(&out, static_cast<int>(N), 1 << log2w, ids.data()));
			else
				Check(qcsim_sv_create(&out, static_cast<int>(N), ids.empty() ? DefaultDevice() : ids[0]));
			return out;
		}

		static size_t MeasuredMask(size_t firstQubit, size_t secondQubit)  // QubitRegisterCalculator.h:1128-1130
		{
			const size_t secondQubitp1 = secondQubit + 1;
			const size_t firstPartMask = (1ULL << firstQubit) - 1;
			return ((1ULL << secondQubitp1) - 1) ^ firstPartMask;
		}

		std::vector<uint64_t> SampleStates(size_t nrTimes)
		{
			std::vector<double> probs(nrTimes);
			for (size_t i = 0; i < nrTimes; ++i) probs[i] = 1. - uniformZeroOne(rng);
			std::vector<uint64_t> out(nrTimes);
			if (nrTimes) Check(qcsim_sv_sample(handle, probs.data(), nrTimes, out.data()));
			return out;
		}

		void Upload(const VectorClass& v)
		{
			Touch();
			Check(qcsim_sv_upload(handle, reinterpret_cast<const double*>(&v(0)), 0, NrBasisStates));
		}

		void Touch() { mirrorValid = false; }

		void SeedFromClock(unsigned int addseed)  // :26-34
		{
			if (addseed == 0)
			{
				std::random_device rdl;
				addseed = rdl();
			}
			const uint64_t timeSeed = std::chrono::high_resolution_clock::now().time_since_epoch().count() + addseed;
			Seed(timeSeed);
		}

		size_t NrQubits;
		size_t NrBasisStates;

		qcsim_sv* handle = nullptr;       // replaces registerStorage / resultsStorage / savedStateStorage (:718-721)
		mutable VectorClass mirror;       // host copy handed out by getRegisterStorage()
		mutable bool mirrorValid = false;

		// QCSim's own QubitRegisterDebug.h reads the protected member `registerStorage(i)` (QubitRegisterDebug.h:30-34).
		// Here the storage is on the device, so the name is a small read-only proxy onto the host mirror: the
		// reference header compiles against this class unchanged.
		struct StorageProxy
		{
			const QubitRegister* owner;
			std::complex<double> operator()(size_t i) const { return owner->getRegisterStorage()(i); }
			std::complex<double> operator[](size_t i) const { return owner->getRegisterStorage()(i); }
			size_t size() const { return owner->NrBasisStates; }
		};
		StorageProxy registerStorage{ this };

		std::mt19937_64 rng;
		std::uniform_real_distribution<double> uniformZeroOne;

		std::vector<Gates::AppliedGate<MatrixClass>> computeGates;
		bool recordGates;
	};

}
