// QuantumFourierTransform.h -- optional shadow of QCSim's header of the same name
// (QuantumFourierTransform.h:1-98).  QCSim's own version also works unchanged on the device-backed
// register (gate by gate through ApplyGate, fused when the register is in fusion mode); this one
// hands the whole transform to the engine in one call (qcsim_sv_qft), which runs it as fused
// shared-memory gate blocks.  Same class, same members, same results to rounding.
#pragma once

#include "QubitsSwapper.h"

#define _USE_MATH_DEFINES
#include <math.h>

namespace QC {

	namespace SubAlgo {

		template<class VectorClass = Eigen::VectorXcd, class MatrixClass = Eigen::MatrixXcd> class QuantumFourierTransform : public QubitsSwapper<VectorClass, MatrixClass>
		{
		public:
			using BaseClass = QubitsSwapper<VectorClass, MatrixClass>;
			using RegisterClass = QubitRegister<VectorClass, MatrixClass>;

			QuantumFourierTransform(size_t N, size_t startQubit = 0, size_t endQubit = INT_MAX)
				: BaseClass(N, startQubit, endQubit)
			{
			}

			size_t Execute(RegisterClass& reg) override  // :23-28
			{
				QFT(reg);
				return reg.MeasureAll();
			}

			void QFT(RegisterClass& reg, bool doSwap = true)  // :35-60
			{
				reg.ApplyQFT(BaseClass::BaseClass::getStartQubit(), BaseClass::BaseClass::getEndQubit(), doSwap, false);
			}

			void IQFT(RegisterClass& reg, bool doSwap = true)  // :62-87
			{
				reg.ApplyQFT(BaseClass::BaseClass::getStartQubit(), BaseClass::BaseClass::getEndQubit(), doSwap, true);
			}

			// public in the reference (:90-91); kept so client code that borrows them still compiles
			Gates::HadamardGate<MatrixClass> hadamard;
			Gates::ControlledPhaseShiftGate<MatrixClass> cPhaseShift;
		};

	}

}
