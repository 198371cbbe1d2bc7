// QubitRegisterDebug.h -- shadows QCSim's header of the same name (QubitRegisterDebug.h:1-94): the
// reference version reads the protected host vector `registerStorage` directly; this one goes
// through the public getters of the device-backed register.  Same members, same output format.
#pragma once

#include "QubitRegister.h"

namespace QC {

	template<class VectorClass = Eigen::VectorXcd, class MatrixClass = Eigen::MatrixXcd> class QubitRegisterDebug : public QubitRegister<VectorClass, MatrixClass>
	{
	public:
		using BaseClass = QubitRegister<VectorClass, MatrixClass>;

		QubitRegisterDebug(size_t N = 3, unsigned int addseed = 0)
			: BaseClass(N, addseed)
		{
		}

		// QubitRegisterDebug.h:20-43
		bool writeToFile(const std::string& name, bool amplitude = true, bool append = false) const
		{
			try {
				std::ofstream thefile;
				thefile.open(name, std::ios::out | (append ? std::ios::app : std::ios::trunc));
				if (!thefile.is_open()) return false;
				if (append) thefile << std::endl << std::endl;

				const VectorClass& psi = BaseClass::getRegisterStorage();
				for (size_t i = 0; i < BaseClass::NrBasisStates; ++i)
				{
					thefile << i << "\t";
					if (amplitude) thefile << std::abs(psi(i));
					else thefile << psi(i);
					thefile << std::endl;
				}
				return true;
			}
			catch (...) {};

			return false;
		}

		void displayState(size_t state) const  // :45-58
		{
			const size_t nQubits = BaseClass::getNrQubits();
			std::cout << "|";
			size_t mask = 1ULL << (nQubits - 1);
			for (size_t qubit = 0; qubit < nQubits; ++qubit)
			{
				std::cout << ((state & mask) ? "1" : "0");
				mask >>= 1;
			}
			std::cout << ">    ";
		}

		void displayRegister() const  // :60-90
		{
			const size_t nQubits = BaseClass::getNrQubits();
			const size_t nStates = BaseClass::getNrBasisStates();
			const VectorClass& psi = BaseClass::getRegisterStorage();

			std::cout << std::setprecision(4);
			for (size_t state = 0; state < nStates; ++state)
			{
				const std::complex<double> val = psi(state);
				if (abs(real(val)) < 1E-10 && abs(imag(val)) < 1E-10) continue;

				bool r = false;
				if (abs(real(val)) > 1E-10) {
					std::cout << real(val) << " ";
					r = true;
				}
				if (abs(imag(val)) > 1E-10) {
					if (r && imag(val) > 0) std::cout << "+ ";
					if (imag(val) < 0) {
						std::cout << "-";
						if (r) std::cout << " ";
					}
					std::cout << abs(imag(val)) << "i ";
				}
				displayState(nQubits);
			}
		}
	};

}
