"""qcsim_b200 -- B200-native statevector engine for QCSim's hot path.

Python host side: `QubitRegister` mirrors QC::QubitRegister, `gates` mirrors the gate classes,
`QuantumFourierTransform` mirrors the QFT sub-algorithm.  All compute happens in hand-written
sm_100a CUDA kernels behind the C ABI in include/qcsim_b200.h (libqcsim_b200.so); there is no
CPU fallback.
"""
from . import gates  # noqa: F401
from ._lib import QcsimError, load  # noqa: F401
from .qft import QuantumFourierTransform, QubitsSwapper  # noqa: F401
from .register import QubitRegister  # noqa: F401

__all__ = ["QubitRegister", "QuantumFourierTransform", "QubitsSwapper", "gates", "QcsimError", "load"]
