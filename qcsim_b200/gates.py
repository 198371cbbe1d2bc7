"""Gate library: the host-side mirror of QCSim's gate classes.

Each factory below has the same name, parameters and matrix as the reference class it cites
(file:line relative to /root/reference/QCSim/) and carries the same structural flags the
reference exposes as virtual methods (SimpleGates.h:27-60).  Matrices are produced with the
same libm calls in the same order so they are bit-identical to the reference's (checked by
tests/test_oracle.py against the compiled reference).

Matrix index convention: row/col bit0 = `qubit` (target), bit1 = `controllingQubit1`,
bit2 = `controllingQubit2` (QubitRegisterCalculator.h:427,749).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# the reference's virtual flags, packed as in include/qcsim_b200.h
CONTROLLED = 1
TWO_CONTROLS = 2
DIAGONAL = 4
ANTIDIAGONAL = 8
SWAP = 16
ISWAP = 32
ISWAPDAG = 64


@dataclass
class Gate:
    """A QuantumGateWithOp (SimpleGates.h:63-132): matrix + the flags that pick the kernel."""

    name: str
    matrix: np.ndarray  # (2^nq, 2^nq) complex128, row-major
    flags: int = 0
    gate_id: int = -1  # id shared with oracle/ref_driver.cpp::makeGate, -1 for ad-hoc matrices
    params: tuple = field(default_factory=tuple)

    @property
    def nq(self) -> int:  # getQubitsNumber(), SimpleGates.h:117
        return int(round(math.log2(self.matrix.shape[0])))

    def getRawOperatorMatrix(self) -> np.ndarray:
        return self.matrix

    def getQubitsNumber(self) -> int:
        return self.nq

    def isControlled(self) -> bool:
        return bool(self.flags & CONTROLLED)

    def isDiagonal(self) -> bool:
        return bool(self.flags & DIAGONAL)

    def isAntidiagonal(self) -> bool:
        return bool(self.flags & ANTIDIAGONAL)

    def isSwapGate(self) -> bool:
        return bool(self.flags & SWAP)

    def adjoint(self) -> "Gate":
        """Flag-less adjoint, what Uncompute builds (QubitRegister.h:584)."""
        return AppliedGate(self.matrix.conj().T.copy())


def _polar(theta: float) -> complex:  # std::polar(1., theta)
    return complex(math.cos(theta), math.sin(theta))


def _m(rows) -> np.ndarray:
    return np.array(rows, dtype=np.complex128)


def _controlled(block: np.ndarray, dim: int) -> np.ndarray:
    """Identity with `block` in the lower-right corner (TwoQubitsControlledGate, QuantumGate.h:94)."""
    out = np.eye(dim, dtype=np.complex128)
    k = block.shape[0]
    out[dim - k:, dim - k:] = block
    return out


def AppliedGate(matrix: np.ndarray) -> Gate:
    """Flag-less gate (SimpleGates.h:439): always classified from its matrix."""
    return Gate("applied", np.ascontiguousarray(matrix, dtype=np.complex128), 0)


# ---- one-qubit gates (SimpleGates.h:582-919) ------------------------------------------------
_S2 = 1.0 / math.sqrt(2.0)


def HadamardGate() -> Gate:  # :582
    return Gate("h", _m([[_S2, _S2], [_S2, -_S2]]), 0, 0)


def HyGate() -> Gate:  # :602
    return Gate("hy", _m([[_S2, complex(0, -_S2)], [complex(0, _S2), -_S2]]), 0, 1)


def SGate() -> Gate:  # :620
    return Gate("s", _m([[1, 0], [0, 1j]]), DIAGONAL, 2)


def SDGGate() -> Gate:  # :638
    return Gate("sdg", _m([[1, 0], [0, -1j]]), DIAGONAL, 3)


def TGate() -> Gate:  # :657
    return Gate("t", _m([[1, 0], [0, _polar(math.pi / 4.0)]]), DIAGONAL, 4)


def TDGGate() -> Gate:  # :675
    return Gate("tdg", _m([[1, 0], [0, _polar(-math.pi / 4.0)]]), DIAGONAL, 5)


def PhaseShiftGate(theta: float = 0.0) -> Gate:  # :693
    return Gate("p", _m([[1, 0], [0, _polar(theta)]]), DIAGONAL, 6, (theta,))


def PauliXGate() -> Gate:  # :717
    return Gate("x", _m([[0, 1], [1, 0]]), ANTIDIAGONAL, 7)


def PauliYGate() -> Gate:  # :735
    return Gate("y", _m([[0, -1j], [1j, 0]]), ANTIDIAGONAL, 8)


def PauliZGate() -> Gate:  # :754
    return Gate("z", _m([[1, 0], [0, -1]]), DIAGONAL, 9)


def SquareRootNOTGate() -> Gate:  # :773
    a, b = complex(0.5, 0.5), complex(0.5, -0.5)
    return Gate("sx", _m([[a, b], [b, a]]), 0, 10)


def SquareRootNOTDagGate() -> Gate:  # :788
    a, b = complex(0.5, -0.5), complex(0.5, 0.5)
    return Gate("sxdg", _m([[a, b], [b, a]]), 0, 11)


def SplitterGate() -> Gate:  # :803
    return Gate("splitter", _m([[_S2, complex(0, _S2)], [complex(0, _S2), _S2]]), 0, 12)


def _rx_block(theta: float) -> np.ndarray:  # :840-847
    t2 = theta * 0.5
    c, s = complex(math.cos(t2), 0), complex(0, -math.sin(t2))
    return _m([[c, s], [s, c]])


def _ry_block(theta: float) -> np.ndarray:  # :862-869
    t2 = theta * 0.5
    return _m([[complex(math.cos(t2), 0), complex(-math.sin(t2), 0)], [complex(math.sin(t2), 0), complex(math.cos(t2), 0)]])


def _rz_block(theta: float) -> np.ndarray:  # :884-888
    t2 = theta * 0.5
    return _m([[_polar(-t2), 0], [0, _polar(t2)]])


def _u_block(theta: float, phi: float, lam: float, gamma: float) -> np.ndarray:  # :907-914
    t2 = theta * 0.5
    c, s = math.cos(t2), math.sin(t2)

    def times(z: complex, x: float) -> complex:  # complex * double, component-wise
        return complex(z.real * x, z.imag * x)

    p01 = _polar(gamma + lam)
    return _m([
        [times(_polar(gamma), c), times(complex(-p01.real, -p01.imag), s)],
        [times(_polar(gamma + phi), s), times(_polar(gamma + phi + lam), c)],
    ])


def RxGate(theta: float = 0.0) -> Gate:  # :829
    return Gate("rx", _rx_block(theta), 0, 13, (theta,))


def RyGate(theta: float = 0.0) -> Gate:  # :851
    return Gate("ry", _ry_block(theta), 0, 14, (theta,))


def RzGate(theta: float = 0.0) -> Gate:  # :873
    return Gate("rz", _rz_block(theta), DIAGONAL, 15, (theta,))


def UGate(theta: float = 0.0, phi: float = 0.0, lam: float = 0.0, gamma: float = 0.0) -> Gate:  # :899
    return Gate("u", _u_block(theta, phi, lam, gamma), 0, 16, (theta, phi, lam, gamma))


# ---- two-qubit gates (QuantumGate.h:10-372) ---------------------------------------------------
def SwapGate() -> Gate:  # :10
    return Gate("swap", _m([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]), SWAP, 20)


def iSwapGate() -> Gate:  # :30
    return Gate("iswap", _m([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]), ISWAP, 21)


def iSwapDagGate() -> Gate:  # :50
    return Gate("iswapdg", _m([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, 1]]), ISWAPDAG, 22)


def DecrementGate() -> Gate:  # :70
    return Gate("dec", _m([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0]]), 0, 23)


def CNOTGate() -> Gate:  # :118
    return Gate("cx", _controlled(_m([[0, 1], [1, 0]]), 4), CONTROLLED | ANTIDIAGONAL, 24)


def ControlledYGate() -> Gate:  # :149
    return Gate("cy", _controlled(_m([[0, -1j], [1j, 0]]), 4), CONTROLLED | ANTIDIAGONAL, 25)


def ControlledZGate() -> Gate:  # :170
    return Gate("cz", _controlled(_m([[1, 0], [0, -1]]), 4), CONTROLLED | DIAGONAL, 26)


def ControlledHadamardGate() -> Gate:  # :187
    return Gate("ch", _controlled(_m([[_S2, _S2], [_S2, -_S2]]), 4), CONTROLLED, 27)


def ControlledSquareRootNOTGate() -> Gate:  # :203
    return Gate("csx", _controlled(SquareRootNOTGate().matrix, 4), CONTROLLED, 28)


def ControlledSquareRootNOTDagGate() -> Gate:  # :218
    return Gate("csxdg", _controlled(SquareRootNOTDagGate().matrix, 4), CONTROLLED, 29)


def ControlledPhaseGate() -> Gate:  # :234
    return Gate("cs", _controlled(_m([[1, 0], [0, 1j]]), 4), CONTROLLED | DIAGONAL, 30)


def ControlledPhaseShiftGate(theta: float = 0.0) -> Gate:  # :251
    return Gate("cp", _controlled(_m([[1, 0], [0, _polar(theta)]]), 4), CONTROLLED | DIAGONAL, 31, (theta,))


def ControlledUGate(theta: float = 0.0, phi: float = 0.0, lam: float = 0.0, gamma: float = 0.0) -> Gate:  # :273
    return Gate("cu", _controlled(_u_block(theta, phi, lam, gamma), 4), CONTROLLED, 32, (theta, phi, lam, gamma))


def ControlledRxGate(theta: float = 0.0) -> Gate:  # :305
    return Gate("crx", _controlled(_rx_block(theta), 4), CONTROLLED, 33, (theta,))


def ControlledRyGate(theta: float = 0.0) -> Gate:  # :327
    return Gate("cry", _controlled(_ry_block(theta), 4), CONTROLLED, 34, (theta,))


def ControlledRzGate(theta: float = 0.0) -> Gate:  # :349
    return Gate("crz", _controlled(_rz_block(theta), 4), CONTROLLED | DIAGONAL, 35, (theta,))


# ---- three-qubit gates (QuantumGate.h:376-469) -------------------------------------------------
def ToffoliGate() -> Gate:  # :402
    return Gate("ccx", _controlled(_m([[0, 1], [1, 0]]), 8), CONTROLLED | TWO_CONTROLS | ANTIDIAGONAL, 40)


def FredkinGate() -> Gate:  # :429  (isSwapGate is tested before isControlled, QubitRegisterCalculator.h:181)
    m = np.eye(8, dtype=np.complex128)
    m[5, 5] = m[6, 6] = 0
    m[5, 6] = m[6, 5] = 1
    return Gate("cswap", m, CONTROLLED | SWAP, 41)


def CCZGate() -> Gate:  # :449
    return Gate("ccz", _controlled(_m([[1, 0], [0, -1]]), 8), CONTROLLED | TWO_CONTROLS | DIAGONAL, 42)


ONE_QUBIT = [HadamardGate, HyGate, SGate, SDGGate, TGate, TDGGate, PauliXGate, PauliYGate, PauliZGate,
             SquareRootNOTGate, SquareRootNOTDagGate, SplitterGate]
ONE_QUBIT_PARAM = [PhaseShiftGate, RxGate, RyGate, RzGate]
TWO_QUBIT = [SwapGate, iSwapGate, iSwapDagGate, DecrementGate, CNOTGate, ControlledYGate, ControlledZGate,
             ControlledHadamardGate, ControlledSquareRootNOTGate, ControlledSquareRootNOTDagGate, ControlledPhaseGate]
TWO_QUBIT_PARAM = [ControlledPhaseShiftGate, ControlledRxGate, ControlledRyGate, ControlledRzGate]
THREE_QUBIT = [ToffoliGate, FredkinGate, CCZGate]


def all_gate_samples(rng=None):
    """One instance of every reference gate class (angles from rng or fixed), for sweeps."""
    import random

    r = rng or random.Random(1234)
    ang = lambda: r.uniform(-2 * math.pi, 2 * math.pi)
    out = [f() for f in ONE_QUBIT] + [f(ang()) for f in ONE_QUBIT_PARAM] + [UGate(ang(), ang(), ang(), ang())]
    out += [f() for f in TWO_QUBIT] + [f(ang()) for f in TWO_QUBIT_PARAM] + [ControlledUGate(ang(), ang(), ang(), ang())]
    out += [f() for f in THREE_QUBIT]
    return out
