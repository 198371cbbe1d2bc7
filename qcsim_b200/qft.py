"""QuantumFourierTransform / QubitsSwapper: mirrors of the reference sub-algorithms
(QuantumFourierTransform.h:9-93, QubitsSwapper.h:12-48) driving a QubitRegister gate by gate,
exactly as the reference does -- H, controlled phase shifts with the phase halved each step,
then the SWAP ladder.  `QubitRegister.QFT` is the single-call engine path for the same circuit.
"""
from __future__ import annotations

import math

from . import gates

INT_MAX = 2 ** 31 - 1


class QubitsSwapper:
    def __init__(self, N: int, startQubit: int = 0, endQubit: int = INT_MAX):
        # QuantumSubAlgorithmOnSubregister ctor (QuantumAlgorithm.h)
        self.sQubit = startQubit
        self.eQubit = max(startQubit, min(N - 1, endQubit))
        self.swapOp = gates.SwapGate()

    def getStartQubit(self) -> int:
        return self.sQubit

    def getEndQubit(self) -> int:
        return self.eQubit

    def Swap(self, reg) -> None:
        s, e = self.sQubit, self.eQubit
        while s < e:
            reg.ApplyGate(self.swapOp, s, e)
            s += 1
            e -= 1

    def Execute(self, reg) -> int:
        self.Swap(reg)
        return reg.MeasureAll()


class QuantumFourierTransform(QubitsSwapper):
    def __init__(self, N: int, startQubit: int = 0, endQubit: int = INT_MAX):
        super().__init__(N, startQubit, endQubit)
        self.hadamard = gates.HadamardGate()

    def Execute(self, reg) -> int:
        self.QFT(reg)
        return reg.MeasureAll()

    def QFT(self, reg, doSwap: bool = True) -> None:
        sq, eq = self.sQubit, self.eQubit
        reg.ApplyGate(self.hadamard, eq)
        for cur in range(eq, sq, -1):
            phase = math.pi / 2  # M_PI_2
            for ctrl in range(cur - 1, sq - 1, -1):
                reg.ApplyGate(gates.ControlledPhaseShiftGate(phase), cur, ctrl)
                phase *= 0.5
            reg.ApplyGate(self.hadamard, cur - 1)
        if doSwap:
            self.Swap(reg)

    def IQFT(self, reg, doSwap: bool = True) -> None:
        sq, eq = self.sQubit, self.eQubit
        if doSwap:
            self.Swap(reg)
        for cur in range(sq + 1, eq + 1):
            reg.ApplyGate(self.hadamard, cur - 1)
            phase = -math.pi / 2
            for ctrl in range(cur - 1, sq - 1, -1):
                reg.ApplyGate(gates.ControlledPhaseShiftGate(phase), cur, ctrl)
                phase *= 0.5
        reg.ApplyGate(self.hadamard, eq)
