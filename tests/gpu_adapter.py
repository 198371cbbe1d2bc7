"""Adapter giving the GPU register the same driving interface as the oracles, so one test body
runs both.  Everything goes through the C ABI via qcsim_b200.QubitRegister."""
import numpy as np

import qcsim_b200
from qcsim_b200 import gates


class GpuSim:
    kind = "gpu"

    def __init__(self, n, fusion=False, strict=False, batch=False):
        self.n = n
        self.dim = 1 << n
        self.reg = qcsim_b200.QubitRegister(n, seed=1)
        self.batch = batch
        if fusion:
            self.reg.set_fusion(True)
        if strict:
            self.reg.set_strict_measure(True)

    def close(self):
        self.reg.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def apply(self, gate, q, c1=0, c2=0):
        self.reg.ApplyGate(gate, q, c1, c2)

    def apply_matrix(self, nq, matrix, q, c1=0, c2=0):
        self.reg.ApplyGate(gates.AppliedGate(matrix), q, c1, c2)

    def apply_circuit(self, circuit):
        if self.batch:
            self.reg.ApplyGates(circuit)
        else:
            for g in circuit:
                self.reg.ApplyGate(*g)

    def state(self):
        return self.reg.getRegisterStorage()

    def set_state(self, v):
        self.reg.setRegisterStorageFastNoNormalize(np.asarray(v, dtype=np.complex128))

    def set_basis_state(self, s):
        self.reg.setToBasisState(s)

    def norm2(self):
        return self.reg.norm2()

    def qubit_probability(self, q):
        return self.reg.GetQubitProbability(q)

    def measure_all(self, prob):
        return self.reg.MeasureAll(prob)

    def measure(self, a, b, prob):
        return self.reg.Measure(a, b, prob)

    def measure_all_nocollapse(self, prob):
        return self.reg.MeasureNoCollapse(prob=prob)

    def measure_nocollapse(self, a, b, prob):
        return self.reg.MeasureNoCollapse(a, b, prob=prob)

    def qft(self, sq=0, eq=2 ** 31 - 1, do_swap=True, inverse=False):
        self.reg.QFT(sq, eq, do_swap, inverse)
