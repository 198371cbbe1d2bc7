"""QubitRegister: host-side mirror of QC::QubitRegister (QubitRegister.h:11-728) on top of the C ABI.

Same method names, argument meaning and error behaviour as the reference class:
std::invalid_argument -> ValueError with the reference's message, silent no-ops where the
reference is silent.  The state lives on the GPU; this class only holds the handle, the RNG
(std::mt19937_64 restated in rng.py) and the recorded gates for Compute/Uncompute.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .gates import AppliedGate, Gate
from .rng import Mt19937_64, time_seed

_INVALID = {_lib.ERR_QUBIT_TOO_HIGH, _lib.ERR_CTRL_TOO_HIGH, _lib.ERR_SAME_QUBITS}


def _matrix_ptr(gate: Gate):
    m = np.ascontiguousarray(gate.matrix, dtype=np.complex128)
    return m, m.ctypes.data_as(C.c_void_p)


class QubitRegister:
    """QC::QubitRegister<> drop-in (Python host side)."""

    OneQubitOmpLimit = 8192  # QubitRegisterCalculator.h:1276, kept for source compatibility

    def __init__(self, N: int = 3, addseed: int = 0, device: int = 0, *, seed: Optional[int] = None,
                 _handle=None, max_host_qubits: int = 31, devices: Optional[Sequence[int]] = None):
        assert N > 0
        self._lib = _lib.load()
        self.NrQubits = int(N)
        self.NrBasisStates = 1 << self.NrQubits
        self._max_host_qubits = max_host_qubits
        if _handle is None:
            h = C.c_void_p()
            if devices is not None and len(devices) > 1:
                # one object, one caller, several GPUs of this process (qcsim_sv_create_multi)
                ids = (C.c_int * len(devices))(*devices)
                _lib.check(self._lib.qcsim_sv_create_multi(C.byref(h), self.NrQubits, len(devices), ids))
            else:
                _lib.check(self._lib.qcsim_sv_create(C.byref(h), self.NrQubits, devices[0] if devices else device))
            self._h = h
        else:
            self._h = _handle
        self.rng = Mt19937_64(seed if seed is not None else time_seed(addseed))
        self.recordGates = False
        self.computeGates: List[Tuple[Gate, int, int, int]] = []
        self._multithreading = True

    # -- lifetime ----------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.qcsim_sv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- sizes -------------------------------------------------------------------------------
    def getNrQubits(self) -> int:
        return self.NrQubits

    def getNrBasisStates(self) -> int:
        return self.NrBasisStates

    # -- state access (QubitRegister.h:62-130) -------------------------------------------------
    def getBasisStateAmplitude(self, State: int) -> complex:
        if State >= self.NrBasisStates or State < 0:
            return 0j
        out = (C.c_double * 2)()
        _lib.check(self._lib.qcsim_sv_get_amplitude(self._h, State, out))
        return complex(out[0], out[1])

    def getBasisStateProbability(self, State: int) -> float:
        a = self.getBasisStateAmplitude(State)
        return a.real * a.real + a.imag * a.imag

    def setToBasisState(self, State: int) -> None:
        if State >= self.NrBasisStates or State < 0:
            return
        _lib.check(self._lib.qcsim_sv_set_basis_state(self._h, State))

    def setToQubitState(self, q: int) -> None:
        if q >= self.NrQubits or q < 0:
            return
        self.setToBasisState(1 << q)

    def setToCatState(self) -> None:
        self.Clear()
        v = 1.0 / math.sqrt(2.0)
        self.setRawAmplitude(0, v)
        self.setRawAmplitude(self.NrBasisStates - 1, v)

    def Reset(self) -> None:
        self.setToBasisState(0)

    def setToEqualSuperposition(self) -> None:
        _lib.check(self._lib.qcsim_sv_fill(self._h, 1.0 / math.sqrt(self.NrBasisStates), 0.0))

    def setRawAmplitude(self, State: int, val: complex) -> None:
        if State >= self.NrBasisStates or State < 0:
            return
        val = complex(val)
        _lib.check(self._lib.qcsim_sv_set_amplitude(self._h, State, val.real, val.imag))

    def Clear(self) -> None:
        _lib.check(self._lib.qcsim_sv_fill(self._h, 0.0, 0.0))

    def Normalize(self) -> None:
        _lib.check(self._lib.qcsim_sv_normalize(self._h))

    def norm2(self) -> float:
        out = C.c_double()
        _lib.check(self._lib.qcsim_sv_norm2(self._h, C.byref(out)))
        return out.value

    # -- measurement (QubitRegister.h:169-224, 619-642, 695-713) ---------------------------------
    def _draw(self) -> float:
        return self.rng.draw()

    def MeasureAll(self, prob: Optional[float] = None) -> int:
        out = C.c_uint64()
        p = self._draw() if prob is None else prob
        _lib.check(self._lib.qcsim_sv_measure_all(self._h, p, C.byref(out)))
        return out.value

    def MeasureQubit(self, qubit: int, prob: Optional[float] = None) -> int:
        return self.Measure(qubit, qubit, prob)

    def Measure(self, firstQubit: int, secondQubit: int, prob: Optional[float] = None) -> int:
        out = C.c_uint64()
        p = self._draw() if prob is None else prob
        _lib.check(self._lib.qcsim_sv_measure(self._h, firstQubit, secondQubit, p, C.byref(out)))
        return out.value

    def MeasureNoCollapse(self, firstQubit: Optional[int] = None, secondQubit: Optional[int] = None,
                          prob: Optional[float] = None) -> int:
        out = C.c_uint64()
        p = self._draw() if prob is None else prob
        if firstQubit is None:
            _lib.check(self._lib.qcsim_sv_measure_all_nocollapse(self._h, p, C.byref(out)))
        else:
            last = firstQubit if secondQubit is None else secondQubit
            _lib.check(self._lib.qcsim_sv_measure_nocollapse(self._h, firstQubit, last, p, C.byref(out)))
        return out.value

    def RepeatedMeasure(self, *args, **kw) -> Dict[int, int]:
        """RepeatedMeasure(nrTimes=1000) or RepeatedMeasure(firstQubit, secondQubit, nrTimes=1000)
        (QubitRegister.h:227-274, 325-376)."""
        if len(args) >= 2:
            first, second = args[0], args[1]
            nr = args[2] if len(args) > 2 else kw.get("nrTimes", 1000)
        else:
            first = second = None
            nr = args[0] if args else kw.get("nrTimes", 1000)
        res: Dict[int, int] = {}
        if nr == 0:
            return res
        mask = None if first is None else ((1 << (second + 1)) - 1) - ((1 << first) - 1)
        if nr == 1:
            # the reference's single-shot shortcut (:231-236, 334-339): MeasureNoCollapse -- outcome 0 when the draw is beyond the
            # total, no table cut; the range variant masks and shifts the ALREADY shifted outcome once more (:337)
            if first is None:
                meas = self.MeasureNoCollapse()
            else:
                meas = (self.MeasureNoCollapse(first, second) & mask) >> first
            return {meas: 1}
        probs = np.array([self._draw() for _ in range(nr)], dtype=np.float64)
        outs = np.zeros(nr, dtype=np.uint64)
        _lib.check(self._lib.qcsim_sv_sample(self._h, probs.ctypes.data_as(C.c_void_p), nr,
                                             outs.ctypes.data_as(C.c_void_p)))
        if first is not None:
            outs = (outs & np.uint64(mask)) >> np.uint64(first)
        for v in outs.tolist():
            res[v] = res.get(v, 0) + 1
        return dict(sorted(res.items()))

    RepeatedMeasureUnordered = RepeatedMeasure

    def GetQubitProbability(self, qubit: int) -> float:
        out = C.c_double()
        _lib.check(self._lib.qcsim_sv_qubit_probability(self._h, qubit, C.byref(out)))
        return out.value

    # -- gates (QubitRegister.h:434-497) ------------------------------------------------------------
    def ApplyGate(self, gate: Gate, qubit: int, controllingQubit1: int = 0, controllingQubit2: int = 0) -> None:
        m, ptr = _matrix_ptr(gate)
        rc = self._lib.qcsim_sv_apply(self._h, gate.nq, ptr, gate.flags, qubit, controllingQubit1, controllingQubit2)
        if rc in _INVALID:
            raise ValueError(self._lib.qcsim_last_error().decode())  # std::invalid_argument
        _lib.check(rc)
        if self.recordGates:  # :484-485
            self.computeGates.append((AppliedGate(gate.matrix), qubit, controllingQubit1, controllingQubit2))

    def ApplyGates(self, gates: Iterable[Tuple[Gate, int, int, int]]) -> None:
        """ApplyGates(vector<AppliedGate>) (:493-497); executed as fused gate blocks."""
        gates = list(gates)
        if not gates:
            return
        arr = (_lib.GateStruct * len(gates))()
        for i, g in enumerate(gates):
            gate, q = g[0], g[1]
            c1 = g[2] if len(g) > 2 else 0
            c2 = g[3] if len(g) > 3 else 0
            arr[i].nq, arr[i].flags, arr[i].q, arr[i].c1, arr[i].c2 = gate.nq, gate.flags, q, c1, c2
            flat = np.ascontiguousarray(gate.matrix, dtype=np.complex128).view(np.float64).ravel()
            C.memmove(arr[i].m, flat.ctypes.data, flat.nbytes)
        rc = self._lib.qcsim_sv_apply_batch(self._h, arr, len(gates))
        if rc in _INVALID:
            raise ValueError(self._lib.qcsim_last_error().decode())
        _lib.check(rc)
        if self.recordGates:
            for g in gates:
                self.computeGates.append((AppliedGate(g[0].matrix), g[1], g[2] if len(g) > 2 else 0, g[3] if len(g) > 3 else 0))

    def ApplyOperatorMatrix(self, m: np.ndarray) -> None:
        """registerStorage = m * registerStorage (QubitRegister.h:499-505): dense 2^n x 2^n operator as a device GEMV;
        small registers only (the engine refuses above QCSIM_MAX_OPERATOR_QUBITS = 13).  Recorded like the reference does."""
        m = np.ascontiguousarray(m, dtype=np.complex128)
        if m.shape != (self.NrBasisStates, self.NrBasisStates):
            raise ValueError("operator matrix must be 2^n x 2^n")
        _lib.check(self._lib.qcsim_sv_apply_operator(self._h, m.ctypes.data_as(C.c_void_p)))
        if self.recordGates:
            self.computeGates.append((AppliedGate(m), 0, 0, 0))

    def ApplyCircuitFile(self, path: str) -> None:
        """Replay a recorded gate stream from a circuit file (circuits.save_circuit; include/qcsim_b200.h "circuit files")."""
        _lib.check(self._lib.qcsim_sv_apply_circuit_file(self._h, str(path).encode()))

    def QFT(self, sq: int = 0, eq: int = 2 ** 31 - 1, doSwap: bool = True, inverse: bool = False) -> None:
        """QuantumFourierTransform::QFT/IQFT as one engine call (QuantumFourierTransform.h:35-87)."""
        _lib.check(self._lib.qcsim_sv_qft(self._h, sq, eq, int(doSwap), int(inverse)))

    # -- storage (QubitRegister.h:507-524) ---------------------------------------------------------
    def getRegisterStorage(self) -> np.ndarray:
        if self.NrQubits > self._max_host_qubits:
            raise MemoryError(f"{self.NrQubits}-qubit state does not fit the host mirror; use download(first, count)")
        return self.download(0, self.NrBasisStates)

    def download(self, first: int, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.complex128)
        _lib.check(self._lib.qcsim_sv_download(self._h, out.ctypes.data_as(C.c_void_p), first, count))
        return out

    def upload(self, vals: np.ndarray, first: int = 0) -> None:
        v = np.ascontiguousarray(vals, dtype=np.complex128)
        _lib.check(self._lib.qcsim_sv_upload(self._h, v.ctypes.data_as(C.c_void_p), first, v.size))

    def setRegisterStorage(self, vals: np.ndarray) -> None:
        if len(vals) != self.NrBasisStates:
            return
        self.upload(vals)
        self.Normalize()

    def setRegisterStorageFastNoNormalize(self, vals: np.ndarray) -> None:
        self.upload(vals)

    def stateFidelity(self, state: np.ndarray) -> float:
        if len(state) != self.NrBasisStates:
            return 0.0
        p = np.vdot(self.getRegisterStorage(), state)
        return p.real * p.real + p.imag * p.imag

    # -- gate recording (QubitRegister.h:536-590) -------------------------------------------------
    def ComputeStart(self) -> None:
        self.recordGates = True
        self.computeGates = []

    def ComputeEnd(self) -> None:
        self.recordGates = False

    def ComputeClear(self) -> None:
        self.computeGates = []

    def _replay(self, recorded) -> None:
        """QubitRegister.h:554-590: recorded gates on more than three qubits go through ApplyOperatorMatrix (:563, 581)"""
        run = []
        for g, q, c1, c2 in recorded:
            if g.nq > 3:
                self.ApplyGates(run)
                run = []
                self.ApplyOperatorMatrix(g.matrix)
            else:
                run.append((g, q, c1, c2))
        self.ApplyGates(run)

    def Compute(self) -> None:
        save, self.recordGates = self.recordGates, False
        self._replay(self.computeGates)
        self.recordGates = save

    def Uncompute(self) -> None:
        save, self.recordGates = self.recordGates, False
        self._replay([(g.adjoint(), q, c1, c2) for (g, q, c1, c2) in reversed(self.computeGates)])
        self.recordGates = save

    # -- save / restore / clone (QubitRegister.h:600-616, 662-674) ------------------------------------
    def SaveState(self) -> None:
        _lib.check(self._lib.qcsim_sv_save_state(self._h))

    def RestoreState(self) -> None:
        _lib.check(self._lib.qcsim_sv_restore_state(self._h, 0))

    def RestoreStateDestructive(self) -> None:
        _lib.check(self._lib.qcsim_sv_restore_state(self._h, 1))

    def Clone(self) -> "QubitRegister":
        h = C.c_void_p()
        _lib.check(self._lib.qcsim_sv_clone(self._h, C.byref(h)))
        r = QubitRegister(self.NrQubits, _handle=h, max_host_qubits=self._max_host_qubits)
        r.computeGates = list(self.computeGates)
        r.recordGates = self.recordGates
        return r

    def ExpectationValue(self, gates: Sequence[Tuple[Gate, int, int, int]]) -> complex:
        """<psi| G_k ... G_1 |psi> (QubitRegister.h:646-660)."""
        if not gates:
            return 1.0 + 0j
        work = self.Clone()
        try:
            work.recordGates = False
            work.ApplyGates(gates)
            out = (C.c_double * 2)()
            _lib.check(self._lib.qcsim_sv_inner_product(self._h, work._h, out))
            return complex(out[0], out[1])
        finally:
            work.close()

    # -- engine controls -----------------------------------------------------------------------------
    def SetMultithreading(self, enable: bool = True) -> None:  # QubitRegisterCalculator.h:1263 (no-op here)
        self._multithreading = bool(enable)

    def GetMultithreading(self) -> bool:
        return self._multithreading

    def set_fusion(self, enabled: bool) -> None:
        _lib.check(self._lib.qcsim_sv_set_fusion(self._h, int(enabled)))

    def set_strict_measure(self, enabled: bool) -> None:
        _lib.check(self._lib.qcsim_sv_set_strict_measure(self._h, int(enabled)))

    def flush(self) -> None:
        _lib.check(self._lib.qcsim_sv_flush(self._h))

    def sync(self) -> None:
        _lib.check(self._lib.qcsim_sv_sync(self._h))

    def stats(self) -> dict:
        s = _lib.Stats()
        _lib.check(self._lib.qcsim_sv_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self) -> None:
        _lib.check(self._lib.qcsim_sv_reset_stats(self._h))
