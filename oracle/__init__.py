"""Parity checkers for qcsim_b200.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; nothing under qcsim_b200/ does.

Two checkers with one interface:

* ``RefOracle``  -- the reference's own headers compiled (unmodified) into
  oracle/_ref/libqcsim_ref_{sse2,avx2}.so by oracle/Makefile.  Built in the dev container where
  /root/reference exists; the binaries travel to the GPU box.
* ``PortOracle`` -- oracle/qcsim_oracle.c, our C restatement, buildable anywhere gcc is.

``best_oracle()`` returns the compiled reference when its binary is present and loads, else the port.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PORT_LIB = os.path.join(HERE, "libqcsim_oracle.so")
REFERENCE_SRC = "/root/reference/QCSim"


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "qcsim_oracle.c")
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)
    return PORT_LIB


def build_ref(force: bool = False) -> bool:
    """Compile the reference where its sources exist; returns False (and does nothing) elsewhere."""
    if not os.path.isdir(REFERENCE_SRC):
        return False
    src = os.path.join(HERE, "ref_driver.cpp")
    libs = [os.path.join(REF_DIR, f"libqcsim_ref_{v}.so") for v in ("sse2", "avx2")]
    stale = force or any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in libs)
    if stale:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    return True


def ref_available(variant: str = "sse2") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libqcsim_ref_{variant}.so"))


_U64 = C.c_uint64
_VP = C.c_void_p


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


class _Base:
    kind = "?"

    def apply(self, gate, q: int, c1: int = 0, c2: int = 0) -> None:
        """reg.ApplyGate(<gate with its reference flags>, q, c1, c2)"""
        raise NotImplementedError

    def apply_circuit(self, circuit) -> None:
        for g in circuit:
            self.apply(*g)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class RefOracle(_Base):
    """The compiled reference (kind == "reference")."""

    kind = "reference"

    def __init__(self, n: int, variant: str = "sse2"):
        path = os.path.join(REF_DIR, f"libqcsim_ref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.ref_create.restype = _VP
        L.ref_create.argtypes = [C.c_int]
        L.ref_destroy.argtypes = [_VP]
        L.ref_last_error.restype = C.c_char_p
        L.ref_apply_named.argtypes = [_VP, C.c_int, _VP, _U64, _U64, _U64]
        L.ref_apply_named_dense.argtypes = [_VP, C.c_int, _VP, _U64, _U64, _U64]
        L.ref_apply_matrix.argtypes = [_VP, C.c_int, _VP, _U64, _U64, _U64]
        L.ref_gate_matrix.argtypes = [C.c_int, _VP, _VP, C.POINTER(C.c_int)]
        L.ref_get_state.argtypes = [_VP, _VP]
        L.ref_set_state.argtypes = [_VP, _VP]
        L.ref_set_basis_state.argtypes = [_VP, _U64]
        L.ref_set_equal_superposition.argtypes = [_VP]
        L.ref_set_cat_state.argtypes = [_VP]
        L.ref_normalize.argtypes = [_VP]
        L.ref_set_multithreading.argtypes = [_VP, C.c_int]
        L.ref_norm2.restype = C.c_double
        L.ref_norm2.argtypes = [_VP]
        L.ref_qubit_probability.restype = C.c_double
        L.ref_qubit_probability.argtypes = [_VP, _U64]
        for name in ("ref_measure_all", "ref_measure_all_nocollapse"):
            getattr(L, name).restype = _U64
            getattr(L, name).argtypes = [_VP, C.c_double]
        for name in ("ref_measure", "ref_measure_nocollapse"):
            getattr(L, name).restype = _U64
            getattr(L, name).argtypes = [_VP, _U64, _U64, C.c_double]
        L.ref_qft.argtypes = [_VP, _U64, _U64, C.c_int, C.c_int]
        for name in ("ref_compute_start", "ref_compute_end", "ref_compute", "ref_uncompute"):
            getattr(L, name).argtypes = [_VP]
        L.ref_ncnot.argtypes = [_VP, _VP, C.c_int, _U64, _U64, C.c_int]
        L.ref_grover_gates.argtypes = [C.c_int, _U64, _VP, C.POINTER(_U64)]
        L.ref_draper_add.argtypes = [C.c_int, _U64, _U64, _VP]
        L.ref_draws.argtypes = [_VP, _U64, C.c_int, _VP]
        L.ref_repeated_measure.restype = C.c_int
        L.ref_repeated_measure.argtypes = [_VP, _U64, _U64, C.c_int, _U64, _U64, _VP, _VP, C.c_int]
        L.ref_state_fidelity.restype = C.c_double
        L.ref_state_fidelity.argtypes = [_VP, _VP, _U64]
        L.ref_expectation_value.argtypes = [_VP, C.c_int, _VP, _VP, _VP, _VP]
        L.ref_save_state.argtypes = [_VP]
        L.ref_restore_state.argtypes = [_VP, C.c_int]
        L.ref_apply_operator_matrix.argtypes = [_VP, _VP]
        L.ref_apply_circuit_file.restype = C.c_long
        L.ref_apply_circuit_file.argtypes = [_VP, C.c_char_p]
        self.L = L
        self.n = n
        self.dim = 1 << n
        self.h = L.ref_create(n)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def _rc(self, rc):
        if rc == -1:
            raise ValueError(self.L.ref_last_error().decode())
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def num_threads(self) -> int:
        return self.L.ref_num_threads()

    def set_num_threads(self, n: int) -> None:
        self.L.ref_set_num_threads(n)

    def set_multithreading(self, on: bool) -> None:
        self.L.ref_set_multithreading(self.h, int(on))

    def apply(self, gate, q, c1=0, c2=0):
        if gate.gate_id >= 0:
            p = (C.c_double * 4)(*(list(gate.params) + [0.0] * 4)[:4])
            self._rc(self.L.ref_apply_named(self.h, gate.gate_id, p, q, c1, c2))
        else:
            self.apply_matrix(gate.nq, gate.matrix, q, c1, c2)

    def apply_dense(self, gate, q, c1=0, c2=0):
        """The reference's second, independent path: full 2^n x 2^n operator."""
        p = (C.c_double * 4)(*(list(gate.params) + [0.0] * 4)[:4])
        self._rc(self.L.ref_apply_named_dense(self.h, gate.gate_id, p, q, c1, c2))

    def apply_matrix(self, nq, matrix, q, c1=0, c2=0):
        m = _as_f64(matrix)
        self._rc(self.L.ref_apply_matrix(self.h, nq, m.ctypes.data_as(_VP), q, c1, c2))

    def gate_matrix(self, gate_id: int, params=()) -> np.ndarray:
        p = (C.c_double * 4)(*(list(params) + [0.0] * 4)[:4])
        out = np.zeros(64, dtype=np.complex128)
        nq = C.c_int()
        if self.L.ref_gate_matrix(gate_id, p, out.ctypes.data_as(_VP), C.byref(nq)) != 0:
            raise KeyError(gate_id)
        d = 1 << nq.value
        return out[: d * d].reshape(d, d).copy()

    def gate_flags(self, gate_id: int) -> int:
        return self.L.ref_gate_flags(gate_id)

    def state(self) -> np.ndarray:
        out = np.empty(self.dim, dtype=np.complex128)
        self.L.ref_get_state(self.h, out.ctypes.data_as(_VP))
        return out

    def set_state(self, v) -> None:
        v = _as_f64(v)
        assert v.size == self.dim
        self.L.ref_set_state(self.h, v.ctypes.data_as(_VP))

    def set_basis_state(self, s: int) -> None:
        self.L.ref_set_basis_state(self.h, s)

    def normalize(self) -> None:
        self.L.ref_normalize(self.h)

    def norm2(self) -> float:
        return self.L.ref_norm2(self.h)

    def qubit_probability(self, q: int) -> float:
        return self.L.ref_qubit_probability(self.h, q)

    def measure_all(self, prob: float) -> int:
        return self.L.ref_measure_all(self.h, prob)

    def measure(self, first: int, last: int, prob: float) -> int:
        return self.L.ref_measure(self.h, first, last, prob)

    def measure_all_nocollapse(self, prob: float) -> int:
        return self.L.ref_measure_all_nocollapse(self.h, prob)

    def measure_nocollapse(self, first: int, last: int, prob: float) -> int:
        return self.L.ref_measure_nocollapse(self.h, first, last, prob)

    def qft(self, sq=0, eq=2 ** 31 - 1, do_swap=True, inverse=False) -> None:
        self._rc(self.L.ref_qft(self.h, sq, eq, int(do_swap), int(inverse)))

    def compute_start(self):
        self.L.ref_compute_start(self.h)

    def compute_end(self):
        self.L.ref_compute_end(self.h)

    def compute(self):
        self.L.ref_compute(self.h)

    def uncompute(self):
        self.L.ref_uncompute(self.h)

    def ncnot(self, controls, target, start_ancilla, clear_ancilla=True):
        c = (C.c_uint64 * len(controls))(*controls)
        self._rc(self.L.ref_ncnot(self.h, c, len(controls), target, start_ancilla, int(clear_ancilla)))

    def draws(self, seed: int, count: int) -> np.ndarray:
        """`1. - uniformZeroOne(rng)` x count after rng.seed(seed) -- pins qcsim_b200.rng."""
        out = np.empty(count, dtype=np.float64)
        self.L.ref_draws(self.h, seed, count, out.ctypes.data_as(_VP))
        return out

    def repeated_measure(self, seed: int, nr_times: int, first=None, last=None, unordered=False) -> dict:
        """RepeatedMeasure[Unordered](nrTimes) / (first, last, nrTimes) after rng.seed(seed) (QubitRegister.h:227-429)"""
        cap = max(16, min(nr_times, 1 << 20) + 1)
        keys = np.zeros(cap, dtype=np.uint64)
        counts = np.zeros(cap, dtype=np.uint64)
        f, l = (1, 0) if first is None else (first, last)
        n = self.L.ref_repeated_measure(self.h, seed, nr_times, int(unordered), f, l, keys.ctypes.data_as(_VP), counts.ctypes.data_as(_VP), cap)
        assert n <= cap
        return {int(k): int(c) for k, c in zip(keys[:n], counts[:n])}

    def state_fidelity(self, state) -> float:
        v = _as_f64(state)
        return self.L.ref_state_fidelity(self.h, v.ctypes.data_as(_VP), v.size)

    def expectation_value(self, circuit) -> complex:
        """ExpectationValue(gates) with gates given as (gate, q, c1, c2) tuples (flag-less AppliedGates, :646-660)"""
        nq = (C.c_int * len(circuit))(*[g.nq for g, *_ in circuit])
        mats = np.concatenate([np.ascontiguousarray(g.matrix, dtype=np.complex128).ravel() for g, *_ in circuit]).view(np.float64)
        qs = np.array([[q, c1, c2] for _, q, c1, c2 in circuit], dtype=np.uint64).ravel()
        out = np.zeros(2)
        self._rc(self.L.ref_expectation_value(self.h, len(circuit), nq, mats.ctypes.data_as(_VP), qs.ctypes.data_as(_VP), out.ctypes.data_as(_VP)))
        return complex(out[0], out[1])

    def save_state(self):
        self.L.ref_save_state(self.h)

    def restore_state(self, destructive=False):
        self.L.ref_restore_state(self.h, int(destructive))

    def apply_operator_matrix(self, m) -> None:
        m = np.ascontiguousarray(m, dtype=np.complex128)
        assert m.shape == (self.dim, self.dim)
        self._rc(self.L.ref_apply_operator_matrix(self.h, m.ctypes.data_as(_VP)))

    def apply_circuit_file(self, path: str) -> int:
        """replay a circuit file written by qcsim_b200.circuits.save_circuit; returns the number of gates applied"""
        n = self.L.ref_apply_circuit_file(self.h, str(path).encode())
        if n < 0:
            self._rc(int(n))
        return int(n)

    def grover_gates(self, n_search: int, marked: int) -> np.ndarray:
        nq = 2 * n_search - 1
        out = np.empty(1 << nq, dtype=np.complex128)
        nqo = C.c_uint64()
        self._rc(self.L.ref_grover_gates(n_search, marked, out.ctypes.data_as(_VP), C.byref(nqo)))
        assert nqo.value == nq
        return out

    def draper_add(self, n_bits: int, n1: int, n2: int) -> np.ndarray:
        out = np.empty(1 << (2 * n_bits), dtype=np.complex128)
        self._rc(self.L.ref_draper_add(n_bits, n1, n2, out.ctypes.data_as(_VP)))
        return out


class PortOracle(_Base):
    """oracle/qcsim_oracle.c (kind == "port")."""

    kind = "port"

    def __init__(self, n: int):
        L = C.CDLL(build_port())
        L.orc_create.restype = _VP
        L.orc_create.argtypes = [C.c_int]
        L.orc_destroy.argtypes = [_VP]
        L.orc_get_state.argtypes = [_VP, _VP]
        L.orc_set_state.argtypes = [_VP, _VP]
        L.orc_set_basis_state.argtypes = [_VP, _U64]
        L.orc_apply.argtypes = [_VP, C.c_int, _VP, C.c_int, _U64, _U64, _U64]
        L.orc_norm2.restype = C.c_double
        L.orc_norm2.argtypes = [_VP]
        L.orc_qubit_probability.restype = C.c_double
        L.orc_qubit_probability.argtypes = [_VP, _U64]
        for name in ("orc_measure_all", "orc_measure_all_nocollapse"):
            getattr(L, name).restype = _U64
            getattr(L, name).argtypes = [_VP, C.c_double]
        for name in ("orc_measure", "orc_measure_nocollapse"):
            getattr(L, name).restype = _U64
            getattr(L, name).argtypes = [_VP, _U64, _U64, C.c_double]
        L.orc_qft.argtypes = [_VP, _U64, _U64, C.c_int, C.c_int]
        self.L = L
        self.n = n
        self.dim = 1 << n
        self.h = L.orc_create(n)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def num_threads(self) -> int:
        return os.cpu_count() or 1

    def apply(self, gate, q, c1=0, c2=0):
        self._apply(gate.nq, gate.matrix, gate.flags, q, c1, c2)

    def apply_matrix(self, nq, matrix, q, c1=0, c2=0):
        self._apply(nq, matrix, 0, q, c1, c2)

    def _apply(self, nq, matrix, flags, q, c1, c2):
        m = _as_f64(matrix)
        rc = self.L.orc_apply(self.h, nq, m.ctypes.data_as(_VP), flags, q, c1, c2)
        if rc in (-1, -2, -3):
            raise ValueError({-1: "Qubit number is too high", -2: "Controlling qubit number is too high",
                              -3: "Qubits must be different"}[rc])
        if rc != 0:
            raise RuntimeError(f"orc_apply rc={rc}")

    def state(self) -> np.ndarray:
        out = np.empty(self.dim, dtype=np.complex128)
        self.L.orc_get_state(self.h, out.ctypes.data_as(_VP))
        return out

    def set_state(self, v) -> None:
        v = _as_f64(v)
        assert v.size == self.dim
        self.L.orc_set_state(self.h, v.ctypes.data_as(_VP))

    def set_basis_state(self, s: int) -> None:
        self.L.orc_set_basis_state(self.h, s)

    def norm2(self) -> float:
        return self.L.orc_norm2(self.h)

    def qubit_probability(self, q: int) -> float:
        return self.L.orc_qubit_probability(self.h, q)

    def measure_all(self, prob: float) -> int:
        return self.L.orc_measure_all(self.h, prob)

    def measure(self, first: int, last: int, prob: float) -> int:
        return self.L.orc_measure(self.h, first, last, prob)

    def measure_all_nocollapse(self, prob: float) -> int:
        return self.L.orc_measure_all_nocollapse(self.h, prob)

    def measure_nocollapse(self, first: int, last: int, prob: float) -> int:
        return self.L.orc_measure_nocollapse(self.h, first, last, prob)

    def qft(self, sq=0, eq=2 ** 31 - 1, do_swap=True, inverse=False) -> None:
        self.L.orc_qft(self.h, sq, eq, int(do_swap), int(inverse))


def best_oracle(n: int, variant: str = "sse2") -> _Base:
    """The compiled reference when it is available, else the C port.  With QCSIM_REQUIRE_REF=1 (set by the GPU
    parity tests) there is no silent fallback: a missing reference build is an error."""
    if ref_available(variant):
        try:
            return RefOracle(n, variant)
        except OSError:
            if os.environ.get("QCSIM_REQUIRE_REF"):
                raise
    if os.environ.get("QCSIM_REQUIRE_REF"):
        raise RuntimeError("oracle/_ref (the compiled reference) is required for the GPU parity tests but is not available; "
                           "build it where /root/reference exists (python -c 'import oracle; oracle.build_ref()')")
    return PortOracle(n)
