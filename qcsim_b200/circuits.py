"""Deterministic workload generators (BASELINE.json configs).

A circuit is a list of (Gate, q, c1, c2) tuples -- the same thing the reference records as
AppliedGate (SimpleGates.h:439) -- so one list drives the GPU engine and both oracles.
Randomness comes from a self-contained splitmix64 so circuits are bit-identical everywhere.
"""
from __future__ import annotations

import math

import numpy as np
from typing import List, Tuple

from . import gates
from .gates import Gate

Circuit = List[Tuple[Gate, int, int, int]]

_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self) -> float:  # [0, 1) with 53 bits
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def below(self, n: int) -> int:
        return self.next() % n

    def shuffle(self, xs: list) -> None:  # Fisher-Yates
        for i in range(len(xs) - 1, 0, -1):
            j = self.below(i + 1)
            xs[i], xs[j] = xs[j], xs[i]


RANDOM_CIRCUIT_SEED = 20260117


def random_layer(n: int, rng: SplitMix64) -> Circuit:
    """One layer of BASELINE config 2: a 1-qubit gate on every qubit, kind uniform in
    {H, RX(t), RZ(t)}, t uniform in [-2pi, 2pi); then a random permutation of the qubits split
    into floor(0.3 n) CNOT (target, control) pairs and floor(n / 7.5) CCNOT triples
    (n = 30: 9 CNOT + 4 CCNOT -> 43 gate applications per layer)."""
    layer: Circuit = []
    for q in range(n):
        kind = rng.below(3)
        theta = (rng.uniform() * 4.0 - 2.0) * math.pi
        if kind == 0:
            layer.append((gates.HadamardGate(), q, 0, 0))
        elif kind == 1:
            layer.append((gates.RxGate(theta), q, 0, 0))
        else:
            layer.append((gates.RzGate(theta), q, 0, 0))
    perm = list(range(n))
    rng.shuffle(perm)
    n_cx = (3 * n) // 10
    n_ccx = (2 * n) // 15
    while 2 * n_cx + 3 * n_ccx > n and n_ccx > 0:
        n_ccx -= 1
    while 2 * n_cx + 3 * n_ccx > n and n_cx > 0:
        n_cx -= 1
    pos = 0
    for _ in range(n_cx):
        layer.append((gates.CNOTGate(), perm[pos], perm[pos + 1], 0))
        pos += 2
    for _ in range(n_ccx):
        layer.append((gates.ToffoliGate(), perm[pos], perm[pos + 1], perm[pos + 2]))
        pos += 3
    return layer


def random_circuit(n: int, layers: int, seed: int = RANDOM_CIRCUIT_SEED) -> Circuit:
    rng = SplitMix64(seed)
    out: Circuit = []
    for _ in range(layers):
        out.extend(random_layer(n, rng))
    return out


def gates_per_layer(n: int) -> int:
    return len(random_layer(n, SplitMix64(1)))


def qft_circuit(n: int, sq: int = 0, eq: int = None, do_swap: bool = True, inverse: bool = False) -> Circuit:
    """The gate list QuantumFourierTransform::QFT / IQFT issues (QuantumFourierTransform.h:35-87)."""
    eq = n - 1 if eq is None else max(sq, min(n - 1, eq))
    h = gates.HadamardGate()
    sw = gates.SwapGate()
    out: Circuit = []

    def swaps():
        s, e = sq, eq
        while s < e:
            out.append((sw, s, e, 0))
            s += 1
            e -= 1

    if not inverse:
        out.append((h, eq, 0, 0))
        for cur in range(eq, sq, -1):
            phase = math.pi / 2
            for ctrl in range(cur - 1, sq - 1, -1):
                out.append((gates.ControlledPhaseShiftGate(phase), cur, ctrl, 0))
                phase *= 0.5
            out.append((h, cur - 1, 0, 0))
        if do_swap:
            swaps()
    else:
        if do_swap:
            swaps()
        for cur in range(sq + 1, eq + 1):
            out.append((h, cur - 1, 0, 0))
            phase = -math.pi / 2
            for ctrl in range(cur - 1, sq - 1, -1):
                out.append((gates.ControlledPhaseShiftGate(phase), cur, ctrl, 0))
                phase *= 0.5
        out.append((h, eq, 0, 0))
    return out


def ncnot_circuit(controls: List[int], target: int, start_ancilla: int, clear_ancilla: bool = True) -> Circuit:
    """NControlledNotWithAncilla::Execute as a flat gate list (NControlledNotWithAncilla.h:24-96).
    The uncompute half is flag-less adjoints, exactly what QubitRegister::Uncompute replays
    (QubitRegister.h:573-590)."""
    cnot, ccnot = gates.CNOTGate(), gates.ToffoliGate()
    out: Circuit = []
    if not controls:
        return out
    if len(controls) == 2:
        return [(ccnot, target, controls[0], controls[1])]
    if len(controls) == 1:
        return [(cnot, target, controls[0], 0)]
    recorded: Circuit = []
    free = start_ancilla
    for i in range(0, len(controls) - 1, 2):
        recorded.append((ccnot, free, controls[i], controls[i + 1]))
        free += 1
    cur = start_ancilla
    final = None
    if len(controls) % 2:
        if free == start_ancilla + 1:
            final = (ccnot, target, controls[-1], cur)
        else:
            recorded.append((ccnot, free, controls[-1], cur))
            free += 1
            cur += 1
    if final is None:
        while free - cur > 2:
            recorded.append((ccnot, free, cur, cur + 1))
            free += 1
            cur += 2
        final = (ccnot, target, cur, cur + 1) if free - cur == 2 else (cnot, target, cur, 0)
    out.extend(recorded)
    out.append(final)
    if clear_ancilla:
        out.extend((g.adjoint(), q, c1, c2) for (g, q, c1, c2) in reversed(recorded))
    return out


def grover_iterations(n_search: int) -> int:
    return int(round(math.pi / 4.0 * math.sqrt(1 << n_search)))  # GroverAlgorithm.h:187


def grover_gates_circuit(n_search: int, marked: int, iterations: int = None, qubit_map=None) -> Circuit:
    """GroverAlgorithmWithGatesOracle::ExecuteWithoutMeasurement (GroverAlgorithm.h:128-242):
    search qubits 0..N-1, oracle target N, ancillas N+1..2N-2.  `qubit_map` relabels qubits
    (SURVEY 8d config 4 puts the search qubits on top so the H walls hit the global qubits)."""
    N = n_search
    its = grover_iterations(N) if iterations is None else iterations
    h, x = gates.HadamardGate(), gates.PauliXGate()
    nc = ncnot_circuit(list(range(N)), N, N + 1, True)
    out: Circuit = []

    def hall():
        for q in range(N):
            out.append((h, q, 0, 0))

    def oracle(state):
        out.append((h, N, 0, 0))
        for q in range(N):
            if not (state >> q) & 1:
                out.append((x, q, 0, 0))
        out.extend(nc)
        for q in range(N):
            if not (state >> q) & 1:
                out.append((x, q, 0, 0))
        out.append((h, N, 0, 0))

    hall()
    out.append((x, N, 0, 0))
    for _ in range(its):
        oracle(marked)
        hall()
        oracle(0)
        hall()
    if qubit_map is not None:
        def relabel(g, q, c1, c2):
            return (g, qubit_map[q], qubit_map[c1] if g.nq >= 2 else 0, qubit_map[c2] if g.nq >= 3 else 0)

        out = [relabel(*t) for t in out]
    return out


# ---- circuit files (include/qcsim_b200.h "circuit files") ---------------------------------------------------
def save_circuit(path: str, n: int, circuit: Circuit) -> None:
    """Write `circuit` (a list of (gate, q, c1, c2)) for an n-qubit register as a circuit file: the form both the engine
    (qcsim_sv_apply_circuit_file) and the compiled reference (oracle/ref_driver.cpp: ref_apply_circuit_file) replay."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    arr = (_lib.CircuitGateStruct * max(len(circuit), 1))()
    for i, (g, q, c1, c2) in enumerate(circuit):
        r = arr[i]
        r.nq, r.flags, r.gate_id, r.reserved, r.q, r.c1, r.c2 = g.nq, g.flags, g.gate_id, 0, q, c1, c2
        for j, v in enumerate((list(g.params) + [0.0] * 4)[:4]):
            r.params[j] = float(v)
        flat = np.ascontiguousarray(g.matrix, dtype=np.complex128).view(np.float64).ravel()
        C.memmove(r.m, flat.ctypes.data, flat.nbytes)
    _lib.check(lib.qcsim_circuit_save(str(path).encode(), n, arr, len(circuit)))


def load_circuit(path: str):
    """-> (n_qubits, [(Gate, q, c1, c2)]) read back through the C ABI loader"""
    import ctypes as C

    from . import _lib
    from .gates import Gate

    lib = _lib.load()
    n = C.c_uint32()
    ptr = C.c_void_p()
    count = C.c_uint64()
    _lib.check(lib.qcsim_circuit_load(str(path).encode(), C.byref(n), C.byref(ptr), C.byref(count)))
    out = []
    try:
        recs = C.cast(ptr, C.POINTER(_lib.CircuitGateStruct))
        for i in range(count.value):
            r = recs[i]
            d = 1 << r.nq
            m = np.frombuffer(r.m, dtype=np.complex128, count=d * d).reshape(d, d).copy()
            out.append((Gate(f"file{i}", m, r.flags, r.gate_id, tuple(r.params)), int(r.q), int(r.c1), int(r.c2)))
    finally:
        lib.qcsim_circuit_free(ptr)
    return n.value, out
