"""CPU tests: pin the oracle.

The C restatement (oracle/qcsim_oracle.c) is checked bit-for-bit (up to the sign of zero)
against the reference's own headers compiled into oracle/_ref, for every gate class and every
kernel branch, QFT/IQFT and all measurement entry points; the Python gate library and RNG are
checked against the reference's; and the committed golden fixtures are re-verified.
"""
import itertools
import math
import os

import numpy as np
import pytest

import oracle
from conftest import draws, random_state
from qcsim_b200 import circuits, gates
from qcsim_b200.rng import Mt19937_64

needs_ref = pytest.mark.skipif(not (oracle.build_ref() or oracle.ref_available()),
                               reason="compiled reference (oracle/_ref) not available")


def same(a, b):
    """bit-exact up to the sign of zero"""
    return np.array_equal(a, b)


@needs_ref
def test_gate_library_matches_reference_classes():
    with oracle.RefOracle(1) as ref:
        for g in gates.all_gate_samples():
            m = ref.gate_matrix(g.gate_id, g.params)
            assert m.shape == g.matrix.shape, g.name
            assert same(m, g.matrix), (g.name, m, g.matrix)
            assert ref.gate_flags(g.gate_id) == g.flags, g.name


@needs_ref
def test_rng_matches_libstdcxx():
    with oracle.RefOracle(1) as ref:
        for seed in (0, 1, 42, 2 ** 63 + 12345):
            want = ref.draws(seed, 700)  # crosses one 312-word refill twice
            r = Mt19937_64(seed)
            got = np.array([r.draw() for _ in range(700)])
            assert np.array_equal(want, got)


def _qubit_choices(n, nq):
    return list(itertools.permutations(range(n), nq))


@needs_ref
@pytest.mark.parametrize("n", [5])
def test_port_equals_reference_every_gate_every_qubit_choice(n):
    psi0 = random_state(n, 11)
    with oracle.RefOracle(n) as ref, oracle.PortOracle(n) as port:
        ref.set_multithreading(False)
        for g in gates.all_gate_samples():
            for qs in _qubit_choices(n, g.nq):
                args = list(qs) + [0] * (3 - len(qs))
                ref.set_state(psi0)
                port.set_state(psi0)
                ref.apply(g, *args)
                port.apply(g, *args)
                assert same(ref.state(), port.state()), (g.name, qs)
                # flag-less path (AppliedGate) through the generic kernels: same numbers to 1 ulp
                ref.set_state(psi0)
                ref.apply_matrix(g.nq, g.matrix, *args)
                port.set_state(psi0)
                port.apply_matrix(g.nq, g.matrix, *args)
                assert same(ref.state(), port.state()), (g.name, qs, "flagless")


@needs_ref
def test_reference_fast_kernels_equal_dense_operator_path():
    """The reference's second, independent path (getOperatorMatrix, SimpleGates.h:211-232):
    confirms there are no latent kernel bugs to be bug-compatible with."""
    n = 5
    psi0 = random_state(n, 3)
    with oracle.RefOracle(n) as a, oracle.RefOracle(n) as b:
        for g in gates.all_gate_samples():
            for qs in _qubit_choices(n, g.nq)[:: 3 if g.nq == 3 else 1]:
                args = list(qs) + [0] * (3 - len(qs))
                a.set_state(psi0)
                b.set_state(psi0)
                a.apply(g, *args)
                b.apply_dense(g, *args)
                assert np.max(np.abs(a.state() - b.state())) < 4e-16, (g.name, qs)


@needs_ref
def test_omp_path_equals_serial_path():
    n = 14  # >= OneQubitOmpLimit (8192 states) so the Omp twins run
    psi0 = random_state(n, 5)
    circ = circuits.random_circuit(n, 3) + [(gates.SwapGate(), 0, 13, 0), (gates.FredkinGate(), 1, 2, 12),
                                            (gates.ControlledRxGate(0.3), 13, 0, 0), (gates.CCZGate(), 0, 5, 13)]
    with oracle.RefOracle(n) as a, oracle.RefOracle(n) as b, oracle.PortOracle(n) as p:
        a.set_multithreading(True)
        b.set_multithreading(False)
        for o in (a, b, p):
            o.set_state(psi0)
            o.apply_circuit(circ)
        assert same(a.state(), b.state())
        assert same(a.state(), p.state())


@needs_ref
@pytest.mark.parametrize("n,sq,eq,swap", [(10, 0, 9, True), (10, 0, 9, False), (9, 2, 6, True), (8, 3, 7, False), (4, 1, 1, True)])
def test_port_qft_equals_reference(n, sq, eq, swap):
    psi0 = random_state(n, 21)
    with oracle.RefOracle(n) as ref, oracle.PortOracle(n) as port:
        for inverse in (False, True):
            ref.set_state(psi0)
            port.set_state(psi0)
            ref.qft(sq, eq, swap, inverse)
            port.qft(sq, eq, swap, inverse)
            assert same(ref.state(), port.state())
            # the python gate list (what the GPU engine is fed) is the same circuit
            port.set_state(psi0)
            port.apply_circuit(circuits.qft_circuit(n, sq, eq, swap, inverse))
            assert same(ref.state(), port.state())


def test_qft_is_sqrtN_ifft():
    """Analytic anchor: QFT(a) = sqrt(N) * ifft(a), IQFT(a) = fft(a) / sqrt(N)."""
    n = 10
    psi0 = random_state(n, 7)
    with oracle.best_oracle(n) as o:
        o.set_state(psi0)
        o.qft()
        assert np.max(np.abs(o.state() - np.sqrt(1 << n) * np.fft.ifft(psi0))) < 5e-15
        o.set_state(psi0)
        o.qft(inverse=True)
        assert np.max(np.abs(o.state() - np.fft.fft(psi0) / np.sqrt(1 << n))) < 5e-15


@needs_ref
def test_measurement_port_equals_reference():
    n = 9
    psi0 = random_state(n, 13)
    with oracle.RefOracle(n) as ref, oracle.PortOracle(n) as port:
        ref.set_multithreading(False)
        for p in draws(40, 1):
            ref.set_state(psi0)
            port.set_state(psi0)
            assert ref.measure_all_nocollapse(p) == port.measure_all_nocollapse(p)
            assert ref.measure_nocollapse(2, 5, p) == port.measure_nocollapse(2, 5, p)
            assert ref.measure_all(p) == port.measure_all(p)
            assert same(ref.state(), port.state())
            for (a, b) in ((0, 0), (8, 8), (3, 3), (0, 8), (2, 6), (7, 8)):
                ref.set_state(psi0)
                port.set_state(psi0)
                assert ref.measure(a, b, p) == port.measure(a, b, p)
                assert same(ref.state(), port.state())
        # fallbacks: a draw above the total mass (unnormalised state)
        half = psi0 * 0.5
        ref.set_state(half)
        port.set_state(half)
        assert ref.measure_all_nocollapse(0.9) == port.measure_all_nocollapse(0.9) == 0
        assert ref.measure_all(0.9) == port.measure_all(0.9) == (1 << n) - 1
        for q in range(n):
            ref.set_state(psi0)
            port.set_state(psi0)
            assert ref.qubit_probability(q) == port.qubit_probability(q)


@needs_ref
def test_reference_measurement_analytic_case():
    """MeasurementsTests.cpp:34-167: 1/2|00> - i/2|01> + 1/sqrt2|11>, MeasureQubit, closed forms."""
    s2 = 1 / math.sqrt(2)
    psi = np.array([0.5, -0.5j, 0, s2], dtype=np.complex128)
    with oracle.RefOracle(2) as ref:
        for p in draws(50, 9):
            ref.set_state(psi)
            r = ref.measure(0, 0, p)
            st = ref.state()
            if r == 0:
                want = np.array([1, 0, 0, 0], dtype=np.complex128)
            else:
                want = np.array([0, -1j / math.sqrt(3), 0, math.sqrt(2.0 / 3.0)], dtype=np.complex128)
            assert np.max(np.abs(st - want)) < 1e-10


@needs_ref
def test_draper_adder_known_answer():
    """AdderTests.cpp:213-327 / DraperAdder.h:38-56: sub-register QFT + CPhase + IQFT = n1 + n2."""
    nb = 3
    with oracle.RefOracle(1) as ref:
        for n1 in range(8):
            for n2 in range(0, 8, 3):
                amp = ref.draper_add(nb, n1, n2)
                expect = n1 | (((n1 + n2) % 8) << nb)
                assert abs(abs(amp[expect]) ** 2 - 1) < 1e-13, (n1, n2)


@needs_ref
def test_grover_with_gates_matches_gate_list():
    """GroverAlgorithm.h:128-242 vs circuits.grover_gates_circuit (what the GPU engine is fed)."""
    N, marked = 5, 0b10110
    with oracle.RefOracle(1) as ref:
        want = ref.grover_gates(N, marked)
    nq = 2 * N - 1
    with oracle.PortOracle(nq) as port:
        port.apply_circuit(circuits.grover_gates_circuit(N, marked))
        got = port.state()
    assert np.max(np.abs(want - got)) < 1e-14
    k = circuits.grover_iterations(N)
    p_marked = sum(abs(got[marked | (hi << N)]) ** 2 for hi in range(1 << (nq - N)))
    assert abs(p_marked - math.sin((2 * k + 1) * math.asin(2 ** (-N / 2))) ** 2) < 1e-12


@needs_ref
def test_ncnot_gate_list_matches_reference_record_uncompute():
    n = 12
    psi0 = random_state(n, 17)
    ctrl = [0, 1, 2, 3, 4]
    with oracle.RefOracle(n) as ref, oracle.PortOracle(n) as port:
        ref.set_state(psi0)
        ref.ncnot(ctrl, 5, 6)
        port.set_state(psi0)
        port.apply_circuit(circuits.ncnot_circuit(ctrl, 5, 6))
        assert same(ref.state(), port.state())


def test_check_qubits_errors():
    with oracle.best_oracle(3) as o:
        with pytest.raises(ValueError):
            o.apply(gates.HadamardGate(), 3)
        with pytest.raises(ValueError):
            o.apply(gates.CNOTGate(), 0, 0)
        with pytest.raises(ValueError):
            o.apply(gates.ToffoliGate(), 0, 1, 1)


def test_golden_fixtures_with_port():
    """tests/golden/*.npz were produced by the compiled reference (make_golden.py); the port
    must reproduce them bit-for-bit, so the oracle stays pinned where /root/reference is absent."""
    import golden_util

    cases = golden_util.load_all()
    assert cases, "no golden fixtures committed"
    for name, case in cases.items():
        n = int(case["n"])
        with oracle.PortOracle(n) as port:
            golden_util.check_case(port, case, exact=True)
