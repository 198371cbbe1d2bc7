"""GPU parity tests (run with `-m gpu` on the B200 box).  Every call goes through the C ABI.

Tolerances (BASELINE.json north_star): amplitudes max-abs <= 1e-12 vs the reference on the same
circuit, |norm^2_gpu - norm^2_ref| <= 1e-12, measurement outcomes identical for identical
injected draws.
"""
import itertools
import math

import numpy as np
import pytest

import golden_util
import oracle
from conftest import draws, random_state
from gpu_adapter import GpuSim
from qcsim_b200 import circuits, gates

pytestmark = pytest.mark.gpu

TOL = 1e-12


def maxdiff(a, b):
    return float(np.max(np.abs(a - b)))


@pytest.mark.parametrize("n", [3, 5, 7])
def test_every_gate_every_qubit_choice(n):
    psi0 = random_state(n, 11)
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        worst = 0.0
        for g in gates.all_gate_samples():
            choices = list(itertools.permutations(range(n), g.nq))
            if n == 7 and g.nq == 3:
                choices = choices[::5]
            for qs in choices:
                args = list(qs) + [0] * (3 - len(qs))
                ref.set_state(psi0)
                ref.apply(g, *args)
                want = ref.state()
                gpu.set_state(psi0)
                gpu.apply(g, *args)
                d = maxdiff(gpu.state(), want)
                assert d <= TOL, (g.name, qs, d)
                gpu.set_state(psi0)  # flag-less path: classified from the matrix
                gpu.apply_matrix(g.nq, g.matrix, *args)
                d2 = maxdiff(gpu.state(), want)
                assert d2 <= TOL, (g.name, qs, "flagless", d2)
                worst = max(worst, d, d2)
        assert worst < 1e-14


@pytest.mark.parametrize("n", [1, 2])
def test_tiny_registers(n):
    psi0 = random_state(n, 2)
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        for g in gates.all_gate_samples():
            if g.nq > n:
                continue
            for qs in itertools.permutations(range(n), g.nq):
                args = list(qs) + [0] * (3 - len(qs))
                ref.set_state(psi0)
                ref.apply(g, *args)
                gpu.set_state(psi0)
                gpu.apply(g, *args)
                assert maxdiff(gpu.state(), ref.state()) <= TOL, (g.name, qs)


def test_every_target_position_mid_size():
    """n = 13: every qubit as target / control of every kernel shape (exercises the 256-bit,
    qubit-0 and generic variants and grids larger than one block)."""
    n = 13
    psi0 = random_state(n, 4)
    shapes = [gates.HadamardGate(), gates.RzGate(0.7), gates.PauliYGate(), gates.TGate(),
              gates.CNOTGate(), gates.ControlledPhaseShiftGate(0.3), gates.ControlledRyGate(1.1), gates.SwapGate(),
              gates.iSwapGate(), gates.DecrementGate(), gates.ToffoliGate(), gates.FredkinGate(), gates.CCZGate(),
              gates.AppliedGate(np.linalg.qr(random_state(6, 1).reshape(8, 8))[0])]
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        for g in shapes:
            for q in range(n):
                others = [(q + 1) % n, (q + n - 1) % n] if q % 2 else [(q + 5) % n, (q + 9) % n]
                args = [q] + others[: g.nq - 1] + [0] * (3 - g.nq)
                ref.set_state(psi0)
                ref.apply(g, *args)
                gpu.set_state(psi0)
                gpu.apply(g, *args)
                assert maxdiff(gpu.state(), ref.state()) <= TOL, (g.name, args)


def test_golden_fixtures():
    for name, case in golden_util.load_all().items():
        with GpuSim(int(case["n"])) as gpu:
            golden_util.check_case(gpu, case, exact=False, tol=TOL)


@pytest.mark.parametrize("n,layers", [(16, 6), (20, 4), (22, 2)])
def test_random_circuit_vs_oracle(n, layers):
    circ = circuits.random_circuit(n, layers)
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        ref.apply_circuit(circ)
        gpu.apply_circuit(circ)
        assert maxdiff(gpu.state(), ref.state()) <= TOL
        assert abs(gpu.norm2() - ref.norm2()) <= TOL


def test_config2_full_depth_vs_reference():
    """BASELINE config 2 at its stated DEPTH (200 layers; the generator is the one bench.py uses: 6800 gate applications
    at 24 qubits, 8600 at 30) at 24 qubits, the largest size the compiled reference finishes in about a minute: amplitudes within 1e-12
    and norm within 1e-12 of the reference, gate by gate and as fused blocks (whose 8x8 products are formed on the
    host and applied with FMA, i.e. a different rounding order)."""
    n, layers = 24, 200
    circ = circuits.random_circuit(n, layers)
    assert len(circ) == 34 * layers
    with oracle.best_oracle(n) as ref:
        assert ref.kind == "reference"
        ref.apply_circuit(circ)
        want, want_norm = ref.state(), ref.norm2()
    for mode in ("unfused", "fused"):
        with GpuSim(n, batch=(mode == "fused")) as gpu:
            gpu.apply_circuit(circ)
            d = maxdiff(gpu.state(), want)
            assert d <= TOL, (mode, d)
            assert abs(gpu.norm2() - want_norm) <= TOL, mode
            print(f"\n[config 2, {n} q x {layers} layers, {mode}] max|d| = {d:.2e}, norm2 = {gpu.norm2():.15f} (reference {want_norm:.15f})")


@pytest.mark.parametrize("n", [10, 20])
def test_qft_iqft_config1(n):
    """BASELINE config 1: QFT then IQFT with MeasureAll, three start states."""
    for start in ("zero", "basis", "random"):
        spec = {"n": n, "start": start, "basis": 0x5A5A5 % (1 << n), "state_seed": 7}
        psi0 = golden_util.start_state(spec)
        with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
            ref.set_state(psi0)
            gpu.set_state(psi0)
            ref.qft()
            gpu.qft()
            got = gpu.state()
            assert maxdiff(got, ref.state()) <= TOL
            assert maxdiff(got, np.sqrt(1 << n) * np.fft.ifft(psi0)) <= 1e-11
            ref.qft(inverse=True)
            gpu.qft(inverse=True)
            got = gpu.state()
            assert maxdiff(got, ref.state()) <= TOL
            assert maxdiff(got, psi0) <= 1e-11
            assert abs(gpu.norm2() - ref.norm2()) <= TOL
            for p in draws(6, 42):
                assert gpu.measure_all_nocollapse(p) == ref.measure_all_nocollapse(p)
            p = draws(1, 5)[0]
            assert gpu.measure_all(p) == ref.measure_all(p)
            assert maxdiff(gpu.state(), ref.state()) == 0.0


def test_qft_gate_by_gate_equals_engine_call():
    n = 12
    psi0 = random_state(n, 9)
    import qcsim_b200

    with GpuSim(n) as a, GpuSim(n) as b:
        for (sq, eq, swap) in ((0, n - 1, True), (2, 9, False), (3, 3, True)):
            a.set_state(psi0)
            b.set_state(psi0)
            a.qft(sq, eq, swap, False)
            qcsim_b200.QuantumFourierTransform(n, sq, eq).QFT(b.reg, swap)
            assert maxdiff(a.state(), b.state()) <= 1e-15
            a.qft(sq, eq, swap, True)
            qcsim_b200.QuantumFourierTransform(n, sq, eq).IQFT(b.reg, swap)
            assert maxdiff(a.state(), b.state()) <= 1e-15
            assert maxdiff(a.state(), psi0) <= 1e-12


@pytest.mark.parametrize("n", [4, 9, 13, 17])
@pytest.mark.parametrize("strict", [False, True])
def test_measurement_outcomes_bit_exact(n, strict):
    psi0 = random_state(n, 13)
    ranges = [(0, 0), (n - 1, n - 1), (n // 2, n // 2), (0, n - 1), (1, n // 2), (n // 2, n - 1)]
    with oracle.best_oracle(n) as ref, GpuSim(n, strict=strict) as gpu:
        for p in draws(12 if n < 17 else 4, n):
            ref.set_state(psi0)
            gpu.set_state(psi0)
            assert gpu.measure_all_nocollapse(p) == ref.measure_all_nocollapse(p)
            assert gpu.measure_nocollapse(1, n // 2, p) == ref.measure_nocollapse(1, n // 2, p)
            assert gpu.measure_all(p) == ref.measure_all(p)
            assert maxdiff(gpu.state(), ref.state()) == 0.0
            for (a, b) in ranges:
                ref.set_state(psi0)
                gpu.set_state(psi0)
                assert gpu.measure(a, b, p) == ref.measure(a, b, p), (a, b, p)
                assert maxdiff(gpu.state(), ref.state()) <= TOL
        for q in range(n):
            ref.set_state(psi0)
            gpu.set_state(psi0)
            assert abs(gpu.qubit_probability(q) - ref.qubit_probability(q)) <= TOL


def test_measurement_fallbacks_and_edges():
    """draw above the total mass: MeasureAll -> last state, others -> 0 (QubitRegister.h:173,623);
    a draw exactly on a bin edge picks that bin (prob <= accum)."""
    n = 6
    psi = random_state(n, 3) * 0.5
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        ref.set_state(psi)
        gpu.set_state(psi)
        assert gpu.measure_all_nocollapse(0.9) == ref.measure_all_nocollapse(0.9) == 0
        assert gpu.measure_nocollapse(1, 3, 0.9) == ref.measure_nocollapse(1, 3, 0.9) == 0
        assert gpu.measure_all(0.9) == ref.measure_all(0.9) == (1 << n) - 1
        # exact edges: amplitudes 1/2 on four states -> cumulative 0.25, 0.5, 0.75, 1.0 exactly
        v = np.zeros(1 << n, dtype=np.complex128)
        v[[3, 17, 40, 63]] = 0.5
        for p, want in ((0.25, 3), (0.5, 17), (0.75, 40), (1.0, 63), (0.2500000000000001, 17)):
            ref.set_state(v)
            gpu.set_state(v)
            assert gpu.measure_all_nocollapse(p) == ref.measure_all_nocollapse(p) == want


def test_analytic_measurement_case():
    """MeasurementsTests.cpp:34-167 through the GPU register."""
    s2 = 1 / math.sqrt(2)
    psi = np.array([0.5, -0.5j, 0, s2], dtype=np.complex128)
    with GpuSim(2) as gpu:
        for p in draws(30, 9):
            gpu.set_state(psi)
            r = gpu.measure(0, 0, p)
            want = np.array([1, 0, 0, 0]) if r == 0 else np.array([0, -1j / math.sqrt(3), 0, math.sqrt(2.0 / 3.0)])
            assert maxdiff(gpu.state(), want) < 1e-10
            gpu.set_state(psi)
            r = gpu.measure(1, 1, p)
            want = np.array([s2, -1j * s2, 0, 0]) if r == 0 else np.array([0, 0, 0, 1])
            assert maxdiff(gpu.state(), want) < 1e-10


def test_reference_error_conventions():
    import qcsim_b200

    with qcsim_b200.QubitRegister(3, seed=1) as reg:
        with pytest.raises(ValueError, match="Qubit number is too high"):
            reg.ApplyGate(gates.HadamardGate(), 3)
        with pytest.raises(ValueError, match="Controlling qubit number is too high"):
            reg.ApplyGate(gates.CNOTGate(), 0, 5)
        with pytest.raises(ValueError, match="Qubit and controlling qubit are the same"):
            reg.ApplyGate(gates.CNOTGate(), 1, 1)
        with pytest.raises(ValueError, match="Qubits must be different"):
            reg.ApplyGate(gates.ToffoliGate(), 0, 1, 1)
        # silent no-ops / zero returns (QubitRegister.h:63,69,76,84,114)
        reg.setToBasisState(5)
        reg.setToBasisState(8)
        reg.setRawAmplitude(99, 1.0)
        reg.setToQubitState(7)
        assert reg.getBasisStateAmplitude(8) == 0
        assert reg.getBasisStateProbability(5) == 1.0
        reg.setRegisterStorage(np.ones(4))  # wrong size: ignored (:514)
        assert reg.getBasisStateProbability(5) == 1.0
        reg.Clear()
        reg.Normalize()  # norm < 1e-20: no-op (:127)
        assert reg.norm2() == 0.0


def test_state_helpers_and_record_uncompute():
    import qcsim_b200

    n = 10
    psi0 = random_state(n, 8)
    with qcsim_b200.QubitRegister(n, seed=1) as reg, oracle.best_oracle(n) as ref:
        reg.setToCatState()
        st = reg.getRegisterStorage()
        assert abs(st[0] - 1 / math.sqrt(2)) < 1e-16 and abs(st[-1] - 1 / math.sqrt(2)) < 1e-16
        reg.setToEqualSuperposition()
        assert maxdiff(reg.getRegisterStorage(), np.full(1 << n, 1 / math.sqrt(1 << n))) == 0
        reg.setRegisterStorage(psi0 * 3.0)  # normalises (:516-517)
        assert maxdiff(reg.getRegisterStorage(), psi0) < 1e-15
        # n-controlled NOT with recorded compute / uncompute (NControlledNotWithAncilla.h:24-96)
        ctrl = [0, 1, 2, 3, 4]
        reg.ComputeStart()
        circ = circuits.ncnot_circuit(ctrl, 5, 6, clear_ancilla=False)
        for g in circ[:-1]:
            reg.ApplyGate(*g)
        reg.ComputeEnd()
        reg.ApplyGate(*circ[-1])
        reg.Uncompute()
        ref.set_state(psi0)
        ref.apply_circuit(circuits.ncnot_circuit(ctrl, 5, 6, clear_ancilla=True))
        assert maxdiff(reg.getRegisterStorage(), ref.state()) <= TOL
        # save / restore / clone / expectation value
        reg.SaveState()
        reg.ApplyGate(gates.HadamardGate(), 3)
        c = reg.Clone()
        reg.RestoreState()
        assert maxdiff(reg.getRegisterStorage(), ref.state()) <= TOL
        ev = reg.ExpectationValue([(gates.PauliZGate(), 0, 0, 0)])
        st = reg.getRegisterStorage()
        want = np.vdot(st, st * np.where(np.arange(1 << n) & 1, -1.0, 1.0))
        assert abs(ev - want) < 1e-13
        c.close()


def test_grover_with_gates_small():
    N, marked = 6, 0b101101
    nq = 2 * N - 1
    circ = circuits.grover_gates_circuit(N, marked)
    with oracle.best_oracle(nq) as ref, GpuSim(nq) as gpu:
        ref.apply_circuit(circ)
        gpu.apply_circuit(circ)
        got = gpu.state()
        assert maxdiff(got, ref.state()) <= TOL
        k = circuits.grover_iterations(N)
        p_marked = sum(abs(got[marked | (hi << N)]) ** 2 for hi in range(1 << (nq - N)))
        assert abs(p_marked - math.sin((2 * k + 1) * math.asin(2 ** (-N / 2))) ** 2) < 1e-12


# ---- fused gate blocks (qcsim_sv_apply_batch / fusion mode) ------------------------------------------
@pytest.mark.parametrize("mode", ["batch", "fusion"])
@pytest.mark.parametrize("n,layers", [(6, 4), (9, 5), (13, 6), (16, 6), (21, 3)])
def test_fused_random_circuit_vs_oracle(n, layers, mode):
    circ = circuits.random_circuit(n, layers, seed=n * 7 + 1)
    with oracle.best_oracle(n) as ref, GpuSim(n, fusion=(mode == "fusion"), batch=(mode == "batch")) as gpu:
        ref.apply_circuit(circ)
        gpu.apply_circuit(circ)
        assert maxdiff(gpu.state(), ref.state()) <= TOL
        assert abs(gpu.norm2() - ref.norm2()) <= TOL
        st = gpu.reg.stats()
        if n >= 9:
            assert st["state_passes"] < len(circ) / 2, st  # really fused


@pytest.mark.parametrize("n", [9, 14, 18])
def test_qubit_probability_between_fused_gates_keeps_the_rest_queued(n):
    """GetQubitProbability(q) in fusion mode runs only the queued gates that can change P(q) (engine.cu:
    engine_flush_for_diagonal_observable); every value along the way and the final state are the reference's."""
    circ = circuits.random_circuit(n, 6, seed=91 + n)
    rng = np.random.default_rng(n)
    with oracle.best_oracle(n) as ref, GpuSim(n, fusion=True) as gpu:
        for i, (g, q, c1, c2) in enumerate(circ):
            ref.apply(g, q, c1, c2)
            gpu.apply(g, q, c1, c2)
            if i % 7 == 6:
                qq = int(rng.integers(0, n))
                assert abs(gpu.reg.GetQubitProbability(qq) - ref.qubit_probability(qq)) <= TOL, (i, qq)
        passes_so_far = gpu.reg.stats()["state_passes"]
        assert maxdiff(gpu.state(), ref.state()) <= TOL
        if n >= 14:
            assert passes_so_far < len(circ) / 2, passes_so_far  # the probability reads did not force gate-by-gate passes


@pytest.mark.parametrize("n", [7, 12, 14])
def test_fused_all_gate_kinds(n):
    """every gate class (flagged and flag-less) through the tile engine, on rotating qubits, twice
    with different qubit offsets so targets land on low, tile and out-of-tile positions"""
    psi0 = random_state(n, 31)
    circ = []
    for shift in (0, 3, n - 3):
        for (g, q, c1, c2) in golden_util.all_gates_circuit(n):
            circ.append((g, (q + shift) % n, (c1 + shift) % n if g.nq >= 2 else 0, (c2 + shift) % n if g.nq >= 3 else 0))
    with oracle.best_oracle(n) as ref, GpuSim(n, batch=True) as gpu:
        ref.set_state(psi0)
        gpu.set_state(psi0)
        ref.apply_circuit(circ)
        gpu.apply_circuit(circ)
        assert maxdiff(gpu.state(), ref.state()) <= TOL


@pytest.mark.parametrize("n", [12, 20])
def test_fused_qft_vs_oracle(n):
    psi0 = random_state(n, 3)
    with oracle.best_oracle(n) as ref, GpuSim(n, fusion=True) as gpu:
        ref.set_state(psi0)
        gpu.set_state(psi0)
        ref.qft()
        gpu.qft()
        assert maxdiff(gpu.state(), ref.state()) <= TOL
        ref.qft(2, n - 3, True, True)
        gpu.qft(2, n - 3, True, True)
        assert maxdiff(gpu.state(), ref.state()) <= TOL


def test_fused_equals_unfused_large():
    """24 qubits: fused and one-gate-per-pass execution of the same circuit agree to 1e-13"""
    n = 24
    circ = circuits.random_circuit(n, 4) + circuits.qft_circuit(n, 3, 20)
    with GpuSim(n) as a, GpuSim(n, batch=True) as b:
        a.apply_circuit(circ)
        b.apply_circuit(circ)
        assert maxdiff(a.state(), b.state()) <= 1e-13
        assert abs(a.norm2() - b.norm2()) <= 1e-13


def test_fusion_mode_flushes_on_observation():
    n = 10
    circ = circuits.random_circuit(n, 2)
    with oracle.best_oracle(n) as ref, GpuSim(n, fusion=True) as gpu:
        for i, g in enumerate(circ):
            ref.apply(*g)
            gpu.apply(*g)
            if i % 17 == 5:
                assert abs(gpu.qubit_probability(i % n) - ref.qubit_probability(i % n)) <= TOL
            if i % 29 == 7:
                p = draws(1, i)[0]
                assert gpu.measure(0, 2, p) == ref.measure(0, 2, p)
        assert maxdiff(gpu.state(), ref.state()) <= TOL


@pytest.mark.parametrize("n", [5, 13, 18])
def test_repeated_measure_equals_the_reference_draw_by_draw(n):
    """RepeatedMeasure (QubitRegister.h:227-273) = many draws against one cumulative table: every
    shot must be the outcome the reference's MeasureNoCollapse gives for the same draw."""
    psi0 = random_state(n, 23)
    shots = 300
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        ref.set_state(psi0)
        gpu.set_state(psi0)
        gpu.reg.rng.seed(1234)
        hist = gpu.reg.RepeatedMeasure(shots)
        gpu.reg.rng.seed(1234)
        d = [gpu.reg._draw() for _ in range(shots)]  # the draws RepeatedMeasure consumed
        want = {}
        for p in d:
            s = ref.measure_all_nocollapse(p)
            want[s] = want.get(s, 0) + 1
        assert hist == dict(sorted(want.items()))
        assert sum(hist.values()) == shots
        # sub-register variant: outcomes masked and shifted (QubitRegister.h:325-375)
        gpu.reg.rng.seed(99)
        sub = gpu.reg.RepeatedMeasure(1, n - 2, 50)
        gpu.reg.rng.seed(99)
        want = {}
        for _ in range(50):
            s = (ref.measure_all_nocollapse(gpu.reg._draw()) >> 1) & ((1 << (n - 2)) - 1)
            want[s] = want.get(s, 0) + 1
        assert sub == dict(sorted(want.items()))
        assert maxdiff(gpu.state(), psi0) == 0.0  # sampling does not touch the state
