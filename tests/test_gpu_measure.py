"""GPU parity of the measurement path against the compiled reference (run with `-m gpu`).

BASELINE.json north_star: "Measurement outcomes must be bit-exact when the same uniform random draws are injected."
The reference's outcome is the first i with prob <= acc_i for its strictly sequential fp64 running sum; the engine
reproduces that sum bit for bit (csrc/reduce_kernels.cuh), so the hardest draws -- exactly ON a bin edge and one ulp to
either side -- must agree too.  RepeatedMeasure is compared with the reference's own RepeatedMeasure (all four
overloads, seeded identically), including its table cut at 1 - epsilon and the end() outcome (QubitRegister.h:250-254, 268).
"""
import numpy as np
import pytest

import oracle
import qcsim_b200
from conftest import draws, random_state
from gpu_adapter import GpuSim
from qcsim_b200 import circuits, gates

pytestmark = pytest.mark.gpu


def edge_draws(psi, count, seed):
    """the legal draws closest to the bin edges of the reference's running sum of |psi|^2.  A real draw is
    1 - k 2^-53, a multiple of 2^-53 (and the oracle driver can only inject those): for an edge acc_i the two
    neighbouring multiples -- the lower one IS the edge whenever acc_i >= 0.5 or happens to be representable."""
    p = psi.real * psi.real + psi.imag * psi.imag           # rounded like std::norm in the -msse2 build
    acc = np.cumsum(p)                                       # sequential fp64 (tests/test_sequential_sum.py)
    rng = np.random.default_rng(seed)
    idx = list(rng.integers(0, len(acc), size=count))
    # the crossings of powers of two are where a parallel sum would go wrong first
    for e in range(1, 12):
        i = int(np.searchsorted(acc, 2.0 ** -e))
        if 0 < i < len(acc):
            idx += [i - 1, i]
    out = []
    for i in idx:
        lo = np.floor(acc[i] * 2.0 ** 53) / 2.0 ** 53
        out += [lo - 2.0 ** -53, lo, lo + 2.0 ** -53]
    return [float(x) for x in out if 0.0 < x <= 1.0]


@pytest.mark.parametrize("n", [9, 14, 19])
def test_outcomes_on_bin_edges_are_the_references(n):
    circ = circuits.random_circuit(n, 3, seed=40 + n)
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        assert ref.kind == "reference"
        ref.apply_circuit(circ)
        gpu.apply_circuit(circ)
        psi = ref.state()
        assert np.max(np.abs(gpu.state() - psi)) <= 1e-12
        gpu.set_state(psi)                                   # identical amplitudes: the comparison is about the scan only
        tested = 0
        for p in edge_draws(psi, 25 if n < 19 else 12, n) + list(draws(10, n)):
            assert gpu.measure_all_nocollapse(p) == ref.measure_all_nocollapse(p), (n, p)
            tested += 1
        assert tested > 40
        a, b = 1, n - 2
        for p in edge_draws(psi, 6, n + 1):
            assert gpu.measure_nocollapse(a, b, p) == ref.measure_nocollapse(a, b, p), (n, p)


def test_unnormalised_and_sparse_states():
    n = 13
    with oracle.best_oracle(n) as ref, GpuSim(n) as gpu:
        psi = np.zeros(1 << n, dtype=np.complex128)
        psi[[5, 4097, 8000]] = [0.6, 0.3j, np.sqrt(1 - 0.45)]
        for scale in (1.0, 0.7, 1.3):
            ref.set_state(psi * scale)
            gpu.set_state(psi * scale)
            for p in (2.0 ** -53, np.floor(0.36 * scale * scale * 2.0 ** 53) / 2.0 ** 53, np.floor(0.36 * scale * scale * 2.0 ** 53 + 1) / 2.0 ** 53, 0.4375, 0.453125, 1 - 2.0 ** -20, 1.0):
                assert gpu.measure_all_nocollapse(p) == ref.measure_all_nocollapse(p), (scale, p)


@pytest.mark.parametrize("n", [6, 12, 16])
def test_repeated_measure_is_the_references(n):
    """QubitRegister.h:227-429: all four overloads, seeded identically; states with norm below / above one exercise the end()
    outcome (= table size) and the table cut at 1 - epsilon; one shot takes the MeasureNoCollapse shortcut."""
    psi = random_state(n, 17)
    with oracle.best_oracle(n) as ref:
        for scale in (1.0, 0.8, 1.25):
            ref.set_state(psi * scale)
            with qcsim_b200.QubitRegister(n, seed=1) as reg:
                reg.setRegisterStorageFastNoNormalize(psi * scale)
                for seed, shots in ((3, 1), (4, 2), (5, 400), (6, 5000 if n > 6 else 300)):
                    reg.rng.seed(seed)
                    assert reg.RepeatedMeasure(shots) == ref.repeated_measure(seed, shots), (n, scale, shots)
                    reg.rng.seed(seed)
                    assert reg.RepeatedMeasure(1, n - 2, shots) == ref.repeated_measure(seed, shots, 1, n - 2), (n, scale, shots, "range")
                    reg.rng.seed(seed)
                    assert reg.RepeatedMeasure(shots) == ref.repeated_measure(seed, shots, unordered=True)
                    reg.rng.seed(seed)
                    assert reg.RepeatedMeasure(0, 0, shots) == ref.repeated_measure(seed, shots, 0, 0, unordered=True)
                if scale == 0.8:
                    reg.rng.seed(9)
                    assert (1 << n) in reg.RepeatedMeasure(400)      # the reference's end() outcome really occurs


def test_expectation_fidelity_save_restore_against_the_reference():
    """QubitRegister.h:527-534, 600-616, 646-660 through the reference's own members"""
    n = 10
    psi = random_state(n, 23)
    obs = [(gates.PauliZGate(), 0, 0, 0), (gates.PauliXGate(), 3, 0, 0), (gates.ControlledZGate(), 5, 7, 0), (gates.RyGate(0.4), 9, 0, 0)]
    with oracle.best_oracle(n) as ref, qcsim_b200.QubitRegister(n, seed=1) as reg:
        ref.set_state(psi)
        reg.setRegisterStorageFastNoNormalize(psi)
        ev_ref = ref.expectation_value(obs)
        ev = reg.ExpectationValue(obs)
        assert abs(ev - ev_ref) <= 1e-12
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) == 0.0     # both restore the state afterwards
        other = random_state(n, 24)
        assert abs(reg.stateFidelity(other) - ref.state_fidelity(other)) <= 1e-12
        # RestoreState without a saved state is a no-op (:607), also with queued gates
        reg.set_fusion(True)
        reg.ApplyGate(gates.HadamardGate(), 2)
        reg.RestoreState()
        ref.apply(gates.HadamardGate(), 2)
        ref.restore_state()
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
        reg.set_fusion(False)
        reg.SaveState()
        ref.save_state()
        for g in circuits.random_circuit(n, 2, seed=5):
            reg.ApplyGate(*g)
            ref.apply(*g)
        reg.RestoreState()
        ref.restore_state()
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
        reg.ApplyGate(gates.RxGate(0.3), 4)
        ref.apply(gates.RxGate(0.3), 4)
        reg.RestoreStateDestructive()
        ref.restore_state(destructive=True)
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
        reg.ApplyGate(gates.RxGate(0.3), 4)
        ref.apply(gates.RxGate(0.3), 4)
        reg.RestoreState()                                                    # nothing saved any more: no-op on both sides
        ref.restore_state()
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
