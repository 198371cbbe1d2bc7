"""Circuit files (include/qcsim_b200.h "circuit files"): one recorded gate stream, replayed by the engine
(qcsim_sv_apply_circuit_file) and by the compiled reference (oracle/ref_driver.cpp: ref_apply_circuit_file)."""
import os

import numpy as np
import pytest

import oracle
from conftest import random_state
from qcsim_b200 import _lib, circuits, gates


def sample_circuit(n):
    return circuits.random_circuit(n, 3, seed=9) + circuits.qft_circuit(n, 1, n - 2) + [
        (gates.AppliedGate(np.linalg.qr(random_state(6, 2).reshape(8, 8))[0]), 0, n - 1, 3), (gates.FredkinGate(), 2, 0, 5),
        (gates.iSwapGate(), 1, 4, 0), (gates.ControlledRyGate(0.3), n - 1, 0, 0)]


def test_file_round_trip_and_reference_replay(tmp_path):
    n = 9
    circ = sample_circuit(n)
    path = str(tmp_path / "circuit.qcc")
    circuits.save_circuit(path, n, circ)
    nn, back = circuits.load_circuit(path)      # the C ABI loader, no GPU needed
    assert nn == n and len(back) == len(circ)
    for (g, q, c1, c2), (g2, q2, c12, c22) in zip(circ, back):
        assert (q, c1, c2) == (q2, c12, c22) and g.flags == g2.flags and g.gate_id == g2.gate_id
        assert np.array_equal(g.matrix, g2.matrix) and tuple(g2.params)[: len(g.params)] == tuple(float(p) for p in g.params)
    if not oracle.ref_available("sse2"):
        pytest.skip("compiled reference not built here")
    psi = random_state(n, 3)
    with oracle.best_oracle(n) as a, oracle.best_oracle(n) as b:
        a.set_state(psi)
        b.set_state(psi)
        a.apply_circuit(circ)
        assert b.apply_circuit_file(path) == len(circ)
        assert np.array_equal(a.state(), b.state())   # same reference kernels either way


def test_loader_rejects_garbage(tmp_path):
    import ctypes as C

    lib = _lib.load()
    bad = tmp_path / "bad.qcc"
    bad.write_bytes(b"not a circuit file at all")
    n, ptr, cnt = C.c_uint32(), C.c_void_p(), C.c_uint64()
    assert lib.qcsim_circuit_load(str(bad).encode(), C.byref(n), C.byref(ptr), C.byref(cnt)) == _lib.ERR_BAD_ARG
    good = tmp_path / "good.qcc"
    circuits.save_circuit(str(good), 5, [(gates.HadamardGate(), 1, 0, 0), (gates.CNOTGate(), 0, 3, 0)])
    data = good.read_bytes()
    (tmp_path / "cut.qcc").write_bytes(data[:-40])
    assert lib.qcsim_circuit_load(str(tmp_path / "cut.qcc").encode(), C.byref(n), C.byref(ptr), C.byref(cnt)) == _lib.ERR_BAD_ARG
    assert lib.qcsim_circuit_load(str(tmp_path / "missing.qcc").encode(), C.byref(n), C.byref(ptr), C.byref(cnt)) == _lib.ERR_BAD_ARG


@pytest.mark.gpu
def test_engine_replays_the_file_like_the_reference(tmp_path):
    import qcsim_b200

    n = 15
    circ = sample_circuit(n)
    path = str(tmp_path / "circuit.qcc")
    circuits.save_circuit(path, n, circ)
    psi = random_state(n, 3)
    with oracle.best_oracle(n) as ref:
        ref.set_state(psi)
        ref.apply_circuit_file(path)
        want = ref.state()
    for fusion in (False, True):
        with qcsim_b200.QubitRegister(n, seed=1) as reg:
            reg.setRegisterStorageFastNoNormalize(psi)
            reg.set_fusion(fusion)
            reg.ApplyCircuitFile(path)
            assert np.max(np.abs(reg.getRegisterStorage() - want)) <= 1e-12
    with qcsim_b200.QubitRegister(n + 1, seed=1) as reg:
        with pytest.raises(_lib.QcsimError):
            reg.ApplyCircuitFile(path)          # recorded for another register size


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 4, 9])
def test_apply_operator_matrix_vs_reference(n):
    """ApplyOperatorMatrix (QubitRegister.h:499-505) and its recording for Compute / Uncompute"""
    import qcsim_b200

    rng = np.random.default_rng(n)
    d = 1 << n
    u = np.linalg.qr(rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d)))[0]
    psi = random_state(n, 8)
    with oracle.best_oracle(n) as ref, qcsim_b200.QubitRegister(n, seed=1) as reg:
        ref.set_state(psi)
        reg.setRegisterStorageFastNoNormalize(psi)
        ref.apply_operator_matrix(u)
        reg.ApplyOperatorMatrix(u)
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
        # recorded and undone: Uncompute applies the adjoint (:573-590)
        reg.ComputeStart()
        reg.ApplyOperatorMatrix(u)
        if n >= 2:
            reg.ApplyGate(gates.CNOTGate(), 0, 1)
        reg.ComputeEnd()
        reg.Uncompute()
        assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= 1e-12
    with qcsim_b200.QubitRegister(14, seed=1) as big:   # above QCSIM_MAX_OPERATOR_QUBITS: refused, not attempted
        import ctypes as C

        rc = big._lib.qcsim_sv_apply_operator(big._h, u.ctypes.data_as(C.c_void_p))
        assert rc == _lib.ERR_UNSUPPORTED
