"""GPU tests of the C++ drop-in facade (qcsim_b200/cpp/QubitRegister.h).

tests/cpp/facade_test[_fast].bin is QCSim's own algorithm code (QuantumFourierTransform.h,
GroverAlgorithm.h, DraperAdder.h, NControlledNotWithAncilla.h, QuantumAlgorithm.h, compiled
unmodified from the reference tree) running on the device-backed register.  What it dumps must
equal what the same reference code gives on the reference's CPU register (the oracle) to 1e-12,
and its seeded measurement outcomes must be identical.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import random_state

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


def binary(fast):
    path = os.path.join(HERE, "cpp", "facade_test_fast.bin" if fast else "facade_test.bin")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run tests/cpp/build_facade.sh where the reference tree exists")
    return path


def run(fast, *args):
    res = subprocess.run([binary(fast), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    out = {}
    for line in res.stdout.splitlines():
        k, _, v = line.partition(" ")
        out.setdefault(k, []).append(v)
    return out, res.stdout


def load(path, dim):
    v = np.fromfile(path, dtype=np.complex128)
    assert v.size % dim == 0
    return v.reshape(-1, dim)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("n,sq,eq,swap", [(12, 0, 11, 1), (14, 3, 10, 0), (20, 0, 19, 1)])
def test_reference_qft_header_on_device_register(tmp_path, fast, fusion, n, sq, eq, swap):
    psi = random_state(n, 17)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    psi.tofile(fin)
    run(fast, "qft", n, sq, eq, swap, fusion, fin, fout)
    got = load(fout, 1 << n)
    with oracle.best_oracle(n) as ref:
        ref.set_state(psi)
        ref.qft(sq, eq, bool(swap), False)
        assert np.max(np.abs(got[0] - ref.state())) <= TOL
        ref.qft(sq, eq, bool(swap), True)
        assert np.max(np.abs(got[1] - ref.state())) <= TOL
    assert np.max(np.abs(got[1] - psi)) <= 1e-11


@pytest.mark.parametrize("n_search,marked", [(4, 0b1010), (6, 0b110101)])
def test_reference_grover_with_gates(tmp_path, n_search, marked):
    fout = tmp_path / "g.bin"
    out, _ = run(False, "grover", n_search, marked, fout)
    n = int(out["qubits"][0])
    got = load(fout, 1 << n)[0]
    with oracle.best_oracle(n) as ref:
        want = ref.grover_gates(n_search, marked)
    assert np.max(np.abs(got - want)) <= TOL
    p_marked = sum(abs(got[i]) ** 2 for i in range(1 << n) if (i & ((1 << n_search) - 1)) == marked)
    assert p_marked > 0.9  # Tests.cpp:192-232


@pytest.mark.parametrize("n1,n2", [(3, 4), (7, 7), (5, 1)])
def test_reference_draper_adder_known_answer(tmp_path, n1, n2):
    nb = 3
    fout = tmp_path / "d.bin"
    out, _ = run(False, "draper", nb, n1, n2, fout)
    got = load(fout, 1 << (2 * nb))[0]
    with oracle.best_oracle(2 * nb) as ref:
        want = ref.draper_add(nb, n1, n2)
    assert np.max(np.abs(got - want)) <= TOL
    measured = int(out["measured"][0])
    assert measured == (n1 | (((n1 + n2) % (1 << nb)) << nb))  # AdderTests.cpp:213-327


def test_reference_ncnot_compute_uncompute(tmp_path):
    n = 9
    fout = tmp_path / "n.bin"
    run(False, "ncnot", n, fout)
    got = load(fout, 1 << n)[0]
    nc = (n + 1) // 2
    with oracle.best_oracle(n) as ref:
        from qcsim_b200 import gates
        for q in range(nc):
            ref.apply(gates.HadamardGate(), q)
        ref.ncnot(list(range(nc)), nc, nc + 1, True)
        assert np.max(np.abs(got - ref.state())) <= TOL


def test_seeded_measurements_match_reference(tmp_path):
    n, seed = 10, 987654321
    psi = random_state(n, 3)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    psi.tofile(fin)
    out, _ = run(False, "measure", n, seed, fin, fout)
    states = load(fout, 1 << n)
    with oracle.best_oracle(n) as ref:
        ref.set_state(psi)
        d = ref.draws(seed, 8)  # the draws `1. - uniformZeroOne(rng)` the facade makes after reseed(seed)
        assert [int(x) for x in out["nocollapse"]] == [ref.measure_all_nocollapse(d[i]) for i in range(4)]
        assert int(out["range_nocollapse"][0]) == ref.measure_nocollapse(1, n - 2, d[4])
        assert abs(float(out["p0"][0]) - ref.qubit_probability(0)) <= TOL
        assert int(out["qubit"][0]) == ref.measure(n - 1, n - 1, d[5])
        assert np.max(np.abs(states[0] - ref.state())) <= TOL
        assert int(out["range"][0]) == ref.measure(0, 2, d[6])
        assert np.max(np.abs(states[1] - ref.state())) <= TOL
        assert int(out["all"][0]) == ref.measure_all(d[7])
        assert np.max(np.abs(states[2] - ref.state())) == 0.0


def test_error_conventions_and_state_helpers():
    out, text = run(False, "misc")
    # QubitRegister.h:677-690: same exception type, same messages
    assert "1q too high: invalid_argument: Qubit number is too high" in text
    assert "2q ctrl too high: invalid_argument: Controlling qubit number is too high" in text
    assert "2q same: invalid_argument: Qubit and controlling qubit are the same" in text
    assert "3q ctrl too high: invalid_argument: Controlling qubit number is too high" in text
    assert "3q same: invalid_argument: Qubits must be different" in text
    assert float(out["amp_out_of_range"][0]) == 0.0 and float(out["amp0"][0]) == 1.0  # silent no-ops :63,76,114
    s = 1 / np.sqrt(2)
    assert [float(x) for x in out["cat"][0].split()] == [s, s]
    assert float(out["equal"][0]) == 0.25
    assert abs(float(out["expect_z_plus"][0])) < 1e-15 and float(out["expect_z_one"][0]) == -1.0  # Tests.cpp:665-723
    a, b = [float(x) for x in out["clone_indep"][0].split()]
    assert a == 1.0 and abs(b - s) < 1e-16
    assert float(out["restored"][0]) == 1.0
    assert abs(float(out["fidelity"][0]) - 1.0) < 1e-15
    re, im, im1 = [float(x) for x in out["adjusted"][0].split()]
    assert abs(re - 1 / np.sqrt(1.25)) < 1e-15 and abs(im) < 1e-15 and abs(im1 + 0.5 / np.sqrt(1.25)) < 1e-15
    assert out["repeated"][0].split()[0] in ("1", "2")
    assert out["threads_ok"][0] == "1"
    # QCSim's own QubitRegisterDebug.h compiled unchanged on the drop-in: writeToFile dumps |amplitude| per basis state
    ok, lines, sumsq = out["debug_dump"][0].split()
    assert ok == "1" and int(lines) == 8 and abs(float(sumsq) - 1.0) < 1e-5
