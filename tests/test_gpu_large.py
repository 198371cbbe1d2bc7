"""Full-size parity properties (BASELINE configs 2 and 3), where the oracle cannot run.

At 30+ qubits the CPU reference needs 2 x 16 x 2^n bytes and minutes per gate, so parity is
asserted through size-independent properties:
  * QFT of a basis state |k> is the analytic phase ramp exp(2 pi i j k / N) / sqrt(N);
  * IQFT(QFT(psi)) = psi and the norm stays 1 to 1e-12;
  * the fused and the unfused execution of the same random circuit agree on sampled ranges.
Sizes adapt to the free device memory (33 qubits = 128 GiB needs a whole B200).
"""
import numpy as np
import pytest

import qcsim_b200
from qcsim_b200 import circuits

pytestmark = pytest.mark.gpu
TOL = 1e-12


def free_gib():
    import torch

    free, _ = torch.cuda.mem_get_info(0)
    return free / (1 << 30)


def pick_qubits(copies=1, cap=33):
    n = cap
    while n > 20 and copies * (16 << n) / (1 << 30) > 0.9 * free_gib():
        n -= 1
    return n


def ranges(n, count=4096):
    dim = 1 << n
    return [0, dim // 3 & ~0xfff, dim // 2 - count // 2, dim - count]


def test_qft_of_basis_state_is_the_analytic_phase_ramp():
    n = pick_qubits()
    dim = 1 << n
    k = 0x2D5A5A5 % dim
    with qcsim_b200.QubitRegister(n, seed=1) as reg:
        reg.setToBasisState(k)
        reg.QFT()
        assert abs(reg.norm2() - 1.0) <= TOL
        for first in ranges(n):
            j = np.arange(first, first + 4096, dtype=np.uint64)
            # exp(2 pi i j k / N): reduce j*k mod N in integers first (exact), then one sincos
            frac = ((j * np.uint64(k)) & np.uint64(dim - 1)).astype(np.float64) / dim
            want = np.exp(2j * np.pi * frac) / np.sqrt(dim)
            got = reg.download(first, 4096)
            assert np.max(np.abs(got - want)) <= TOL, (n, first)
        reg.QFT(inverse=True)
        assert abs(reg.norm2() - 1.0) <= TOL
        a = reg.getBasisStateAmplitude(k)
        assert abs(a - 1.0) <= 1e-11, a
        print(f"\n[large] QFT/IQFT at {n} qubits ({16 << n >> 30} GiB): basis-state ramp and round trip ok")


def test_sub_register_qft_round_trip_large():
    n = pick_qubits(cap=31)
    with qcsim_b200.QubitRegister(n, seed=1) as reg:
        reg.setToBasisState(0)
        reg.set_fusion(True)
        reg.ApplyGates(circuits.random_circuit(n, 1))
        before = [reg.download(f, 4096) for f in ranges(n)]
        reg.QFT(3, n - 4, False, False)
        reg.QFT(3, n - 4, False, True)
        reg.QFT(0, n - 1, True, False)
        reg.QFT(0, n - 1, True, True)
        for f, b in zip(ranges(n), before):
            assert np.max(np.abs(reg.download(f, 4096) - b)) <= TOL


def test_fused_equals_unfused_at_full_size():
    """BASELINE config 2 at its stated size and depth: 30 qubits, 200 layers = 8600 gate applications, executed as
    fused gate blocks (TMA-staged DMMA passes) and gate by gate (one in-place kernel per gate, the reference's own
    rounding order); amplitudes on sampled ranges and the norm must agree to 1e-12, and the norm must not drift."""
    n = pick_qubits(copies=2, cap=30)
    layers = 200
    circ = circuits.random_circuit(n, layers)
    assert len(circ) == 43 * layers or n < 30
    with qcsim_b200.QubitRegister(n, seed=1) as a, qcsim_b200.QubitRegister(n, seed=1) as b:
        a.set_fusion(True)
        a.ApplyGates(circ)
        for g in circ:
            b.ApplyGate(*g)
        na, nb = a.norm2(), b.norm2()
        assert abs(na - nb) <= TOL and abs(na - 1.0) <= 1e-11, (na, nb)
        worst = 0.0
        dim = 1 << n
        for f in ranges(n) + [(dim // 16) * j + 4096 * j for j in range(1, 16)]:
            worst = max(worst, float(np.max(np.abs(a.download(f, 4096) - b.download(f, 4096)))))
        assert worst <= TOL, worst
        st = a.stats()
        print(f"\n[large] fused == unfused at {n} qubits over {layers} layers ({len(circ)} gates): max|d| = {worst:.2e}, "
              f"norm2 {na:.15f} / {nb:.15f}, {st['state_passes']} fused passes vs {len(circ)} gate passes")
