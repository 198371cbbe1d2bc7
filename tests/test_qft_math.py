"""CPU check of the algebra behind the radix-8 QFT passes (csrc/qft_kernels.cuh).

The kernel replaces, for a group of up to 3 adjacent qubits, every gate of QCSim's QFT whose target
lies in the group by [butterfly: H CP CP H CP H] x [diagonal factor P^rev(x), P = exp(+-i pi R / 2^top)]
with R = value of all lower transform qubits.  Here the same decomposition is restated in numpy and
compared with the gate-by-gate circuit of QuantumFourierTransform.h:35-87 -- no GPU involved, so a
mistake in the derivation (group order, twiddle exponent, inverse direction) is caught on the CPU."""
import numpy as np
import pytest

from conftest import random_state
from qcsim_b200 import circuits
from test_planner import full_matrix_apply


def radix8_qft(psi, n, sq, eq, inverse):
    idx = np.arange(1 << n)
    s = 1.0 / np.sqrt(2.0)
    sign = -1.0 if inverse else 1.0
    ph2 = complex(np.cos(sign * np.pi / 2), np.sin(sign * np.pi / 2))
    ph4 = complex(np.cos(sign * np.pi / 4), np.sin(sign * np.pi / 4))
    groups = []
    top = eq
    while top >= sq:
        size = min(3, top - sq + 1)
        groups.append((top, size))
        top -= size
    if inverse:
        groups.reverse()
    out = psi.copy()
    for top, G in groups:
        c0 = top - G + 1
        N = 1 << G
        gm = ((1 << G) - 1) << c0
        base = idx[(idx & gm) == 0]
        v = [out[base | (x << c0)] for x in range(N)]
        R = base & ((1 << c0) - 1) & ~((1 << sq) - 1)
        P = np.exp(sign * 1j * np.pi * R / float(1 << top))

        def hadamard(bit):
            for x in range(N):
                if not (x >> bit) & 1:
                    a, b = v[x], v[x | (1 << bit)]
                    v[x], v[x | (1 << bit)] = s * (a + b), s * (a - b)

        def cphase(t, c, ph):
            for x in range(N):
                if (x >> t) & 1 and (x >> c) & 1:
                    v[x] = v[x] * ph

        def twiddle():
            for x in range(1, N):
                rev = sum(1 << (G - 1 - b) for b in range(G) if (x >> b) & 1)
                v[x] = v[x] * P ** rev

        fwd = {3: [("h", 2), ("cp", 2, 1, ph2), ("cp", 2, 0, ph4), ("h", 1), ("cp", 1, 0, ph2), ("h", 0)],
               2: [("h", 1), ("cp", 1, 0, ph2), ("h", 0)], 1: [("h", 0)]}[G]
        seq = fwd if not inverse else list(reversed(fwd))
        if inverse:
            twiddle()
        for step in seq:
            if step[0] == "h":
                hadamard(step[1])
            else:
                cphase(step[1], step[2], step[3])
        if not inverse:
            twiddle()
        for x in range(N):
            out[base | (x << c0)] = v[x]
    return out


@pytest.mark.parametrize("n,sq,eq", [(7, 0, 6), (9, 0, 8), (10, 2, 9), (10, 0, 7), (8, 3, 4), (11, 1, 10), (6, 5, 5)])
@pytest.mark.parametrize("inverse", [False, True])
def test_radix8_decomposition_equals_gate_by_gate_qft(n, sq, eq, inverse):
    psi = random_state(n, 13)
    want = psi.copy()
    for g, q, c1, c2 in circuits.qft_circuit(n, sq, eq, False, inverse):
        want = full_matrix_apply(want, g, [q, c1, c2], n)
    got = radix8_qft(psi, n, sq, eq, inverse)
    assert np.max(np.abs(got - want)) < 1e-13


def test_full_qft_is_the_inverse_dft_up_to_qubit_reversal():
    n = 9
    psi = random_state(n, 2)
    got = radix8_qft(psi, n, 0, n - 1, False)
    rev = np.array([int(format(i, f"0{n}b")[::-1], 2) for i in range(1 << n)])
    assert np.max(np.abs(got[rev] - np.sqrt(1 << n) * np.fft.ifft(psi))) < 1e-12
