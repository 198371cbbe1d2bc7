"""CPU tests of the engine's host logic: gate classification (classify.h) and the fusion planner
(planner.h), compiled with g++ and driven through ctypes; semantics checked with numpy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import random_state
from qcsim_b200 import _lib, circuits, gates

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "cpp", "hostlogic.so")

OP_NOP, OP_PAIR, OP_DENSE2, OP_DENSE3, OP_DIAG = range(5)


class HlOp(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_ctrl", C.c_int), ("ctrl", C.c_int * 3), ("n_tgt", C.c_int), ("tgt", C.c_int * 3),
                ("m", C.c_double * 128)]


@pytest.fixture(scope="module")
def hl():
    src = os.path.join(HERE, "cpp", "hostlogic.cpp")
    deps = [src] + [os.path.join(HERE, "..", "qcsim_b200", "csrc", f) for f in ("classify.h", "planner.h", "dist_plan.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", src, "-o", SO], check=True)
    lib = C.CDLL(SO)
    lib.hl_plan.restype = C.c_int
    return lib


def pack(circ):
    arr = (_lib.GateStruct * len(circ))()
    for i, (g, q, c1, c2) in enumerate(circ):
        arr[i].nq, arr[i].flags, arr[i].q, arr[i].c1, arr[i].c2 = g.nq, g.flags, q, c1, c2
        flat = np.ascontiguousarray(g.matrix, dtype=np.complex128).view(np.float64).ravel()
        C.memmove(arr[i].m, flat.ctypes.data, flat.nbytes)
    return arr


def op_matrix(op, k):
    return np.frombuffer(op.m, dtype=np.complex128)[:k].copy()


def apply_op(state, op, n):
    """numpy semantics of a classified op (what the CUDA kernels implement)"""
    idx = np.arange(1 << n)
    ok = np.ones(1 << n, dtype=bool)
    for i in range(op.n_ctrl):
        ok &= ((idx >> op.ctrl[i]) & 1).astype(bool)
    out = state.copy()
    if op.kind == OP_NOP:
        return out
    if op.kind == OP_DIAG:
        sel = np.zeros(1 << n, dtype=int)
        for k in range(op.n_tgt):
            sel |= ((idx >> op.tgt[k]) & 1) << k
        table = op_matrix(op, 8)
        out[ok] = state[ok] * table[sel[ok]]
        return out
    if op.kind == OP_PAIR:
        m = op_matrix(op, 4).reshape(2, 2)
        if op.n_tgt == 1:
            t = 1 << op.tgt[0]
            lo = idx[ok & ((idx & t) == 0)]
            hi = lo | t
        else:
            t0, t1 = 1 << op.tgt[0], 1 << op.tgt[1]
            base = idx[ok & ((idx & (t0 | t1)) == 0)]
            lo, hi = base | t0, base | t1
        a, b = state[lo], state[hi]
        out[lo] = m[0, 0] * a + m[0, 1] * b
        out[hi] = m[1, 0] * a + m[1, 1] * b
        return out
    k = 2 if op.kind == OP_DENSE2 else 3
    d = 1 << k
    m = op_matrix(op, d * d).reshape(d, d)
    tm = sum(1 << op.tgt[j] for j in range(k))
    base = idx[ok & ((idx & tm) == 0)]
    offs = [sum((1 << op.tgt[j]) for j in range(k) if (c >> j) & 1) for c in range(d)]
    vin = np.stack([state[base | o] for o in offs])
    vout = m @ vin
    for r, o in enumerate(offs):
        out[base | o] = vout[r]
    return out


def full_matrix_apply(state, g, qs, n):
    """reference semantics: tensor-product operator (SimpleGates.h:211-232 etc.)"""
    idx = np.arange(1 << n)
    k = g.nq
    bits = [1 << q for q in qs[:k]]
    tm = sum(bits)
    base = idx[(idx & tm) == 0]
    offs = [sum(bits[j] for j in range(k) if (c >> j) & 1) for c in range(1 << k)]
    vin = np.stack([state[base | o] for o in offs])
    vout = g.matrix @ vin
    out = state.copy()
    for r, o in enumerate(offs):
        out[base | o] = vout[r]
    return out


def test_classify_semantics_every_gate(hl):
    n = 5
    psi = random_state(n, 3)
    shapes = {}
    for g in gates.all_gate_samples():
        for flagged in (True, False):
            gg = g if flagged else gates.AppliedGate(g.matrix)
            qs = [3, 0, 4]
            arr = pack([(gg, qs[0], qs[1] if g.nq > 1 else 0, qs[2] if g.nq > 2 else 0)])
            op = HlOp()
            hl.hl_classify(arr, C.byref(op))
            want = full_matrix_apply(psi, g, qs, n)
            got = apply_op(psi, op, n)
            assert np.max(np.abs(got - want)) < 1e-15, (g.name, flagged)
            shapes[(g.name, flagged)] = (op.kind, op.n_ctrl, op.n_tgt)
    # structure is recovered from the matrix alone (flag-less == flagged shape)
    for name in ("cx", "ccx", "cz", "ccz", "cp", "swap", "iswap", "cswap", "crx", "rz", "t", "x", "crz"):
        assert shapes[(name, True)] == shapes[(name, False)], name
    assert shapes[("ccx", False)] == (OP_PAIR, 2, 1)
    assert shapes[("cswap", False)] == (OP_PAIR, 1, 2)
    assert shapes[("cp", True)] == (OP_DIAG, 2, 0)      # pure phase on the |11> quarter
    assert shapes[("ccz", True)] == (OP_DIAG, 3, 0)
    assert shapes[("crz", True)] == (OP_DIAG, 1, 1)
    assert shapes[("t", True)] == (OP_DIAG, 1, 0)
    assert shapes[("dec", True)] == (OP_DENSE2, 0, 2)


@pytest.mark.parametrize("n,K,L", [(8, 6, 2), (10, 8, 3), (12, 9, 4), (10, 10, 4)])
def test_planner_preserves_semantics_and_fuses(hl, n, K, L):
    circ = circuits.random_circuit(n, 4, seed=5) + circuits.qft_circuit(n, 1, n - 2) + circuits.ncnot_circuit([0, 1, 2, 3], 4, 5)
    N = len(circ)
    arr = pack(circ)
    fused = (C.c_int * N)()
    tile = (C.c_ulonglong * N)()
    nops = (C.c_int * N)()
    order = (C.c_int * N)()
    ops = (HlOp * N)()
    ns = hl.hl_plan(arr, N, n, K, L, fused, tile, nops, order, ops)
    assert 0 < ns < N / 2, (ns, N)  # really fuses
    flat = list(order[: sum(nops[:ns])])
    # NOPs are dropped; every other op appears exactly once
    live = [i for i in range(N) if ops[i].kind != OP_NOP]
    assert sorted(flat) == live
    # tile constraints
    pos = 0
    for s in range(ns):
        members = flat[pos: pos + nops[s]]
        pos += nops[s]
        assert members == sorted(members)  # program order kept inside a pass
        if fused[s]:
            t = tile[s]
            assert bin(t).count("1") == K and (t & ((1 << L) - 1)) == (1 << L) - 1
            for i in members:
                if ops[i].kind != OP_DIAG:
                    for j in range(ops[i].n_tgt):
                        assert (t >> ops[i].tgt[j]) & 1, "non-diagonal target outside the tile"
    # semantics: planned order == program order
    psi = random_state(n, 9)
    a = psi.copy()
    for i in range(N):
        a = apply_op(a, ops[i], n)
    b = psi.copy()
    for i in flat:
        b = apply_op(b, ops[i], n)
    assert np.max(np.abs(a - b)) < 1e-13


def _match(hl, circ, start=0, min_qubits=4):
    out = (C.c_int * 5)()
    hl.hl_match_qft(pack(circ), len(circ), start, min_qubits, out)
    return list(out)


@pytest.mark.parametrize("n,sq,eq", [(8, 0, 7), (10, 2, 8), (6, 1, 4)])
@pytest.mark.parametrize("swap", [True, False])
@pytest.mark.parametrize("inverse", [False, True])
def test_qft_stream_recognition(hl, n, sq, eq, swap, inverse):
    """the gate stream QCSim's QuantumFourierTransform emits (QuantumFourierTransform.h:35-87) is
    recognised exactly -- range, direction, swap -- and nothing else is"""
    circ = circuits.qft_circuit(n, sq, eq, swap, inverse)
    assert _match(hl, circ) == [len(circ), sq, eq, int(swap), int(inverse)]
    # embedded after other gates: found at its offset, not before
    pre = [(gates.RxGate(0.3), 0, 0, 0), (gates.CNOTGate(), 1, 0, 0)]
    assert _match(hl, pre + circ, start=len(pre))[0] == len(circ)
    assert _match(hl, pre + circ, start=0)[0] == 0
    # one phase off by an ulp-scale amount, or one gate missing: no match of the full transform
    broken = list(circ)
    k = next(i for i, g in enumerate(broken) if g[0].name.startswith("cp"))
    broken[k] = (gates.ControlledPhaseShiftGate(0.123), *broken[k][1:])
    got = _match(hl, broken)
    assert got[0] != len(circ)



def test_qft_recognition_ignores_short_and_foreign_streams(hl):
    assert _match(hl, circuits.qft_circuit(5, 0, 2))[0] == 0          # below min_qubits
    assert _match(hl, circuits.random_circuit(8, 2, seed=3))[0] == 0
    h = (gates.HadamardGate(), 3, 0, 0)
    assert _match(hl, [h, h, h, h, h])[0] == 0
