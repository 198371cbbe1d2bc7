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


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_round_scheduler_invariants(hl, seed):
    """schedule_rounds: every op lands in exactly one round whose 3 register bits hold its
    non-diagonal targets and whose <= 2 variant qubits + register bits hold everything it merely
    looks at; executing round by round equals program order; the item-index bit map is a bijection
    onto the non-round tile bits with conflict-free low lanes."""
    n, K = 12, 12
    circ = [g for g in circuits.random_circuit(n, 3, seed=seed)] + circuits.ncnot_circuit([0, 1, 2, 3], 4, 5)
    N = len(circ)
    arr = pack(circ)
    ops = (HlOp * N)()
    for i in range(N):
        hl.hl_classify(C.byref(arr, i * C.sizeof(_lib.GateStruct)), C.byref(ops[i]))
    R = N + 4
    rbits, nvar, vq, nops, order = (C.c_int * (3 * R))(), (C.c_int * R)(), (C.c_int * (3 * R))(), (C.c_int * R)(), (C.c_int * N)()
    item_bits = (C.c_int * (9 * R))()
    hl.hl_rounds.restype = C.c_int
    nr = hl.hl_rounds(arr, N, (1 << K) - 1, 2, rbits, nvar, vq, nops, order, item_bits)
    assert 0 < nr < N
    flat = list(order[: sum(nops[:nr])])
    live = [i for i in range(N) if ops[i].kind != OP_NOP]
    assert sorted(flat) == live
    pos = 0
    for r in range(nr):
        rb = set(rbits[3 * r: 3 * r + 3])
        assert len(rb) == 3
        var = set(vq[3 * r + j] for j in range(nvar[r]))
        assert nvar[r] <= 2 and not (var & rb)
        for i in flat[pos: pos + nops[r]]:
            op = ops[i]
            nd = set(op.tgt[j] for j in range(op.n_tgt)) if op.kind != OP_DIAG else set()
            dg = set(op.ctrl[j] for j in range(op.n_ctrl)) | (set(op.tgt[j] for j in range(op.n_tgt)) if op.kind == OP_DIAG else set())
            assert nd <= rb, "non-diagonal target outside the round's register bits"
            assert dg <= (rb | var), "control / selector that is neither a register bit nor a variant bit"
        pos += nops[r]
        ib = list(item_bits[9 * r: 9 * r + K - 3])
        assert sorted(ib) == sorted(set(range(K)) - rb)                      # bijection onto the other tile bits
        lanes_free = set(range(K)) - rb - var                                # variant bits are kept off the lane bits
        assert len({b % 3 for b in ib[:3]}) == min(3, len({b % 3 for b in lanes_free}))  # quarter-warp: distinct bank groups when possible
        for v in var:                                                         # variant bits are warp-uniform item bits
            assert ib.index(v) in (5, 6, 7)
    psi = random_state(n, 9)
    a = psi.copy()
    for i in range(N):
        a = apply_op(a, ops[i], n)
    b = psi.copy()
    for i in flat:
        b = apply_op(b, ops[i], n)
    assert np.max(np.abs(a - b)) < 1e-13


@pytest.mark.parametrize("seed,tile", [(1, 0xFFF), (2, 0x3F0F), (3, 0xFFF)])
def test_round_matrices_reproduce_the_circuit(hl, seed, tile):
    """What k_tile_pass computes, emulated in numpy: for every round, every amplitude group is
    multiplied by the 8x8 matrix selected by its variant bits.  The product over the rounds must equal
    the gate-by-gate circuit -- this pins schedule_rounds + build_round_matrices (all gate kinds,
    controls inside / outside the round, diagonal selectors) without a GPU."""
    n = 14
    tq = [q for q in range(n) if (tile >> q) & 1]
    rng = np.random.default_rng(seed)
    # gates whose non-diagonal targets lie in the tile; controls / selectors anywhere
    circ = []
    samples = gates.all_gate_samples()
    while len(circ) < 60:
        g = samples[int(rng.integers(0, len(samples)))]
        qs = [int(x) for x in rng.permutation(n)[:3]]
        gg = g if rng.integers(0, 2) else gates.AppliedGate(g.matrix)
        circ.append((gg, qs[0], qs[1] if g.nq > 1 else 0, qs[2] if g.nq > 2 else 0))
    arr = pack(circ)
    N = len(circ)
    ops = (HlOp * N)()
    keep = []
    for i in range(N):
        hl.hl_classify(C.byref(arr, i * C.sizeof(_lib.GateStruct)), C.byref(ops[i]))
        op = ops[i]
        nd = [op.tgt[j] for j in range(op.n_tgt)] if op.kind != OP_DIAG else []
        if all(q in tq for q in nd):
            keep.append(i)
    circ = [circ[i] for i in keep]
    arr = pack(circ)
    N = len(circ)
    assert N > 20
    R = N + 4
    rbits, nvar, vq = (C.c_int * (3 * R))(), (C.c_int * R)(), (C.c_int * (3 * R))()
    max_mats = 8 * R
    mats = (C.c_double * (128 * max_mats))()
    hl.hl_round_matrices.restype = C.c_int
    nr = hl.hl_round_matrices(arr, N, tile, 3, rbits, nvar, vq, mats, max_mats)
    assert 0 < nr < N
    M = np.frombuffer(mats, dtype=np.complex128).reshape(max_mats, 8, 8)
    psi = random_state(n, 5)
    want = psi.copy()
    for g, q, c1, c2 in circ:
        want = full_matrix_apply(want, g, [q, c1, c2], n)
    got = psi.copy()
    idx = np.arange(1 << n)
    used = 0
    for r in range(nr):
        rq = [tq[rbits[3 * r + j]] for j in range(3)]          # register bit j <-> qubit
        var = [vq[3 * r + j] for j in range(nvar[r])]
        tm = sum(1 << q for q in rq)
        base = idx[(idx & tm) == 0]
        offs = [sum((1 << rq[j]) for j in range(3) if (x >> j) & 1) for x in range(8)]
        vin = np.stack([got[base | o] for o in offs])            # 8 x groups
        vidx = np.zeros(base.shape, dtype=int)
        for j, q in enumerate(var):
            vidx |= ((base >> q) & 1) << j
        vout = np.einsum("gij,jg->ig", M[used + vidx], vin)
        for x, o in enumerate(offs):
            got[base | o] = vout[x]
        used += 1 << nvar[r]
    assert np.max(np.abs(got - want)) < 1e-13


# ---- TMA-staged pass (tile_pipe.cuh): geometry + the kernel's index arithmetic, emulated -------------------
def _geometry(hl, tile_mask, n_local):
    out = (C.c_int * 64)()
    for i in range(64):
        out[i] = 0
    hl.hl_tma_geometry.restype = C.c_int
    ok = hl.hl_tma_geometry(C.c_ulonglong(tile_mask), n_local, out)
    if not ok:
        return None
    return _parse_geometry(list(out))


_CHAINED = []  # chained rounds seen by each emulated pass (checked after all of them ran)


def _parse_geometry(o):
    g = {"n_dims": o[0], "n_enum": o[1], "box_log2": o[2], "dim_lo": o[3:8], "dim_bits": o[8:13], "box_bits": o[13:18],
         "enum_pos": o[18:27][: o[1]], "slot_qubit": o[27:39]}
    return g


def _tswz(s):
    return s ^ ((s >> 3) & 7)


def _tma_ops(g, gbase):
    """(smem slot base, [global index per box element in TMA linear order]) of every op of the tile at gbase:
    coordinates exactly as the producer warp computes them, box traversal as the hardware does (dim 0 fastest)"""
    ops = []
    for e in range(1 << g["n_enum"]):
        idx = gbase
        for j, p in enumerate(g["enum_pos"]):
            idx |= ((e >> j) & 1) << p
        coord = [0] + [(idx >> g["dim_lo"][d]) & ((1 << g["dim_bits"][d]) - 1) for d in range(1, 5)]
        elems = []
        box = [1 << b for b in g["box_bits"]]
        for b4 in range(box[4]):
            for b3 in range(box[3]):
                for b2 in range(box[2]):
                    for b1 in range(box[1]):
                        for b0 in range(box[0]):
                            c = [coord[0] + b0, coord[1] + b1, coord[2] + b2, coord[3] + b3, coord[4] + b4]
                            # element address = sum coordinate * stride (stride of dim d = 2^dim_lo[d] amplitudes)
                            elems.append(sum(c[d] << g["dim_lo"][d] for d in range(5)))
        ops.append((e << g["box_log2"], elems))
    return ops


def _mask(bits):
    return sum(1 << b for b in bits)


@pytest.mark.parametrize("n_local,tile_bits", [
    (12, range(12)), (14, [0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13]), (20, [0, 1, 2, 3, 5, 7, 9, 11, 13, 15, 17, 19]),
    (30, [0, 1, 2, 3, 22, 23, 24, 25, 26, 27, 28, 29]), (30, [0, 1, 2, 5, 9, 12, 13, 14, 15, 16, 17, 18]),
    (33, [0, 1, 2, 3, 10, 12, 14, 16, 18, 20, 30, 32]), (30, [0, 1, 2, 21, 22, 23, 24, 25, 26, 27, 28, 29]),
    (24, [0, 1, 2, 3, 6, 7, 10, 11, 14, 15, 20, 21]),
    (11, range(11)), (30, [0, 1, 2, 4, 7, 11, 12, 20, 25, 28, 29]), (33, [0, 1, 2, 24, 25, 26, 27, 28, 30, 31, 32]),
    (30, [0, 1, 2, 3, 5, 7, 9, 11, 13, 15, 17]),
])
def test_tma_tile_geometry_addresses_every_tile_amplitude_once(hl, n_local, tile_bits):
    tile_mask = _mask(tile_bits)
    """tma_tile_geometry: the TMA ops of a tile (coordinates as in the producer warp) touch exactly the tile's 2^12
    amplitudes, each once, and shared-memory slot bit j holds index bit slot_qubit[j]."""
    k = bin(tile_mask).count("1")
    assert k in (11, 12)
    g = _geometry(hl, tile_mask, n_local)
    assert g is not None
    # tensor-map constraints (cuTensorMapEncodeTiled): box <= 256, dims partition the index bits, strides < 2^40
    assert g["dim_lo"][0] == 0 and g["dim_bits"][0] == 3 and g["box_bits"][0] == 3 and g["dim_lo"][1] == 3
    for d in range(1, g["n_dims"]):
        nxt = g["dim_lo"][d + 1] if d + 1 < g["n_dims"] else n_local
        assert g["dim_lo"][d] + g["dim_bits"][d] == nxt and g["box_bits"][d] <= min(8, g["dim_bits"][d])
        assert (16 << g["dim_lo"][d]) < (1 << 40)
    assert g["box_log2"] >= 6 and g["box_log2"] + g["n_enum"] == k
    assert sorted(g["slot_qubit"][:k]) == [q for q in range(64) if (tile_mask >> q) & 1]
    rng = np.random.default_rng(1)
    free = [q for q in range(n_local) if not (tile_mask >> q) & 1]
    for _ in range(3):
        t = int(rng.integers(0, 1 << len(free))) if free else 0
        gbase = sum(((t >> j) & 1) << q for j, q in enumerate(free))
        seen = {}
        for slot_base, elems in _tma_ops(g, gbase):
            for off, idx in enumerate(elems):
                seen[slot_base + off] = idx
        assert sorted(seen) == list(range(1 << k))
        for slot, idx in seen.items():
            want = gbase | sum(((slot >> j) & 1) << g["slot_qubit"][j] for j in range(k))
            assert idx == want, (slot, idx, want)


@pytest.mark.parametrize("layout_search", [1, 0])
@pytest.mark.parametrize("seed,n,tile_bits", [(1, 13, range(11)), (2, 14, [0, 1, 2, 3, 6, 7, 8, 9, 10, 12, 13]),
                                              (3, 15, [0, 1, 2, 3, 5, 6, 8, 9, 12, 13, 14]), (4, 16, [0, 1, 2, 8, 9, 10, 11, 12, 13, 14, 15])])
def test_pipe_kernel_data_path_emulated(hl, seed, n, tile_bits, layout_search):
    tile_mask = _mask(tile_bits)
    """k_tile_pipe emulated in numpy from exactly what launch_pass_pipe gives it: TMA ops into swizzled slots, rounds in
    the kernel's MMA mapping (lane (g, t): B fragment = amplitudes t, 4 + t of item g; D fragment = output amplitude g of
    items 2t, 2t + 1; warp-uniform variant selection), TMA stores.  Equals the gate-by-gate circuit."""
    K = 11                                                                # kPipeTileBits
    assert bin(tile_mask).count("1") == K
    tq = [q for q in range(n) if (tile_mask >> q) & 1]
    rng = np.random.default_rng(seed)
    circ = []
    samples = gates.all_gate_samples()
    while len(circ) < 50:
        g = samples[int(rng.integers(0, len(samples)))]
        qs = [int(x) for x in rng.permutation(n)[:3]]
        gg = g if rng.integers(0, 2) else gates.AppliedGate(g.matrix)
        circ.append((gg, qs[0], qs[1] if g.nq > 1 else 0, qs[2] if g.nq > 2 else 0))
    arr = pack(circ)
    keep = []
    for i in range(len(circ)):
        op = HlOp()
        hl.hl_classify(C.byref(arr, i * C.sizeof(_lib.GateStruct)), C.byref(op))
        nd = [op.tgt[j] for j in range(op.n_tgt)] if op.kind != OP_DIAG else []
        if all(q in tq for q in nd):
            keep.append(i)
    circ = [circ[i] for i in keep]
    arr = pack(circ)
    N = len(circ)
    assert N > 15
    base_geom = _geometry(hl, tile_mask, n)
    assert base_geom is not None
    R = N + 4
    desc = (C.c_uint * (6 * R))()
    max_mats = 8 * R
    mats = (C.c_double * (128 * max_mats))()
    hl.hl_pipe_pass.restype = C.c_int
    geom_out = (C.c_int * 64)()
    nr = hl.hl_pipe_pass(arr, N, C.c_ulonglong(tile_mask), n, desc, mats, max_mats, geom_out, layout_search)
    assert 0 < nr < N
    geom = _parse_geometry(list(geom_out))                                # dimension order chosen against bank conflicts
    assert sorted(geom["slot_qubit"][:K]) == sorted(base_geom["slot_qubit"][:K]) and geom["slot_qubit"][:3] == [0, 1, 2]
    assert geom["box_log2"] == base_geom["box_log2"] and geom["enum_pos"] == base_geom["enum_pos"]
    if not layout_search:
        assert geom == base_geom
    M = np.frombuffer(mats, dtype=np.complex128).reshape(max_mats, 8, 8)
    psi = random_state(n, 5)
    want = psi.copy()
    for g, q, c1, c2 in circ:
        want = full_matrix_apply(want, g, [q, c1, c2], n)
    got = psi.copy()
    free = [q for q in range(n) if q not in tq]
    lane = np.arange(32)
    g, tq4 = lane >> 2, lane & 3
    degrees = []
    prev_chain, n_chained = None, 0
    for t in range(1 << len(free)):
        gbase = sum(((t >> j) & 1) << q for j, q in enumerate(free))
        ops = _tma_ops(geom, gbase)
        smem = np.zeros(1 << K, dtype=np.complex128)
        for slot_base, elems in ops:                                     # UTMALDG with the 128 B swizzle
            for off, idx in enumerate(elems):
                smem[_tswz(slot_base + off)] = got[idx]
        for r in range(nr):
            rb, tb0, tb1, tb2, var, mat_off = [int(x) for x in desc[6 * r: 6 * r + 6]]
            tb = [tb0, tb1, tb2]
            rbit = [1 << ((rb >> (8 * j)) & 31) for j in range(3)]
            ib = [1 << ((tb[j >> 2] >> (8 * (j & 3))) & 31) for j in range(K - 3)]
            assert sorted(rbit + ib) == [1 << j for j in range(K)]       # register + item bits = all slot bits
            nvar = var & 0x7F
            chain_next = (var >> 7) & 1                                  # next round: same warp bits, no group barrier
            touched_d = np.zeros(2 << K, dtype=int)
            warp_slots = []
            new_smem = smem.copy()
            for warp in range(8):                                        # the 8 warps of one consumer group
                hi = sum(ib[5 + j] for j in range(3) if (warp >> j) & 1)
                vidx = 0
                for j in range(nvar):
                    e = (var >> (8 + 8 * j)) & 0xFF
                    bit = ((gbase >> (e >> 1)) & 1) if (e & 1) else ((hi >> (e >> 1)) & 1)
                    vidx |= bit << j
                Mv = M[mat_off + vidx]
                ld0 = _tswz(hi | np.where(g & 1, ib[0], 0) | np.where(g & 2, ib[1], 0) | np.where(g & 4, ib[2], 0)
                            | np.where(tq4 & 1, rbit[0], 0) | np.where(tq4 & 2, rbit[1], 0))
                st0 = _tswz(hi | np.where(tq4 & 1, ib[1], 0) | np.where(tq4 & 2, ib[2], 0) | np.where(g & 1, rbit[0], 0)
                            | np.where(g & 2, rbit[1], 0) | np.where(g & 4, rbit[2], 0))
                x_hi, x_i0, x_p0, x_p1 = _tswz(rbit[2]), _tswz(ib[0]), _tswz(ib[3]), _tswz(ib[4])
                lq, sq = tq4 & 1, g & 1                                  # half fetched / stored first by the lane
                td = smem.view(np.float64)                               # td[2 * slot + part]
                tdw = new_smem.view(np.float64)
                mine = set()
                for p in range(4):
                    a0 = ld0 ^ (x_p0 if p & 1 else 0) ^ (x_p1 if p & 2 else 0)
                    mine |= set(a0.tolist()) | set((a0 ^ x_hi).tolist())
                    loads = [2 * a0 + lq, 2 * (a0 ^ x_hi) + lq, 2 * a0 + (lq ^ 1), 2 * (a0 ^ x_hi) + (lq ^ 1)]   # 4 LDS.64, K-block j
                    for addr in loads:
                        touched_d[addr] += 1
                    bfr = [td[addr] for addr in loads]
                    # what the lanes hold: first / second fetched half of amplitudes tq and 4 + tq of item g.  With
                    # v'_a = (-i)^(a & 1) v_a:  Re v' = first half,  Im v' = second half, negated on odd k-lanes
                    x = np.zeros((8, 8))                                 # [item][amplitude]
                    y = np.zeros((8, 8))
                    for j in range(2):
                        amp_i = tq4 + 4 * j
                        x[g, amp_i] = bfr[j]
                        y[g, amp_i] = np.where(lq == 1, -bfr[2 + j], bfr[2 + j])
                    # M'[o][a] = (-i)^(o & 1) i^(a & 1) M[o][a] = P + iQ;  three real products (6 DMMA per panel):
                    # S = P (x + y),  Re o' = S - (P + Q) y,  Im o' = S + (Q - P) x
                    oo, aa = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
                    Mp = Mv * ((-1j) ** (oo & 1)) * (1j ** (aa & 1))
                    P, Q = Mp.real, Mp.imag
                    S = (x + y) @ P.T
                    re_o = S + y @ (-(P + Q)).T                          # [item][o]
                    im_o = S + x @ (Q - P).T
                    # o'_o = (-i)^(o & 1) o_o: the first stored half is Re o', the second is Im o', negated on odd rows
                    odd = (np.arange(8) & 1)[None, :]
                    out = np.where(odd == 1, -im_o + 1j * re_o, re_o + 1j * im_o)   # back to o (checked against the kernel's stores below)
                    s0 = st0 ^ (x_p0 if p & 1 else 0) ^ (x_p1 if p & 2 else 0)
                    o0, o1 = out[2 * tq4, g], out[2 * tq4 + 1, g]        # D fragments: row o = g, columns 2 tq, 2 tq + 1
                    first0 = np.where(sq == 1, o0.imag, o0.real)
                    first1 = np.where(sq == 1, o1.imag, o1.real)
                    second0 = np.where(sq == 1, o0.real, o0.imag)
                    second1 = np.where(sq == 1, o1.real, o1.imag)
                    stores = [(2 * s0 + sq, first0), (2 * (s0 ^ x_i0) + sq, first1), (2 * s0 + (sq ^ 1), second0), (2 * (s0 ^ x_i0) + (sq ^ 1), second1)]
                    for addr, val in stores:
                        tdw[addr] = val
                    assert sorted(np.concatenate([a for a, _ in stores]).tolist()) == sorted(np.concatenate(loads).tolist())
                    if t == 0:                                           # half-warp 8-byte columns of the LDS.64 / STS.64
                        for addr in loads + [a for a, _ in stores]:
                            degrees.append(max(max(np.bincount(addr[q * 16: q * 16 + 16] & 15, minlength=16)) for q in range(2)))
                warp_slots.append(mine)
            assert np.all(touched_d == 1)                                # the round reads every 8-byte half exactly once
            if t == 0:
                if prev_chain is not None:                               # chained rounds: every warp stays on its own 256 slots
                    assert prev_chain == warp_slots, r
                    n_chained += 1
                prev_chain = warp_slots if chain_next else None
            smem = new_smem
        for slot_base, elems in ops:                                     # UTMASTG
            for off, idx in enumerate(elems):
                got[idx] = smem[_tswz(slot_base + off)]
    assert np.max(np.abs(got - want)) < 1e-13
    # bank conflicts of the fragment accesses: the planner picks the register-bit order and item bits 0..2 to dodge them;
    # what remains is forced by register bits that sit above slot bit 5 (the TMA swizzle does not fold those)
    assert max(degrees) <= 2 and np.mean(degrees) < 1.7, (max(degrees), np.mean(degrees))
    _CHAINED.append(n_chained)


def test_some_emulated_rounds_were_chained():
    """runs after test_pipe_kernel_data_path_emulated: the chained (warp-local, barrier-free) round hand-over was exercised"""
    assert _CHAINED and sum(_CHAINED) >= 1, _CHAINED


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_lazy_flush_for_a_qubit_probability(hl, seed):
    """GetQubitProbability(q) only needs the queued gates that act non-diagonally on q and what those depend on
    (planner.h: split_queue_for_diagonal_observable): running the needed gates first and the rest later is the same
    circuit, and P(q) after the needed gates alone is already the final P(q)."""
    n = 7
    rng = np.random.default_rng(seed)
    samples = [g for g in gates.all_gate_samples() if g.nq <= 2] + [g for g in gates.all_gate_samples() if g.nq == 3][:4]
    circ = []
    for _ in range(40):
        g = samples[int(rng.integers(0, len(samples)))]
        qs = [int(x) for x in rng.permutation(n)[:3]]
        gg = g if rng.integers(0, 2) else gates.AppliedGate(g.matrix)
        circ.append((gg, qs[0], qs[1] if g.nq > 1 else 0, qs[2] if g.nq > 2 else 0))
    arr = pack(circ)
    psi0 = random_state(n, seed)
    full = psi0.copy()
    for g, q, c1, c2 in circ:
        full = full_matrix_apply(full, g, [q, c1, c2], n)
    kept_any = False
    for q in range(n):
        needed = (C.c_int * len(circ))()
        hl.hl_split_for_observable(arr, len(circ), C.c_ulonglong(1 << q), needed)
        need = [i for i in range(len(circ)) if needed[i]]
        rest = [i for i in range(len(circ)) if not needed[i]]
        kept_any = kept_any or len(rest) > 0
        part = psi0.copy()
        for i in need:
            g, a, c1, c2 = circ[i]
            part = full_matrix_apply(part, g, [a, c1, c2], n)
        idx = np.arange(1 << n)
        p_part = float(np.sum(np.abs(part[(idx >> q) & 1 == 1]) ** 2))
        p_full = float(np.sum(np.abs(full[(idx >> q) & 1 == 1]) ** 2))
        assert abs(p_part - p_full) < 1e-12, (q, p_part, p_full)
        for i in rest:
            g, a, c1, c2 = circ[i]
            part = full_matrix_apply(part, g, [a, c1, c2], n)
        assert np.max(np.abs(part - full)) < 1e-12, q
    assert kept_any
