"""CPU tests of the sharded-register planner (csrc/dist_plan.h).

The planner is pure host code; here it drives SIMULATED shards: every rank's slice is a numpy
array, local steps are applied with the numpy semantics of the classified ops (the same
`apply_op` the fusion-planner test uses), exchange steps move blocks between the slices exactly as
csrc/dist.cu does with ncclSend/ncclRecv.  The result must equal the unsharded circuit.
The second half runs the same thing as two real processes over torch.distributed/gloo.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import random_state
from qcsim_b200 import circuits, gates
from test_planner import HlOp, apply_op, full_matrix_apply, hl, pack  # noqa: F401  (hl is a fixture)

HERE = os.path.dirname(os.path.abspath(__file__))


def plan(hl, circ, n, n_local, rank, phys_of, canonicalize=False):
    N = max(len(circ), 1)
    arr = pack(circ) if circ else None
    max_steps, max_ops = 4 * N + 64, 4 * N + 256
    kind = (C.c_int * max_steps)()
    nops = (C.c_int * max_steps)()
    ex_k = (C.c_int * max_steps)()
    ex_g = (C.c_int * (3 * max_steps))()
    ex_l = (C.c_int * (3 * max_steps))()
    ops = (HlOp * max_ops)()
    po = (C.c_int * 64)(*phys_of, *([0] * (64 - len(phys_of))))
    ns = hl.hl_dist_plan(arr, len(circ), n, n_local, rank, po, int(canonicalize), max_steps, max_ops, kind, nops, ops,
                         ex_k, ex_g, ex_l)
    assert ns >= 0
    steps, o = [], 0
    for s in range(ns):
        if kind[s]:
            steps.append(("x", ex_k[s], list(ex_g[3 * s: 3 * s + ex_k[s]]), list(ex_l[3 * s: 3 * s + ex_k[s]])))
        else:
            steps.append(("l", [ops[o + i] for i in range(nops[s])]))
            # copy the structs out (the array is reused by the caller's next plan)
            steps[-1] = ("l", [HlOp.from_buffer_copy(bytes(x)) for x in steps[-1][1]])
            o += nops[s]
    return steps, list(po[:n])


def sub_block(nl, lpos, v):
    """local indices whose partner bits lpos hold the value v (ascending) -- the strided block k_exchange_swap trades"""
    idx = np.arange(1 << nl)
    keep = np.ones(1 << nl, dtype=bool)
    for j, p in enumerate(lpos):
        keep &= ((idx >> p) & 1) == ((v >> j) & 1)
    return idx[keep]


def exchange_blocks(rank, k, gpos, n_local):
    """(peer, t, a) triples of rank: my amplitudes with partner bits = t <-> the peer's with partner bits = a (dist.cu do_exchange)"""
    a = sum(((rank >> (gpos[j] - n_local)) & 1) << j for j in range(k))
    out = []
    for t in range(1 << k):
        if t == a:
            continue
        peer = rank
        for j in range(k):
            bit = 1 << (gpos[j] - n_local)
            peer = (peer | bit) if (t >> j) & 1 else (peer & ~bit)
        out.append((peer, t, a))
    return out


def run_sharded(hl, circ, n, world, psi):
    """apply `circ` to psi on `world` simulated shards; returns the canonical full state"""
    g = world.bit_length() - 1
    nl = n - g
    shards = [psi[r << nl: (r + 1) << nl].copy() for r in range(world)]
    layout = list(range(n))
    n_exch = 0
    for phase in ("circuit", "canonicalize"):
        plans, layouts = zip(*[plan(hl, circ, n, nl, r, layout, canonicalize=(phase == "canonicalize")) for r in range(world)])
        assert all(l == layouts[0] for l in layouts), "layout evolution must not depend on the rank"
        layout = layouts[0]
        shape = [(s[0], s[1:] if s[0] == "x" else None) for s in plans[0]]
        for p in plans:  # exchange steps are rank-independent and in lock step
            assert [(s[0], s[1:] if s[0] == "x" else None) for s in p] == shape or \
                [s[0] for s in p if s[0] == "x"] == [s[0] for s in plans[0] if s[0] == "x"]
        # local steps may be empty on some ranks (all ops skipped): align by walking exchanges
        cursors = [0] * world
        while True:
            # run local steps up to the next exchange on every rank
            for r in range(world):
                while cursors[r] < len(plans[r]) and plans[r][cursors[r]][0] == "l":
                    for op in plans[r][cursors[r]][1]:
                        assert all(op.tgt[j] < nl for j in range(op.n_tgt)) and all(op.ctrl[j] < nl for j in range(op.n_ctrl))
                        shards[r] = apply_op(shards[r], op, nl)
                    cursors[r] += 1
            if all(cursors[r] >= len(plans[r]) for r in range(world)):
                break
            ex = [plans[r][cursors[r]] for r in range(world)]
            assert all(e == ex[0] for e in ex)
            _, k, gpos, lpos = ex[0]
            assert len(set(lpos)) == k and all(0 <= p < nl for p in lpos) and all(nl <= p < n for p in gpos)
            new = [s.copy() for s in shards]
            for r in range(world):
                for peer, t, a in exchange_blocks(r, k, gpos, nl):
                    new[r][sub_block(nl, lpos, t)] = shards[peer][sub_block(nl, lpos, a)]
            shards = new
            n_exch += 1
            for r in range(world):
                cursors[r] += 1
        if phase == "circuit":
            circ_layout = layout
            circ = []
    assert layout == list(range(n))
    return np.concatenate(shards), n_exch, circ_layout


def run_unsharded(circ, n, psi):
    out = psi.copy()
    for g, q, c1, c2 in circ:
        out = full_matrix_apply(out, g, [q, c1, c2], n)
    return out


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n", [6, 9])
def test_sharded_random_circuit(hl, n, world):
    circ = circuits.random_circuit(n, 5, seed=11 + world)
    psi = random_state(n, 21)
    got, n_exch, _ = run_sharded(hl, circ, n, world, psi)
    want = run_unsharded(circ, n, psi)
    assert np.max(np.abs(got - want)) < 1e-13
    assert n_exch > 0


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_every_gate_kind_on_global_qubits(hl, world):
    n = 7
    psi = random_state(n, 5)
    circ = []
    rng = np.random.default_rng(3)
    for g in gates.all_gate_samples():
        for flagged in (True, False):
            qs = [int(x) for x in rng.permutation(n)[:3]]
            qs[int(rng.integers(0, g.nq))] = n - 1 - int(rng.integers(0, world.bit_length() - 1))  # force a global qubit
            if len(set(qs[:g.nq])) < g.nq:
                continue
            gg = g if flagged else gates.AppliedGate(g.matrix)
            circ.append((gg, qs[0], qs[1] if g.nq > 1 else 0, qs[2] if g.nq > 2 else 0))
    got, _, _ = run_sharded(hl, circ, n, world, psi)
    want = run_unsharded(circ, n, psi)
    assert np.max(np.abs(got - want)) < 1e-13


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_qft_exchanges_are_minimal(hl, world):
    n = 10
    psi = random_state(n, 8)
    circ = circuits.qft_circuit(n)
    got, n_exch, layout = run_sharded(hl, circ, n, world, psi)
    want = run_unsharded(circ, n, psi)
    assert np.max(np.abs(got - want)) < 1e-13
    # the global qubits need non-diagonal work only for their Hadamards; the controlled
    # phases are diagonal and the final SWAPs are relabellings -> two all-to-all phases for the
    # circuit (out and back) + what canonicalisation needs to undo the bit reversal
    assert n_exch <= 2 + 2
    assert layout != list(range(n))  # the bit reversal stayed virtual until canonicalisation


def test_swap_only_circuit_moves_no_data(hl):
    n, world = 8, 4
    circ = [(gates.SwapGate(), 0, 7, 0), (gates.SwapGate(), 1, 6, 0), (gates.SwapGate(), 2, 5, 0)]
    steps, layout = plan(hl, circ, n, n - 2, 1, list(range(n)))
    assert steps == []
    assert layout[0] == 7 and layout[7] == 0 and layout[1] == 6 and layout[2] == 5


# ---- two real processes over gloo ---------------------------------------------------------------

WORKER = r"""
import os, sys, ctypes as C
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "tests")); sys.path.insert(0, sys.argv[1])
from conftest import random_state
from qcsim_b200 import circuits
import test_planner, test_dist_plan
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
hl = C.CDLL(test_planner.SO); hl.hl_plan.restype = C.c_int
n = 9; nl = n - (world.bit_length() - 1)
psi = random_state(n, 4)
circ = circuits.random_circuit(n, 4, seed=2) + circuits.qft_circuit(n)
mine = psi[rank << nl: (rank + 1) << nl].copy()
layout = list(range(n))
for phase in (circ, None):
    steps, layout = test_dist_plan.plan(hl, phase or [], n, nl, rank, layout, canonicalize=phase is None)
    for s in steps:
        if s[0] == "l":
            for op in s[1]:
                mine = test_planner.apply_op(mine, op, nl)
        else:
            _, k, gpos, lpos = s
            new = mine.copy()
            reqs, bufs = [], []
            for peer, t, a in test_dist_plan.exchange_blocks(rank, k, gpos, nl):
                send = torch.from_numpy(np.ascontiguousarray(mine[test_dist_plan.sub_block(nl, lpos, t)]).view(np.float64).copy())
                recv = torch.empty_like(send)
                reqs += [dist.isend(send, peer), dist.irecv(recv, peer)]
                bufs.append((t, recv, send))
            for r in reqs:
                r.wait()
            for t, recv, _ in bufs:
                new[test_dist_plan.sub_block(nl, lpos, t)] = recv.numpy().view(np.complex128)
            mine = new
want = test_dist_plan.run_unsharded(circ, n, psi)[rank << nl: (rank + 1) << nl]
err = float(np.max(np.abs(mine - want)))
t = torch.tensor([err], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("MAXERR", float(t.item()))
dist.destroy_process_group()
"""


def test_two_process_gloo(hl, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    root = os.path.dirname(HERE)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29731", str(script), root], capture_output=True, text=True,
                         env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    line = [l for l in res.stdout.splitlines() if l.startswith("MAXERR")][0]
    assert float(line.split()[1]) < 1e-13
