"""Golden-vector plumbing shared by the CPU tests (port vs fixtures) and the GPU tests.

A fixture (tests/golden/<name>.npz) holds: a JSON `spec` describing how to rebuild the input
(seeded start state + circuit generator arguments), and the outputs the COMPILED REFERENCE
produced for it (amplitudes, measurement outcomes for injected draws).  make_golden.py writes
them; nothing here needs /root/reference.
"""
import glob
import json
import os

import numpy as np

from conftest import draws, random_state
from qcsim_b200 import circuits, gates

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def all_gates_circuit(n: int):
    """Every reference gate class once, flagged and flag-less, on rotating qubit choices."""
    out = []
    k = 0
    for g in gates.all_gate_samples():
        qs = [(k + (0, 2, 1)[j]) % n for j in range(g.nq)]
        k += 1
        args = qs + [0] * (3 - g.nq)
        out.append((g, *args))
        out.append((gates.AppliedGate(g.matrix), *args))
    return out


def build_circuit(spec):
    kind = spec["circuit"]
    n = spec["n"]
    if kind == "all_gates":
        return all_gates_circuit(n)
    if kind == "random":
        return circuits.random_circuit(n, spec["layers"], spec.get("seed", circuits.RANDOM_CIRCUIT_SEED))
    if kind == "qft":
        return circuits.qft_circuit(n, spec.get("sq", 0), spec.get("eq"), spec.get("swap", True), spec.get("inverse", False))
    if kind == "qft_iqft":
        return circuits.qft_circuit(n) + circuits.qft_circuit(n, inverse=True)
    if kind == "grover":
        return circuits.grover_gates_circuit(spec["n_search"], spec["marked"])
    if kind == "none":
        return []
    raise KeyError(kind)


def start_state(spec):
    n = spec["n"]
    s = spec.get("start", "zero")
    if s == "zero":
        v = np.zeros(1 << n, dtype=np.complex128)
        v[0] = 1
        return v
    if s == "basis":
        v = np.zeros(1 << n, dtype=np.complex128)
        v[spec["basis"]] = 1
        return v
    if s == "random":
        return random_state(n, spec.get("state_seed", 7))
    raise KeyError(s)


def load_all():
    out = {}
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        z = np.load(p, allow_pickle=False)
        case = {k: z[k] for k in z.files}
        case["spec"] = json.loads(str(case["spec"]))
        case["n"] = case["spec"]["n"]
        out[os.path.splitext(os.path.basename(p))[0]] = case
    return out


def run_case(sim, case):
    """Drive `sim` (an oracle or a GPU adapter with the oracle interface) through the case and
    return the observed outputs in the same layout as the fixture."""
    spec = case["spec"]
    n = spec["n"]
    got = {}
    sim.set_state(start_state(spec))
    sim.apply_circuit(build_circuit(spec))
    st = sim.state()
    stride = spec.get("stride", 1)
    got["amps"] = st[::stride].copy()
    got["norm2"] = np.array(sim.norm2())
    meas = spec.get("measure")
    if meas:
        ps = draws(meas["count"], meas.get("seed", 42))
        base = st
        outs = []
        for p in ps:
            row = [sim.measure_all_nocollapse(p)]
            for (a, b) in meas.get("ranges", []):
                row.append(sim.measure_nocollapse(a, b, p))
            outs.append(row)
        got["outcomes"] = np.array(outs, dtype=np.uint64)
        if "collapse_range" in meas:
            a, b = meas["collapse_range"]
            sim.set_state(base)
            got["collapse_outcome"] = np.array(sim.measure(a, b, ps[0]), dtype=np.uint64)
            got["collapse_amps"] = sim.state()[::stride].copy()
            sim.set_state(base)
        got["qubit_prob"] = np.array([sim.qubit_probability(q) for q in range(n)])
    return got


def check_case(sim, case, exact=False, tol=1e-12):
    got = run_case(sim, case)
    for key, val in got.items():
        want = case[key]
        if val.dtype.kind in "ui":
            assert np.array_equal(val, want), key
        elif exact and key == "amps":
            assert np.array_equal(val, want), key
        elif exact and key == "collapse_amps":
            # collapse norm: the reference's own OpenMP reduction (QubitRegisterCalculator.h:1039,
            # 1205) makes its low bits depend on the thread count, so "exact" means 1e-15 here
            assert np.max(np.abs(val - want)) <= 1e-15, (key, float(np.max(np.abs(val - want))))
        else:
            assert np.max(np.abs(val - want)) <= tol, (key, float(np.max(np.abs(val - want))))
