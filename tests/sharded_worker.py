"""Worker of tests/test_gpu_sharded.py: one process per GPU under torch.distributed.run.

Every rank builds the same circuits, runs them on its shard of a ShardedQubitRegister (through
the C ABI), and checks its slice / the measurement outcomes against the oracle, which every rank
computes in full at these sizes.  Prints one line `SHARDED_OK <world> <checks>` from rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402  (test infrastructure: the checker)
from conftest import draws, random_state  # noqa: E402
from qcsim_b200 import circuits, gates  # noqa: E402
from qcsim_b200.sharded import create_register  # noqa: E402

TOL = 1e-12


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    checks = 0
    worst = 0.0

    def slice_of(v, reg):
        return v[reg.slice_first: reg.slice_first + reg.slice_count]

    def check_state(reg, ref, what):
        nonlocal checks, worst
        got = reg.getRegisterStorage()
        want = slice_of(ref.state(), reg)
        err = float(np.max(np.abs(got - want)))
        worst = max(worst, err)
        assert err <= TOL, (what, rank, err)
        assert abs(reg.norm2() - ref.norm2()) <= TOL, what
        checks += 1

    for mode in ("immediate", "batch", "fusion"):
        for n, layers in ((8, 4), (13, 5), (17, 4)):
            circ = circuits.random_circuit(n, layers, seed=100 + n)
            # make sure global qubits see every kind of action
            circ += [(gates.HadamardGate(), n - 1, 0, 0), (gates.CNOTGate(), n - 1, 0, 0), (gates.CNOTGate(), 0, n - 1, 0),
                     (gates.ToffoliGate(), n - 1, n - 2, 1), (gates.SwapGate(), 0, n - 1, 0), (gates.RzGate(0.3), n - 1, 0, 0),
                     (gates.ControlledPhaseShiftGate(0.7), n - 1, n - 2, 0), (gates.FredkinGate(), 2, n - 1, n - 2),
                     (gates.iSwapGate(), n - 2, 1, 0), (gates.RxGate(0.9), n - 2, 0, 0)]
            psi0 = random_state(n, 31)
            reg = create_register(n, local, rank, world, dist)
            with oracle.best_oracle(n) as ref:
                ref.set_state(psi0)
                reg.upload_slice(slice_of(psi0, reg))
                if mode == "fusion":
                    reg.set_fusion(True)
                if mode == "immediate":
                    for g in circ:
                        reg.ApplyGate(*g)
                else:
                    reg.ApplyGates(circ)
                ref.apply_circuit(circ)
                check_state(reg, ref, f"random {mode} n={n}")
                for q in (0, n // 2, n - 1):
                    assert abs(reg.GetQubitProbability(q) - ref.qubit_probability(q)) <= TOL
                # measurement: identical outcomes for identical injected draws, then identical collapse
                for strict in (False, True):
                    reg.set_strict_measure(strict)
                    for p in draws(4, 7 + n):
                        assert reg.MeasureNoCollapse(prob=p) == ref.measure_all_nocollapse(p), (mode, n, p)
                p = draws(1, 3)[0]
                assert reg.Measure(n - 2, n - 1, p) == ref.measure(n - 2, n - 1, p)
                check_state(reg, ref, f"partial collapse {mode} n={n}")
                p = draws(1, 4)[0]
                assert reg.MeasureAll(p) == ref.measure_all(p)
                check_state(reg, ref, f"full collapse {mode} n={n}")
            reg.close()

    # QFT / IQFT (BASELINE config 1 shape), engine call and gate by gate
    for n in (10, 16, 20):
        psi0 = random_state(n, 7)
        reg = create_register(n, local, rank, world, dist)
        with oracle.best_oracle(n) as ref:
            ref.set_state(psi0)
            reg.upload_slice(slice_of(psi0, reg))
            ref.qft()
            reg.QFT()
            st0 = reg.stats()
            check_state(reg, ref, f"qft n={n}")
            want = np.sqrt(1 << n) * np.fft.ifft(psi0)
            assert np.max(np.abs(reg.getRegisterStorage() - slice_of(want, reg))) <= 1e-11
            ref.qft(inverse=True)
            reg.QFT(inverse=True)
            check_state(reg, ref, f"iqft n={n}")
            assert np.max(np.abs(reg.getRegisterStorage() - slice_of(psi0, reg))) <= 1e-11
            # sub-register, no swap (Draper adder shape)
            ref.qft(2, n - 2, False, False)
            reg.QFT(2, n - 2, False, False)
            check_state(reg, ref, f"sub qft n={n}")
            assert st0["exchange_calls"] >= 1 and st0["exchange_bytes"] > 0
        reg.close()

    # chained transforms without looking at the state in between: the register stays in whatever layout the
    # previous transform left (virtual qubit reversal, exchanged qubits) -- nothing is canonicalised
    for n in (12, 17):
        psi0 = random_state(n, 9)
        reg = create_register(n, local, rank, world, dist)
        with oracle.best_oracle(n) as ref:
            ref.set_state(psi0)
            reg.upload_slice(slice_of(psi0, reg))
            reg.reset_stats()
            for args in ((0, n - 1, True, False), (0, n - 1, True, False), (1, n - 2, True, True), (0, n - 1, False, False),
                         (0, n - 1, True, True), (2, n - 1, False, True)):
                ref.qft(*args)
                reg.QFT(*args)
            passes = reg.stats()["state_passes"]
            check_state(reg, ref, f"chained qft n={n}")
            if rank == 0:
                print(f"chained QFT n={n}: {passes} state passes for 6 transforms, {reg.stats()['exchange_calls']} exchanges", flush=True)
        reg.close()

    # Grover with gates (config 4 shape, small): the n-controlled NOT ladders over all shards
    n_search = 5
    n = 2 * n_search - 1
    marked = 0b10110
    circ = circuits.grover_gates_circuit(n_search, marked)
    reg = create_register(n, local, rank, world, dist)
    with oracle.best_oracle(n) as ref:
        reg.set_fusion(True)
        reg.ApplyGates(circ)
        ref.apply_circuit(circ)
        check_state(reg, ref, "grover")
    reg.close()

    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_OK {world} {checks} maxerr={float(t.item()):.3e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
