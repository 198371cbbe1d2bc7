"""Single-process multi-GPU register (qcsim_sv_create_multi): one caller, one handle, the state sharded over several
GPUs of this process (run with `-m gpu` on a box with >= 2 GPUs; skipped otherwise).  Same checks as the one-process-
per-GPU tests: the compiled reference's amplitudes to 1e-12, its measurement outcomes exactly -- plus QCSim's own
GroverAlgorithm.h running on the C++ drop-in class with QCSIM_B200_DEVICES set."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import qcsim_b200
from conftest import draws, random_state
from qcsim_b200 import circuits, gates

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


def n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_device_register_matches_the_reference(world):
    if n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    devices = list(range(world))
    for n, layers in ((13, 4), (18, 3)):
        circ = circuits.random_circuit(n, layers, seed=200 + n) + [
            (gates.HadamardGate(), n - 1, 0, 0), (gates.CNOTGate(), 0, n - 1, 0), (gates.ToffoliGate(), n - 1, n - 2, 1),
            (gates.SwapGate(), 0, n - 1, 0), (gates.RxGate(0.9), n - 2, 0, 0), (gates.ControlledPhaseShiftGate(0.7), n - 1, n - 3, 0)]
        psi0 = random_state(n, 31)
        for fusion in (False, True):
            with oracle.best_oracle(n) as ref, qcsim_b200.QubitRegister(n, seed=1, devices=devices) as reg:
                ref.set_state(psi0)
                reg.setRegisterStorageFastNoNormalize(psi0)
                reg.set_fusion(fusion)
                reg.ApplyGates(circ) if fusion else [reg.ApplyGate(*g) for g in circ]
                ref.apply_circuit(circ)
                assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= TOL
                assert abs(reg.norm2() - ref.norm2()) <= TOL
                for args in ((0, n - 1, True, False), (0, n - 1, True, False), (1, n - 2, False, True)):
                    reg.QFT(*args)
                    ref.qft(*args)
                assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= TOL
                for q in (0, n - 1):
                    assert abs(reg.GetQubitProbability(q) - ref.qubit_probability(q)) <= TOL
                for p in draws(4, n):
                    assert reg.MeasureNoCollapse(prob=p) == ref.measure_all_nocollapse(p)
                reg.rng.seed(11)
                assert reg.RepeatedMeasure(300) == ref.repeated_measure(11, 300)
                p = draws(1, 3)[0]
                assert reg.Measure(n - 2, n - 1, p) == ref.measure(n - 2, n - 1, p)
                assert np.max(np.abs(reg.getRegisterStorage() - ref.state())) <= TOL
                assert reg.stats()["exchange_calls"] >= 1
                a = reg.getBasisStateAmplitude(5)
                assert abs(a - ref.state()[5]) <= TOL


def test_reference_grover_header_on_a_multi_device_facade(tmp_path):
    """QCSim's GroverAlgorithm.h (unmodified) on QC::QubitRegister with QCSIM_B200_DEVICES=0,1: same amplitudes as on one GPU"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    binary = os.path.join(HERE, "cpp", "facade_test.bin")
    if not os.path.exists(binary):
        pytest.skip("facade binary not built")
    outs = []
    for devs in ("0", "0,1"):
        out = tmp_path / f"grover_{devs.replace(',', '_')}.bin"
        env = dict(os.environ, QCSIM_B200_DEVICES=devs)
        res = subprocess.run([binary, "grover", "7", "77", str(out)], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        outs.append(np.fromfile(out, dtype=np.complex128))
    assert outs[0].shape == outs[1].shape == (1 << 13,)
    assert np.max(np.abs(outs[0] - outs[1])) <= TOL
    with oracle.best_oracle(13) as ref:
        want = ref.grover_gates(7, 77)
    assert np.max(np.abs(outs[1] - want)) <= TOL
