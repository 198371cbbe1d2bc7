"""BASELINE config 4 at scale: Grover with a gates oracle (GroverAlgorithm.h:128-242) on a sharded 32-qubit register
(16 search qubits -> 31 qubits used, the 32nd idle), the reference's own iteration count
round(pi/4 sqrt(2^16)) = 201 (GroverAlgorithm.h:187) unless a smaller number is given.  Checks the marked state's
probability against sin^2((2k+1) asin 2^-8) (>= 0.99 at 201 iterations) and writes 16 sampled amplitude ranges to
gpurun_out/ so that runs on 2, 4 and 8 GPUs can be compared (tools/compare_grover.py).  torchrun, one rank per GPU.

    torchrun ... tools/grover_sharded.py [n_search=16] [iterations|full] [top]
"""
import math, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcsim_b200 import circuits
from qcsim_b200.sharded import create_register
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 16
K = circuits.grover_iterations(NS) if len(sys.argv) <= 2 or sys.argv[2] == "full" else int(sys.argv[2])
top = len(sys.argv) > 3 and sys.argv[3] == "top"
n = 2 * NS          # 2 NS - 1 used + 1 idle (config 4: "the 32nd is idle")
marked = 0xB6A5 & ((1 << NS) - 1)
# layout: reference (search qubits 0..NS-1) or SURVEY's stress layout (search qubits on top)
qmap = None
if top:
    qmap = {q: n - 1 - q for q in range(NS)}            # search qubits -> top
    qmap[NS] = NS - 1                                    # oracle target
    for a in range(NS - 2):
        qmap[NS + 1 + a] = a + 1                         # ancillas low, qubit 0 idle
circ = circuits.grover_gates_circuit(NS, marked, iterations=K, qubit_map=qmap)
reg = create_register(n, local, rank, world, dist)
reg.set_fusion(True)
torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
CH = 4096
for i in range(0, len(circ), CH):                       # bounded host staging; the engine's queue fuses across the chunks
    reg.ApplyGates(circ[i:i + CH])
reg.sync(); torch.cuda.synchronize(); dist.barrier(); t1 = time.time()
st = reg.stats()
phys = (lambda s: s) if qmap is None else (lambda s: sum(((s >> q) & 1) << qmap[q] for q in range(2 * NS - 1)))
p = sum(reg.getBasisStateProbability(phys(marked | (t << NS))) for t in (0, 1))
want = math.sin((2 * K + 1) * math.asin(2.0 ** (-NS / 2))) ** 2
nrm = reg.norm2()
# 16 global ranges of 1024 amplitudes, each saved by the rank that owns it
os.makedirs("gpurun_out", exist_ok=True)
dim = 1 << n
for j in range(16):
    first = ((dim // 16) * j + 1024 * j) & ~1023
    if reg.slice_first <= first < reg.slice_first + reg.slice_count:
        np.save(f"gpurun_out/grover{n}_{'top' if top else 'ref'}_k{K}_w{world}_r{j}.npy", reg.download(first, 1024))
dist.barrier()
if rank == 0:
    print(f"GROVER n={n} world={world} layout={'top' if top else 'reference'} iterations={K} gates={len(circ)} time={t1-t0:.3f}s "
          f"({(t1-t0)/K*1e3:.1f} ms/iteration) P(marked)={p:.15e} analytic={want:.15e} diff={abs(p-want):.2e} norm2-1={nrm-1:.2e} "
          f"exchanges={st['exchange_calls']} exch_GB={st['exchange_bytes']/1e9:.1f} exch_ms={st['exchange_ms']:.0f} passes={st['state_passes']}", flush=True)
    assert abs(p - want) < 1e-10 and abs(nrm - 1) < 1e-10, (p, want, nrm)
    if K == circuits.grover_iterations(NS):
        assert p >= 0.99, p
reg.close(); dist.destroy_process_group()
