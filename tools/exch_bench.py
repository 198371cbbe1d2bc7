import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qcsim_b200 import gates
from qcsim_b200.sharded import create_register
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 29
reg = create_register(n, local, rank, world, dist)
H = gates.HadamardGate()
g = world.bit_length() - 1
# each H on a currently-global qubit forces an exchange (victim = a local qubit, which becomes global)
for rep in range(3):
    reg.reset_stats()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    multi = len(sys.argv) > 2 and sys.argv[2] == "multi"
    for i in range(4):
        if multi:   # all global qubits in one batch -> one k = log2(world) exchange (all-to-all), and back
            reg.ApplyGates([(H, q, 0, 0) for q in range(n - g, n)])
            reg.ApplyGates([(H, n - g - 1 - q, 0, 0) for q in range(0, g)])
            continue
        for q in range(n - g, n):
            reg.ApplyGate(H, q)
        for q in range(0, g):
            reg.ApplyGate(H, n - g - 1 - q)   # touch parked qubits so they come back
    reg.sync(); torch.cuda.synchronize(); dist.barrier(); t1 = time.time()
    st = reg.stats()
    if rank == 0:
        gb = st["exchange_bytes"] / 1e9
        print(f"n={n} world={world} exchanges={st['exchange_calls']} bytes/rank={gb:.2f} GB exch_ms={st['exchange_ms']:.1f} -> {gb/ (st['exchange_ms']*1e-3):.1f} GB/s per direction; wall {t1-t0:.3f}s", flush=True)
reg.close()
dist.destroy_process_group()
