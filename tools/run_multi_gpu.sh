#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-kernel-sweep 2>$O/r01_bench${N}_random.err | grep "^{" > $O/r01_bench${N}_random.json; cut -c1-200 $O/r01_bench${N}_random.json
timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload qft --qubits 30 --no-kernel-sweep 2>$O/r01_bench${N}_qft.err | grep "^{" > $O/r01_bench${N}_qft30.json; cut -c1-200 $O/r01_bench${N}_qft30.json
if [ "$N" = 8 ]; then
timeout 400 $TR bench.py --gpus 8 --steps 2 --warmup 3 --workload qft --qubits 33 --no-kernel-sweep 2>$O/r01_bench8_qft36.err | grep "^{" > $O/r01_bench8_qft36.json; cut -c1-200 $O/r01_bench8_qft36.json
timeout 300 $TR tools/grover_sharded.py 16 2 2>&1 | grep "GROVER\|rror" | tee $O/r01_grover32_8gpu.log
timeout 300 $TR tools/grover_sharded.py 16 2 top 2>&1 | grep "GROVER\|rror" | tee -a $O/r01_grover32_8gpu.log
timeout 200 $TR tools/exch_bench.py 32 multi 2>&1 | grep "n=32\|rror" | tee $O/r01_exchange8_alltoall.log
fi
