#!/bin/bash
# Refresh of the multi-GPU bench lines for one shard count N after a kernel change (run under `gpurun --gpus N`):
# the driver's exact bench command, the QFT workload and (optionally) the sharded parity tests.  tools/run_scale.sh is
# the full session (Grover, exchange bandwidth, facade).
N=${1:-2}
R=${2:-r02}
TESTS=${3:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 600 "$@" 2>$O/${R}_${name}.err | grep "^{" > $O/${R}_${name}.json; echo "$name rc=$? $(cut -c1-150 $O/${R}_${name}.json)"; }
run bench${N}_random $TR bench.py --gpus $N --steps 20 --warmup 5
if [ "$N" = 8 ]; then
  run bench8_qft36 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload qft --qubits 33
else
  run bench${N}_qft30 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload qft --qubits 30
fi
if [ "$TESTS" = 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi.py -m gpu -x -q -k "$N" 2>&1 | tail -4 | tee $O/${R}_pytest_sharded${N}.log
fi
