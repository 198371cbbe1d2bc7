#!/bin/bash
# Multi-GPU evidence for one shard count N (run under `gpurun --gpus N`): the driver's exact bench command, the QFT
# workload, the sharded / multi-device parity tests for this N, Grover at 32 qubits with the reference's 201 iterations,
# and (N = 8) the 36-qubit QFT + random circuit of BASELINE config 5 and the C++ facade on 8 GPUs.
N=${1:-8}
R=${2:-r02}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
O=gpurun_out
mkdir -p $O
run() { name=$1; shift; timeout 600 "$@" 2>$O/${R}_${name}.err | grep "^{" > $O/${R}_${name}.json; echo "$name rc=$? $(cut -c1-150 $O/${R}_${name}.json)"; }
if [ "$N" != 1 ]; then
run bench${N}_random $TR bench.py --gpus $N --steps 20 --warmup 5
run bench${N}_qft30 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload qft --qubits 30
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi.py -m gpu -x -q -k "$N" 2>&1 | tail -4 | tee $O/${R}_pytest_sharded${N}.log
timeout 600 $TR tools/grover_sharded.py 16 full 2>&1 | grep "GROVER\|rror\|ssert" | tee $O/${R}_grover32_${N}gpu.log
timeout 200 $TR tools/exch_bench.py 33 multi 2>&1 | grep "n=33\|rror" | tail -1 | tee $O/${R}_exchange${N}.log
fi
if [ "$N" = 1 ]; then
  # QCSim's own GroverAlgorithm.h on the C++ drop-in class, 31 qubits, ONE GPU: the amplitudes the 8-GPU run is compared with
  ( time QCSIM_B200_FUSION=1 QCSIM_B200_DEVICES=0 timeout 900 tests/cpp/facade_test.bin grover_ranges 16 46757 $O/facade_g31_1.bin ) 2>&1 | grep -v "^$" | tr '\n' ' ' | tee $O/${R}_facade_grover31_1gpu.log; echo
  exit 0
fi
if [ "$N" = 8 ]; then
  run bench8_qft36 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload qft --qubits 33
  run bench8_random36 $TR bench.py --gpus 8 --steps 10 --warmup 3 --qubits 33
fi
if [ "$N" -ge 4 ]; then
  # QCSim's own GroverAlgorithm.h on the C++ drop-in class, 31 qubits, N GPUs of one process
  DEVS=$(seq -s, 0 $((N-1)))
  ( time QCSIM_B200_FUSION=1 QCSIM_B200_DEVICES=$DEVS timeout 600 tests/cpp/facade_test.bin grover_ranges 16 46757 $O/facade_g31_$N.bin ) 2>&1 | grep -v "^$" | tr '\n' ' ' | tee $O/${R}_facade_grover31_${N}gpu.log; echo
fi
python tools/compare_grover.py | tee $O/${R}_grover_compare${N}.log
