#!/bin/bash
# one 8-GPU session: parity, exchange bandwidth, benches (outputs under gpurun_out/)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "8" 2>&1 | tail -4 | tee gpurun_out/r01_sharded8_pytest.log
timeout 200 $TR scratch/exch_bench.py 32 2>&1 | grep "n=32\|rror" | tee gpurun_out/r01_exchange8.log
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-kernel-sweep 2>gpurun_out/r01_bench8_random.err | tee gpurun_out/r01_bench8_random.json | cut -c1-400
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload qft --qubits 30 --no-kernel-sweep 2>gpurun_out/r01_bench8_qft33.err | tee gpurun_out/r01_bench8_qft33.json | cut -c1-400
timeout 400 $TR bench.py --gpus 8 --steps 2 --warmup 3 --workload qft --qubits 33 --no-kernel-sweep 2>gpurun_out/r01_bench8_qft36.err | tee gpurun_out/r01_bench8_qft36.json | cut -c1-400
tail -3 gpurun_out/r01_bench8_qft36.err
