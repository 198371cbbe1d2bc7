#!/bin/bash
# Single-GPU evidence of a round: full GPU test suite, smoke, default bench (both arms), QFT benches, ncu launch lists and
# one `--set full` capture per dominant kernel, compute-sanitizer on the TMA / DMMA kernels.  Run under gpurun (1 GPU).
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -25 > $O/${R}_pytest_gpu.log; tail -3 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>$O/${R}_bench_reference.err | grep "^{" > $O/${R}_bench_reference.json; cut -c1-200 $O/${R}_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 2>$O/${R}_bench_random_30q.err | grep "^{" > $O/${R}_bench_random_30q.json; cut -c1-200 $O/${R}_bench_random_30q.json
timeout 300 python bench.py --steps 5 --warmup 3 --workload qft --no-kernel-sweep 2>$O/${R}_bench_qft_30q.err | grep "^{" > $O/${R}_bench_qft_30q.json; cut -c1-160 $O/${R}_bench_qft_30q.json
timeout 300 python bench.py --steps 3 --warmup 3 --workload qft --qubits 33 --no-kernel-sweep --no-cpu-baseline 2>$O/${R}_bench_qft_33q.err | grep "^{" > $O/${R}_bench_qft_33q.json; cut -c1-160 $O/${R}_bench_qft_33q.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/${R}_launches_random.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
timeout 400 ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/${R}_launches_qft.csv python bench.py --steps 2 --warmup 1 --workload qft --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tile_pipe -s 6 -c 1 -o $O/${R}_k_tile_pipe_30q python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_qft_pipe$ -s 4 -c 1 -o $O/${R}_k_qft_pipe_30q python bench.py --steps 2 --warmup 3 --workload qft --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bit_reverse -s 1 -c 1 -o $O/${R}_k_bit_reverse_30q python bench.py --steps 2 --warmup 3 --workload qft --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ls -la $O/*.ncu-rep | tail -4
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_random_circuit_vs_oracle and 13 or fused_qft_vs_oracle and 12" 2>&1 | tail -6 ) > $O/${R}_sanitizer.log; tail -4 $O/${R}_sanitizer.log
# per-launch times of k_tile_pipe by round count (serialised launches; diagnostics)
QCSIM_DEBUG_PLAN=3 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-sweep 2>&1 | grep "pipe launch" > $O/${R}_launch_times.log; wc -l $O/${R}_launch_times.log
