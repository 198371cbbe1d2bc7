#!/bin/bash
# round-1 single-GPU evidence run: tests, benches, ncu launch lists and full captures (-> gpurun_out/)
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/r01_pytest_gpu.log; cat $O/r01_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r01_smoke.log
python bench.py > $O/r01_bench_random_30q.json 2> $O/r01_bench_random_30q.err; cut -c1-300 $O/r01_bench_random_30q.json
python bench.py --workload qft --no-kernel-sweep > $O/r01_bench_qft_30q.json 2> $O/r01_bench_qft_30q.err; cut -c1-300 $O/r01_bench_qft_30q.json
python bench.py --workload qft --qubits 33 --steps 3 --no-kernel-sweep --no-cpu-baseline > $O/r01_bench_qft_33q.json 2> $O/r01_bench_qft_33q.err; cut -c1-300 $O/r01_bench_qft_33q.json
python bench.py --fusion 0 --steps 5 --no-kernel-sweep --no-cpu-baseline > $O/r01_bench_random_30q_unfused.json 2>/dev/null; cut -c1-200 $O/r01_bench_random_30q_unfused.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/r01_launches_random.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/r01_launches_qft.csv python bench.py --workload qft --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_tile_pass -s 4 -c 1 -o $O/r01_k_tile_pass_30q python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_qft_pass -s 2 -c 1 -o $O/r01_k_qft_pass_30q python bench.py --workload qft --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_pair_v2 -s 3 -c 1 -o $O/r01_k_pair_v2_30q python bench.py --fusion 0 --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_bit_reverse -s 1 -c 1 -o $O/r01_k_bit_reverse_30q python bench.py --workload qft --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-sweep > /dev/null 2>&1
python bench.py --impl reference > $O/r01_bench_reference.json 2> $O/r01_bench_reference.err; cut -c1-300 $O/r01_bench_reference.json
ls -la $O | tail -20
