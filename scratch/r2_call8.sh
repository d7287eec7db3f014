#!/bin/bash
# Round 2, GPU call 8: evidence run at HEAD — full GPU suite, default bench (all configs + CPU baseline), reference arm, GPU<->oracle soak,
# the reference's own ShmemVecEnv on this host, rollout throughput with DRL_GAT, ncu launch list + full captures of every discrete kernel
O=gpurun_out/r2_c8; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
( timeout 600 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_reference.json 2>&1; echo "ref rc=$?" | tee -a $O/summary.txt
( timeout 300 python bench.py --steps 20 --warmup 5 ) > $O/bench_driver_flags.json 2> $O/bench_driver_flags.err; echo "bench(driver flags) rc=$?" | tee -a $O/summary.txt
( timeout 600 python scratch/soak_gpu_vs_oracle.py --envs 4096 --steps 300 ) > $O/soak_discrete.txt 2>&1; echo "soak discrete rc=$?" | tee -a $O/summary.txt
( timeout 600 python scratch/soak_gpu_vs_oracle.py --envs 2048 --steps 150 --continuous ) > $O/soak_continuous.txt 2>&1; echo "soak continuous rc=$?" | tee -a $O/summary.txt
( timeout 600 python scratch/shmem_baseline.py --procs 1,64,128 --steps 200 --warmup 20 ) > $O/shmem_baseline_s1.txt 2>&1; echo "shmem rc=$?" | tee -a $O/summary.txt
( timeout 300 python scratch/shmem_baseline.py --procs 64 --steps 200 --warmup 20 --setting 2 ) > $O/shmem_baseline_s2.txt 2>&1
( timeout 300 python scratch/shmem_baseline.py --procs 64 --steps 100 --warmup 10 --continuous ) > $O/shmem_baseline_cont.txt 2>&1
( timeout 600 python scratch/bench_rollout.py ) > $O/rollout.txt 2>&1; echo "rollout rc=$?" | tee -a $O/summary.txt
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 80 --csv --log-file $O/launches.csv $B2 > $O/ncu_launches.log 2>&1
for k in pct_apply pct_candidates pct_walk_light "pct_walk_kernel" pct_emit; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 80 -c 1 -o $O/head_$k $B2 > $O/ncu_$k.log 2>&1
done
nproc > $O/nproc.txt; lscpu | head -20 > $O/lscpu.txt
tail -c 600 $O/soak_discrete.txt $O/soak_continuous.txt $O/shmem_baseline_s1.txt $O/rollout.txt | tee -a $O/summary.txt
ls -la $O
