#!/bin/bash
# Round 2, final evidence run at HEAD: full GPU suite, smoke(), ncu launch list + full captures of every kernel of the step (-> profiles/traffic.json of THIS build,
# regenerated on the box before the benches run), reference arm, default bench (all configs, CPU baseline), driver-flag bench, GPU-vs-oracle soaks
O=gpurun_out/r2_final; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_final.so  # frozen copy of the in-tree library at the commit under test
NOTE="$(cat scratch/variants/final_note.txt 2>/dev/null)"
( timeout 1200 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log | tee -a $O/summary.txt
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 70 --csv --log-file $O/launches.csv $B2 > $O/ncu_launches.log 2>&1
for k in pct_apply pct_candidates pct_walk_light "pct_walk_kernel" pct_emit; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 80 -c 1 -o $O/head_$k $B2 > $O/ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pctc_ -s 400 -c 6 -o $O/cont_head $B2 --continuous > $O/ncu_cont.log 2>&1
( python scratch/collect_final.py $O "$NOTE" --traffic-only ) 2>&1 | tail -1 | tee -a $O/summary.txt
( timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/bench_reference.json 2>&1; echo "ref rc=$?" | tee -a $O/summary.txt
( timeout 600 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
( timeout 300 python bench.py --steps 20 --warmup 5 ) > $O/bench_driver_flags.json 2> $O/bench_driver_flags.err; echo "bench(driver flags) rc=$?" | tee -a $O/summary.txt
( timeout 900 python scratch/soak_gpu_vs_oracle.py --envs 2048 --steps 200 ) > $O/soak_discrete.log 2>&1; tail -1 $O/soak_discrete.log | tee -a $O/summary.txt
( timeout 600 python scratch/soak_gpu_vs_oracle.py --continuous --envs 1024 --steps 150 ) > $O/soak_continuous.log 2>&1; tail -1 $O/soak_continuous.log | tee -a $O/summary.txt
ls -la $O | tee -a $O/summary.txt
