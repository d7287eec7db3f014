#!/bin/bash
# Round 2, GPU call 24: shared-memory slab for the hull scratch of the continuation kernels — parity suite, A/B, local-memory instruction counts
O=gpurun_out/r2_c24; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c24.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu --skip-configs"
for rep in 1 2; do
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 200 $B > $O/bench_old$rep.log 2>&1
timeout 200 $B > $O/bench_new$rep.log 2>&1
done
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 200 $B --continuous > $O/bench_cont_old.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_new.log 2>&1
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40"
M=gpu__time_duration.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,sass__inst_executed_shared_loads,sass__inst_executed_shared_stores,smsp__inst_executed.sum
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 300 ncu --metrics $M --clock-control none -k regex:pct_walk_kernel -s 20 -c 1 --csv --log-file $O/walk_old.csv $B2 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:pct_walk_kernel -s 20 -c 1 --csv --log-file $O/walk_new.csv $B2 > /dev/null 2>&1
grep -h "pct_walk_kernel" $O/walk_old.csv | awk -F'","' '{print "old", $(NF-2), $NF}' | tee -a $O/summary.txt
grep -h "pct_walk_kernel" $O/walk_new.csv | awk -F'","' '{print "new", $(NF-2), $NF}' | tee -a $O/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_c24/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c24/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
