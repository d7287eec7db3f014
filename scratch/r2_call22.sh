#!/bin/bash
# Round 2, GPU call 22: full suite at HEAD (fork-join opt-in test included), compute-sanitizer memcheck + racecheck on the new kernels (bucket order, stored polygons, fork-join queue)
O=gpurun_out/r2_c22; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c22.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -5 $O/tests.log | tee -a $O/summary.txt
( timeout 900 compute-sanitizer --tool memcheck python scratch/sani.py ) > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; grep -E "ERROR SUMMARY|workload" $O/memcheck.log | tee -a $O/summary.txt
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python scratch/sani.py ) > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; grep -E "RACECHECK SUMMARY|workload" $O/racecheck.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
timeout 200 $B > $O/bench_head.log 2>&1
timeout 200 $B --continuous > $O/bench_cont.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c22/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c22/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
