#!/bin/bash
# Round 2, GPU call 28: the apply kernels' descent scratch aliased onto their EMS temp areas in shared memory (no extra shared memory) — suite, A/B against the previous commit
O=gpurun_out/r2_c28; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c28.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu --skip-configs"
for rep in 1 2 3; do
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 200 $B > $O/bench_old_$rep.log 2>&1
timeout 200 $B > $O/bench_new_$rep.log 2>&1
done
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 200 $B --continuous > $O/bench_cont_old.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_new.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c28/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c28/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
