#!/bin/bash
# Round 2, GPU call 26: apply kernel with the descent's working arrays per warp in shared memory (instead of lane 0's local-memory stack) at 5 / 6 / 7 blocks per SM, against the previous commit
O=gpurun_out/r2_c26; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c26.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu --skip-configs"
for rep in 1 2; do
for v in final c26 scr5 scr7; do
PCT_B200_LIB=$PWD/scratch/variants/lib_$v.so timeout 200 $B > $O/bench_${v}_$rep.log 2>&1
done
done
PCT_B200_LIB=$PWD/scratch/variants/lib_final.so timeout 200 $B --continuous > $O/bench_cont_final.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_c26.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c26/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c26/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
