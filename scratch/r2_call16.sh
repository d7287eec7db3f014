#!/bin/bash
# Round 2, GPU call 16: stored shrunk polygons + stored two-support split + cached contact centres (A/B against the previous commit's library), LPT on/off
O=gpurun_out/r2_c16; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
for rep in 1 2; do
PCT_B200_LIB=scratch/variants/lib_head.so timeout 200 $B > $O/bench_head_old$rep.log 2>&1
timeout 200 $B > $O/bench_head_new$rep.log 2>&1
PCT_B200_LPT=0 timeout 200 $B > $O/bench_head_new_nolpt$rep.log 2>&1
done
PCT_B200_LIB=scratch/variants/lib_head.so timeout 200 $B --continuous > $O/bench_cont_old.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_new.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c16/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c16/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
