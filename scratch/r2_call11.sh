#!/bin/bash
# Round 2, GPU call 11: observation split (K2 writes internal + item rows) + continuous staged insertion: full GPU suite, benches
O=gpurun_out/r2_c11; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -q --tb=short -x ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -6 $O/tests.log | cut -c1-300 | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 150 --skip-cpu --skip-configs"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_head $B
PCT_B200_K3=block run bench_block $B
run bench_cont $B --continuous
run bench_s2 $B --setting 2 --envs-per-gpu 8192
run bench_s3 $B --setting 3
python - <<'PY' | tee -a gpurun_out/r2_c11/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c11/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40 --continuous"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pctc_candidates -s 80 -c 1 -o $O/k2c $B2 > $O/ncu_k2c.log 2>&1
