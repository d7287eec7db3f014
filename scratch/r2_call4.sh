#!/bin/bash
# Round 2, GPU call 4: pooled-walk pipeline (K2 classify -> pct_walk_kernel -> pct_emit_kernel): parity suite, A/B bench, ncu
O=gpurun_out/r2_c4; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -5 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 100 --skip-cpu --skip-configs"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_pool python bench.py --steps 400 --warmup 200 --e2e-steps 100
PCT_B200_K3=block run bench_block $B
PCT_B200_OVERLAP=0 run bench_pool_nooverlap $B
PCT_B200_LPT=0 run bench_pool_nolpt $B
run bench_s2 $B --setting 2 --envs-per-gpu 8192
run bench_s3 $B --setting 3
run bench_pool_8192 $B --envs-per-gpu 8192
python - <<'PY' | tee -a gpurun_out/r2_c4/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c4/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_walk -s 40 -c 1 -o $O/walk $B2 > $O/ncu_walk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_candidates -s 40 -c 1 -o $O/k2 $B2 > $O/ncu_k2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_emit -s 40 -c 1 -o $O/emit $B2 > $O/ncu_emit.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file $O/launches.csv $B2 > $O/ncu_launches.log 2>&1
ls -la $O
