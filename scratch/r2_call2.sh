#!/bin/bash
# Round 2, GPU call 2: the warp-per-env feasibility kernel (K3 v2) — parity suite, then A/B against round 1's block kernel.
O=gpurun_out/r2_c2; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -5 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 150 --skip-cpu"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_v2 $B
PCT_B200_K3=block run bench_block $B
PCT_B200_OBS_DELTA=0 run bench_v2_nodelta $B
run bench_s2 $B --setting 2 --envs-per-gpu 8192
PCT_B200_K3=block run bench_s2_block $B --setting 2 --envs-per-gpu 8192
run bench_cont $B --continuous
python - <<'PY' | tee -a gpurun_out/r2_c2/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c2/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
