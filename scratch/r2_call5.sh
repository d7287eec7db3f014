#!/bin/bash
# Round 2, GPU call 5: two-stage pooled walks (light prefix -> continuation): discrete parity files, bench, ncu of both walk kernels
O=gpurun_out/r2_c5; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_discrete_parity.py tests/test_gpu_discrete_cases.py tests/test_zz_gpu_golden_replay.py tests/test_zzz_gpu_alias.py tests/test_shuffle.py tests/test_gpu_hostapi.py tests/test_zzz_gpu_obs_delta.py tests/test_gpu_rollout.py -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 100 --skip-cpu --skip-configs"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_pool2 $B
PCT_B200_K3=block run bench_block $B
run bench_s3 $B --setting 3
run bench_pool2_8192 $B --envs-per-gpu 8192
python - <<'PY' | tee -a gpurun_out/r2_c5/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c5/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_walk -s 80 -c 2 -o $O/walk2 $B2 > $O/ncu_walk2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 70 --csv --log-file $O/launches.csv $B2 > $O/ncu_launches.log 2>&1
ls -la $O
