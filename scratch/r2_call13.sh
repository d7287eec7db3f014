#!/bin/bash
# Round 2, GPU call 13: height-aware dealing of the continuations (tall walks: fewer lanes per warp): parity, sweep
O=gpurun_out/r2_c13; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_discrete_parity.py tests/test_gpu_discrete_cases.py tests/test_zzz_gpu_alias.py tests/test_gpu_continuous_parity.py tests/test_zz_gpu_continuous_full_size.py tests/test_zz_gpu_golden_replay.py -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
for cfg in "8 3" "16 3" "16 2" "16 1" "8 2" "8 1" "16 4" "32 2" "8 8"; do set -- $cfg; PCT_B200_WALK_LANES=$1 PCT_B200_WALK_LANES_TALL=$2 timeout 200 $B > $O/bench_L$1_T$2.log 2>&1; done
PCT_B200_WALK_LANES=16 PCT_B200_WALK_LANES_TALL=2 timeout 200 $B --continuous > $O/bench_cont_L16_T2.log 2>&1
PCT_B200_WALK_LANES=8 PCT_B200_WALK_LANES_TALL=3 timeout 200 $B --continuous > $O/bench_cont_L8_T3.log 2>&1
PCT_B200_WALK_LANES=8 PCT_B200_WALK_LANES_TALL=1 timeout 200 $B --continuous > $O/bench_cont_L8_T1.log 2>&1
( timeout 200 python scratch/e2e_breakdown.py ) 2>&1 | tee -a gpurun_out/r2_c13/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_c13/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c13/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
