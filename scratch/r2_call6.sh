#!/bin/bash
# Round 2, GPU call 6: continuations per warp sweep (+ orient3 horizontal shortcut build): parity, bench sweep
O=gpurun_out/r2_c6; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_discrete_parity.py tests/test_gpu_discrete_cases.py tests/test_zzz_gpu_alias.py -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
for L in 8 2 4 16 32; do PCT_B200_WALK_LANES=$L timeout 200 $B > $O/bench_L$L.log 2>&1; done
PCT_B200_WALK_LANES=4 timeout 200 $B --envs-per-gpu 8192 > $O/bench_8192_L4.log 2>&1
PCT_B200_WALK_LANES=8 timeout 200 $B --envs-per-gpu 8192 > $O/bench_8192_L8.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c6/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c6/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
