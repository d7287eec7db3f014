#!/bin/bash
# Round 2, GPU call 27: pct_step_host zero-copy path — leaf indices read from the mapped host buffer, info records stored into it by the emit kernel: host-path tests, e2e A/B
O=gpurun_out/r2_c27; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c27.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 300 --warmup 200 --e2e-steps 400 --skip-cpu --skip-configs"
for rep in 1 2; do
PCT_B200_HOST_MAP_IDX=0 PCT_B200_HOST_MAP_INFO=0 timeout 200 $B > $O/bench_staged_both_$rep.log 2>&1
PCT_B200_HOST_MAP_IDX=1 PCT_B200_HOST_MAP_INFO=0 timeout 200 $B > $O/bench_map_idx_$rep.log 2>&1
PCT_B200_HOST_MAP_IDX=0 PCT_B200_HOST_MAP_INFO=1 timeout 200 $B > $O/bench_map_info_$rep.log 2>&1
timeout 200 $B > $O/bench_map_both_$rep.log 2>&1
done
python - <<'PY' | tee -a gpurun_out/r2_c27/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c27/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.3fM  vec %.2fM  ms/step %.4f" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6, (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"]))
PY
