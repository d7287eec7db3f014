#!/bin/bash
# Round 2, GPU call 9: rollout with the real policy (graph warm-up fix), emit-PDL A/B, host-path variants, continuous at HEAD (+ ncu)
O=gpurun_out/r2_c9; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_rollout_policy.py tests/test_gpu_rollout.py tests/test_gpu_continuous_parity.py tests/test_zz_gpu_continuous_full_size.py tests/test_gpu_discrete_parity.py tests/test_zzz_gpu_host_zerocopy.py tests/test_zzz_gpu_obs_delta.py -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -6 $O/tests.log | tee -a $O/summary.txt
( timeout 600 python scratch/bench_rollout.py ) > $O/rollout.txt 2>&1; echo "rollout rc=$?" | tee -a $O/summary.txt; tail -c 900 $O/rollout.txt | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 150 --skip-cpu --skip-configs"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_head $B
PCT_B200_EMIT_PDL=0 run bench_noemitpdl $B
PCT_B200_HOST_ZEROCOPY=0 run bench_staged $B
PCT_B200_K3=block run bench_block $B
PCT_B200_OBS_DELTA=0 run bench_nodelta $B
run bench_cont $B --continuous
PCT_B200_EMIT_PDL=0 run bench_cont_noemitpdl $B --continuous
run bench_cont_s2 $B --continuous --setting 2
python - <<'PY' | tee -a gpurun_out/r2_c9/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c9/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40 --continuous"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pctc_ -s 400 -c 5 -o $O/cont_head $B2 > $O/ncu_cont.log 2>&1
ls -la $O
