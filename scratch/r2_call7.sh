#!/bin/bash
# Round 2, GPU call 7: continuous pooled pipeline (parity + bench), discrete: internal stream groups / LPT off / K1 register variants
O=gpurun_out/r2_c7; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_continuous_parity.py tests/test_zz_gpu_continuous_full_size.py tests/test_zzz_gpu_continuous_pre.py tests/test_gpu_heuristics_continuous.py tests/test_f32_rows.py tests/test_shuffle.py tests/test_zzz_gpu_alias.py tests/test_zz_gpu_golden_replay.py -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
run() { name=$1; shift; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "$name rc=$?" >> $O/summary.txt; }
run bench_cont_pool $B --continuous
PCT_B200_K3=block run bench_cont_block $B --continuous
PCT_B200_WALK_LANES=4 run bench_cont_pool_L4 $B --continuous
PCT_B200_WALK_LANES=16 run bench_cont_pool_L16 $B --continuous
run bench_cont_s2 $B --continuous --setting 2
PCT_B200_LPT=0 run bench_nolpt $B
PCT_B200_LPT=0 PCT_B200_GROUPS=2 run bench_g2 $B
PCT_B200_LPT=0 PCT_B200_GROUPS=4 run bench_g4 $B
PCT_B200_LPT=0 PCT_B200_GROUPS=8 run bench_g8 $B
PCT_B200_LPT=0 PCT_B200_LIB=$PWD/scratch/variants/lib_k1m8.so run bench_k1m8 $B
PCT_B200_LPT=0 PCT_B200_LIB=$PWD/scratch/variants/lib_k1m5.so run bench_k1m5 $B
python - <<'PY' | tee -a gpurun_out/r2_c7/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c7/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu --skip-configs --preroll 40 --continuous"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pctc_ -s 400 -c 6 -o $O/cont_pool $B2 > $O/ncu_cont.log 2>&1
ls -la $O
