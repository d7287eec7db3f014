#!/bin/bash
# First GPU call of round 2 (one box, ~6-8 min): runs the hardware-unverified test files of round 1 and measures the three prepared
# switches next to the default, everything into gpurun_out/r2_first/.  From the repo root:
#   gpurun --timeout 900 -- 'bash scratch/round2_first.sh'
O=gpurun_out/r2_first; mkdir -p $O
run() { name=$1; shift; echo "== $name: $*" | tee -a $O/summary.txt; ( timeout 240 "$@" ) > $O/$name.log 2>&1; echo "   rc=$?" | tee -a $O/summary.txt; tail -3 $O/$name.log | cut -c1-400 >> $O/summary.txt; }
run tests_zz   python -m pytest tests/test_zz_gpu_golden_replay.py tests/test_zz_gpu_continuous_full_size.py -q --tb=short -m gpu
run tests_zzz  python -m pytest tests/test_zzz_gpu_continuous_pre.py tests/test_zzz_gpu_host_zerocopy.py tests/test_zzz_gpu_obs_delta.py tests/test_zzz_gpu_alias.py -q --tb=short -m gpu
B="python bench.py --steps 400 --warmup 200 --e2e-steps 150 --skip-cpu"
run bench_default            $B
PCT_B200_HOST_ZEROCOPY=1 run bench_zerocopy          $B
PCT_B200_OBS_DELTA=1 run bench_delta                 $B
PCT_B200_HOST_ZEROCOPY=1 PCT_B200_OBS_DELTA=1 run bench_zerocopy_delta $B
PCT_B200_ALIAS=1 run bench_alias                     $B
run bench_cont_default       $B --continuous
PCT_B200_ALIAS=1 run bench_cont_alias                $B --continuous
PCT_B200_CONT_PRE=1 run bench_cont_pre               $B --continuous
python - <<'PY' | tee -a gpurun_out/r2_first/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_first/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.3f  K3 %s ms" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], (j["roofline"].get("kernel_ms") and "%.3f" % j["roofline"]["kernel_ms"])))
PY
