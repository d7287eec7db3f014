#!/bin/bash
# Round 2, GPU call 14/15: host path with mapped reward / done (write-only) buffers: host-path tests + e2e
O=gpurun_out/r2_c15; mkdir -p $O
( timeout 900 python -m pytest tests/test_zzz_gpu_host_zerocopy.py tests/test_zzz_gpu_obs_delta.py tests/test_gpu_hostapi.py tests/test_gpu_discrete_parity.py tests/test_gpu_continuous_parity.py -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
( timeout 200 python scratch/e2e_breakdown.py ) 2>&1 | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 200 --skip-cpu --skip-configs"
timeout 200 $B > $O/bench_head.log 2>&1
timeout 200 $B --continuous > $O/bench_cont.log 2>&1
PCT_B200_HOST_ZEROCOPY=0 timeout 200 $B > $O/bench_staged.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c15/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c15/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  vec %.2fM ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  (j["vec_env"]["value"] or 0) / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
