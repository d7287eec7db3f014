#!/bin/bash
# Round 2, GPU call 18: continuous kernels — packed containment test in the EMS purge, coordinate hashes shared by the four corner lanes
O=gpurun_out/r2_c18; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c18.so  # frozen copy of the library under test (the tree moves on while the call waits for a GPU slot)
( timeout 900 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
timeout 200 $B --continuous > $O/bench_cont_new1.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_new2.log 2>&1
timeout 200 $B > $O/bench_head.log 2>&1
( timeout 600 python scratch/soak_gpu_vs_oracle.py --continuous --envs 2048 --steps 150 ) > $O/soak_cont.log 2>&1; tail -2 $O/soak_cont.log | tee -a $O/summary.txt
python - <<'PY' | tee -a gpurun_out/r2_c18/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c18/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
