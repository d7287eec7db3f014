#!/bin/bash
# Round 2, GPU call 25: suite at HEAD (fork-join buffers allocated only in fork mode); apply kernel: smaller staging areas / 5 blocks per SM -> smaller shared-memory carve-out, more L1 for the descent's stack
O=gpurun_out/r2_c25; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c25.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -4 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu --skip-configs"
for rep in 1 2; do
for v in c25 es32 es16 k1m5; do
PCT_B200_LIB=$PWD/scratch/variants/lib_$v.so timeout 200 $B > $O/bench_${v}_$rep.log 2>&1
done
done
python - <<'PY' | tee -a gpurun_out/r2_c25/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c25/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
