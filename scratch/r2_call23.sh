#!/bin/bash
# Round 2, GPU call 23: more, smaller warps for the continuation kernel — launch bounds 8 / 10 / 12 blocks per SM (128 / 96 / 80 registers) x 4 / 6 / 8 / 16 continuations per warp
O=gpurun_out/r2_c23; mkdir -p $O
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu --skip-configs"
for lib in final m10 m12; do
  for l in 4 6 8 16; do
    t=$(( l < 4 ? l : 4 ))
    PCT_B200_LIB=$PWD/scratch/variants/lib_$lib.so PCT_B200_WALK_LANES=$l PCT_B200_WALK_LANES_TALL=$t timeout 200 $B > $O/bench_${lib}_l${l}_t${t}.log 2>&1
  done
done
PCT_B200_LIB=$PWD/scratch/variants/lib_m12.so PCT_B200_WALK_LANES=4 PCT_B200_WALK_LANES_TALL=2 timeout 200 $B > $O/bench_m12_l4_t2.log 2>&1
PCT_B200_LIB=$PWD/scratch/variants/lib_m12.so PCT_B200_WALK_LANES=6 PCT_B200_WALK_LANES_TALL=2 timeout 200 $B > $O/bench_m12_l6_t2.log 2>&1
python - <<'PY' | tee gpurun_out/r2_c23/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c23/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
