#!/bin/bash
# Round 2, GPU call 3: register-cap / block-shape variants of K3 v2 + ncu of the default v2 build
O=gpurun_out/r2_c3; mkdir -p $O
B="python bench.py --steps 400 --warmup 200 --e2e-steps 20 --skip-cpu"
for v in base m12 m16 w4; do
  if [ $v = base ]; then unset PCT_B200_LIB; else export PCT_B200_LIB=$PWD/scratch/variants/lib_$v.so; fi
  ( timeout 200 $B ) > $O/bench_$v.log 2>&1
done
unset PCT_B200_LIB
python - <<'PY' | tee $O/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c3/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.3f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
B2="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_feas_emit2 -s 40 -c 1 -o $O/k3v2 $B2 > $O/ncu_k3v2.log 2>&1
export PCT_B200_LIB=$PWD/scratch/variants/lib_m16.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_feas_emit2 -s 40 -c 1 -o $O/k3v2_m16 $B2 > $O/ncu_k3v2_m16.log 2>&1
ls -la $O
