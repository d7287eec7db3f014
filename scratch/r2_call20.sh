#!/bin/bash
# Round 2, GPU call 20: fork-join continuation kernel with per-lane tickets (no CAS), both domains — parity suite, A/B against the sequential kernel, blocks / lanes sweep
O=gpurun_out/r2_c20; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c20.so
( timeout 900 python -m pytest tests -m gpu -x -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -3 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
PCT_B200_WALK=seq timeout 200 $B > $O/bench_seq.log 2>&1
timeout 200 $B > $O/bench_fork_b6_l16.log 2>&1
for b in 3 4 8; do PCT_B200_WALK_BLOCKS=$b timeout 200 $B > $O/bench_fork_b${b}_l16.log 2>&1; done
for l in 4 8 32; do PCT_B200_WALK_LANES=$l timeout 200 $B > $O/bench_fork_b6_l$l.log 2>&1; done
PCT_B200_WALK=seq timeout 200 $B --continuous > $O/bench_cont_seq.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_fork.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c20/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c20/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
