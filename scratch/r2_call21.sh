#!/bin/bash
# Round 2, GPU call 21: fork-join continuation kernel v3 (static deal of the initial pieces, tickets for forked pieces, few helper warps, DONE flag) — parity suite, sweeps
O=gpurun_out/r2_c21; mkdir -p $O
export PCT_B200_LIB=$PWD/scratch/variants/lib_c21.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -5 $O/tests.log | tee -a $O/summary.txt
B="python bench.py --steps 400 --warmup 200 --e2e-steps 50 --skip-cpu --skip-configs"
PCT_B200_WALK=seq timeout 200 $B > $O/bench_seq.log 2>&1
timeout 200 $B > $O/bench_fork_k296_b6_l16.log 2>&1
for k in 74 148 592 1184; do PCT_B200_WALK_KEEP=$k timeout 200 $B > $O/bench_fork_k${k}_b6_l16.log 2>&1; done
for b in 4 8; do PCT_B200_WALK_BLOCKS=$b timeout 200 $B > $O/bench_fork_k296_b${b}_l16.log 2>&1; done
for l in 4 8; do PCT_B200_WALK_LANES=$l timeout 200 $B > $O/bench_fork_k296_b6_l$l.log 2>&1; done
PCT_B200_WALK=seq timeout 200 $B --continuous > $O/bench_cont_seq.log 2>&1
timeout 200 $B --continuous > $O/bench_cont_fork.log 2>&1
python - <<'PY' | tee -a gpurun_out/r2_c21/summary.txt
import glob, json
for f in sorted(glob.glob("gpurun_out/r2_c21/bench_*.log")):
    for line in open(f):
        if line.startswith("{"):
            j = json.loads(line)
            print("%-28s value %.2fM  e2e %.2fM  ms/step %.4f  kernels %s" % (f.split("/")[-1][:-4], j["value"] / 1e6, j["e2e"]["value"] / 1e6,
                  j["ms_per_step"], j["roofline"].get("all_kernels_ms")))
PY
