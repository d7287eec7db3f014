#!/bin/bash
# Round 2, GPU call 1: measure the prepared switches (round2_first.sh) + fresh ncu captures of HEAD's discrete kernels.
bash scratch/round2_first.sh
O=gpurun_out/r2_first
B="python bench.py --steps 3 --warmup 60 --e2e-steps 3 --skip-cpu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_feas_emit -s 40 -c 2 -o $O/k3_head $B > $O/ncu_k3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pct_apply -s 40 -c 2 -o $O/k1_head $B > $O/ncu_k1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pctc_ -s 120 -c 3 -o $O/cont_head $B --continuous > $O/ncu_cont.log 2>&1
ls -la $O
