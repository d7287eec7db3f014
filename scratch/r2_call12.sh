#!/bin/bash
# Round 2, GPU call 12: compute-sanitizer memcheck + racecheck on the round-2 kernels (small configs), then initcheck
O=gpurun_out/r2_c12; mkdir -p $O
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scratch/sani.py ) > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee $O/summary.txt; tail -4 $O/memcheck.log | tee -a $O/summary.txt
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python scratch/sani.py ) > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; tail -4 $O/racecheck.log | tee -a $O/summary.txt
