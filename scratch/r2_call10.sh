#!/bin/bash
# Round 2, GPU call 10: rollout with the real policy (asserts compiled out for graph capture)
O=gpurun_out/r2_c10; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_rollout_policy.py tests/test_gpu_rollout.py -m gpu -q --tb=short ) > $O/tests.log 2>&1; echo "tests rc=$?" | tee $O/summary.txt; tail -8 $O/tests.log | cut -c1-300 | tee -a $O/summary.txt
( timeout 600 python scratch/bench_rollout.py ) > $O/rollout.txt 2>&1; echo "rollout rc=$?" | tee -a $O/summary.txt; tail -c 1200 $O/rollout.txt | tee -a $O/summary.txt
