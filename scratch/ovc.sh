for ov in 0 1; do echo "overlap=$ov"; PCT_B200_OVERLAP=$ov timeout 120 python scratch/ov_c.py; done
for ov in 0 1; do echo "overlap=$ov (lpt default)"; PCT_B200_OVERLAP=$ov timeout 120 python scratch/ov_time.py 4096 1; done
timeout 300 python -m pytest tests/test_gpu_hostapi.py tests/test_gpu_continuous_parity.py tests/test_gpu_discrete_parity.py -x -q 2>&1 | tail -2
