for cfg in "1 0" "1 1" "1 0" "1 1"; do set -- $cfg; echo "overlap=$1 lpt=$2"; PCT_B200_OVERLAP=$1 PCT_B200_LPT=$2 timeout 120 python scratch/ov_time.py 4096 1; done
for cfg in "1 0" "1 1"; do set -- $cfg; echo "overlap=$1 lpt=$2"; PCT_B200_OVERLAP=$1 PCT_B200_LPT=$2 timeout 120 python scratch/ov_time.py 8192 2; done
PCT_B200_LPT=1 timeout 300 python -m pytest tests/test_gpu_discrete_parity.py tests/test_gpu_discrete_cases.py -x -q 2>&1 | tail -2
