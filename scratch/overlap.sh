for ov in 0 1 0 1; do echo "overlap=$ov"; PCT_B200_OVERLAP=$ov timeout 120 python scratch/ov_time.py 4096 1; done
for ov in 0 1; do echo "overlap=$ov"; PCT_B200_OVERLAP=$ov timeout 120 python scratch/ov_time.py 8192 2; done
for ov in 0 1; do echo "overlap=$ov"; PCT_B200_OVERLAP=$ov timeout 120 python scratch/ov_time.py 1024 1; done
