python scratch/pcie.py
for cfg in "1 1" "1 4"; do set -- $cfg
  echo "prio=$1 host_groups=$2"; PCT_B200_HOST_PRIO=$1 PCT_B200_HOST_GROUPS=$2 timeout 120 python scratch/e2e_one.py 4096 1
done
timeout 120 python scratch/e2e_one.py 1024 1
